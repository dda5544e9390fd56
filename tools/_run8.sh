run() {
  name=$1; shift
  env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02ac_$name.json 2> gpurun_out/r02ac_$name.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02ac_$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"],1), round(j["e2e"]["value"],1), {k:round(x,1) for k,x in j["roofline"]["kernel_ms_per_step"].items()})
except Exception as e: print("$name failed", e)
PY
}
run base VT_X=1
run l2 VT_WF_LANES=2
run l2_s4t3 VT_WF_LANES=2 VT_WF_SHADE_CTAS=4 VT_WF_TRACE_CTAS=3
run l2_s5t3 VT_WF_LANES=2 VT_WF_SHADE_CTAS=5 VT_WF_TRACE_CTAS=3
run l2_s4t2 VT_WF_LANES=2 VT_WF_SHADE_CTAS=4 VT_WF_TRACE_CTAS=2
run l2_s6t4 VT_WF_LANES=2 VT_WF_SHADE_CTAS=6 VT_WF_TRACE_CTAS=4
run l3_s3t2 VT_WF_LANES=3 VT_WF_SHADE_CTAS=3 VT_WF_TRACE_CTAS=2
run l4_s2t2 VT_WF_LANES=4 VT_WF_SHADE_CTAS=2 VT_WF_TRACE_CTAS=2
run l4_s3t2 VT_WF_LANES=4 VT_WF_SHADE_CTAS=3 VT_WF_TRACE_CTAS=2
