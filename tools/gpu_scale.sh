#!/bin/bash
# One multi-GPU session on a box with N GPUs: the default bench line (C2 weak, with its strong-scaling C2 and tiled C4 extras) and,
# optionally, BASELINE config 5 and the group tests. usage: gpurun --gpus N -- 'bash tools/gpu_scale.sh TAG N [c5] [tests]'
TAG=${1:-scale}; N=${2:-8}; shift 2
OUT=gpurun_out; mkdir -p $OUT
PORT=29711
run() {
    local name=$1; shift
    PORT=$((PORT+1))
    NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $N --warmup 3 "$@" > $OUT/${TAG}_${name}_n${N}.json 2> $OUT/${TAG}_${name}_n${N}.err
    echo "$name n=$N rc=$?"
}
run c2weak --steps 5
for extra in "$@"; do
    case $extra in
        c5) run c5 --steps 2 --config c5 ;;
        c4) run c4 --steps 4 --config c4 ;;
        tests) timeout 600 python -m pytest tests/test_gpu_group.py -q -x --timeout 600 > $OUT/${TAG}_group_tests_n${N}.log 2>&1; tail -2 $OUT/${TAG}_group_tests_n${N}.log ;;
    esac
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*_n${N}.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "N", j["n_gpus"], round(j["value"], 1), j["scaling"], "ms/step", round(j["ms_per_step"], 2), "| exchange ms", (j.get("collective_ms") or {}).get("per_exchange_device"),
              "| render", round(j.get("ms_render_per_step") or 0, 2), "combine+wait", round(j.get("ms_combine_per_step_incl_wait") or 0, 2), "|", j.get("multi_gpu_check"))
        for k, v in (j.get("extras") or {}).items():
            print("   ", k, round(v["value"], 1) if v.get("value") else v.get("error"), "ms/step", v.get("ms_per_step"), "exchange", (v.get("collective_ms") or {}).get("per_exchange_device"), v.get("multi_gpu_check"))
    except Exception as e:
        print(f, "unreadable", e)
PY
