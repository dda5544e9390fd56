#!/bin/bash
# Multi-GPU session: usage: gpurun --gpus N --timeout T -- 'bash tools/gpu_multi.sh TAG "2 4 8" [steps]'
TAG=${1:-multi}; NS=${2:-2}; STEPS=${3:-5}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,memory.used --format=csv > $OUT/${TAG}_smi.txt 2>&1
nvidia-smi topo -m >> $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_group.py -q -x --timeout 600 > $OUT/${TAG}_group_tests.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_group_tests.log; tail -4 $OUT/${TAG}_group_tests.log
PORT=29511
run() {  # n, name, extra flags...
    local n=$1 name=$2; shift 2
    PORT=$((PORT+1))
    NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $n --warmup 3 "$@" > $OUT/${TAG}_${name}_n${n}.json 2> $OUT/${TAG}_${name}_n${n}.err
    echo "$name n=$n rc=$?"; tail -c 300 $OUT/${TAG}_${name}_n${n}.err | tail -2
}
for n in $NS; do
    run $n c2weak --steps $STEPS
    run $n c2strong --steps $STEPS --config c2 --scaling strong
    run $n c4 --steps 3 --config c4
    run $n c5 --steps 2 --config c5
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*_n*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "N", j["n_gpus"], round(j["value"], 1), j["scaling"], "ms/step", round(j["ms_per_step"], 2), "| coll", (j.get("collective_ms") or {}).get("per_exchange_device"),
              "| render", j.get("ms_render_per_step"), "combine+wait", j.get("ms_combine_per_step_incl_wait"), "|", j.get("multi_gpu_check"), "| extras", {k: (round(v["value"], 1) if v.get("value") else v.get("error")) for k, v in (j.get("extras") or {}).items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
