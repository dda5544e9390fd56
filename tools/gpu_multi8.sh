#!/bin/bash
# 8-GPU session: group tests, the default bench line (with its strong-scaling C2 / tiled C4 extras) and BASELINE config 5 at
# N = 2, 4, 8, and the C++ RendererGroup demo on all GPUs. usage: gpurun --gpus 8 -- 'bash tools/gpu_multi8.sh TAG "2 4 8" steps'
TAG=${1:-m8}; NS=${2:-"2 4 8"}; STEPS=${3:-5}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,memory.used --format=csv > $OUT/${TAG}_smi.txt 2>&1
nvidia-smi topo -m >> $OUT/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_group.py -q -x --timeout 600 > $OUT/${TAG}_group_tests.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_group_tests.log; tail -3 $OUT/${TAG}_group_tests.log
PORT=29611
run() {
    local n=$1 name=$2; shift 2
    PORT=$((PORT+1))
    NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $n --warmup 3 "$@" > $OUT/${TAG}_${name}_n${n}.json 2> $OUT/${TAG}_${name}_n${n}.err
    echo "$name n=$n rc=$?"
}
for n in $NS; do
    run $n c2weak --steps $STEPS
    run $n c5 --steps 2 --config c5
done
# the C++ host program (no Python): RendererGroup over all GPUs, both partitions, both exchanges
python - <<PY
import gzip, os, numpy as np, sys
sys.path.insert(0, ".")
from voxeltoy_b200 import host, scenes
open("/tmp/scene_fall.vox", "wb").write(gzip.open("tests/golden/scene_fall.vox.gz").read())
host.write_pfm("/tmp/env.pfm", scenes.synthetic_env(1024, 512))
PY
NG=$(nvidia-smi -L | wc -l); DEVS=$(seq -s, 0 $((NG-1)))
for mode in samples tiles; do for ex in nccl peer; do
    timeout 300 ./voxeltoy_b200/vt_group_demo --vox /tmp/scene_fall.vox --env /tmp/env.pfm --devices $DEVS --mode $mode --exchange $ex --passes 256 --steps 3 \
        >> $OUT/${TAG}_cpp_demo.jsonl 2>> $OUT/${TAG}_cpp_demo.err
done; done
cat $OUT/${TAG}_cpp_demo.jsonl
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*_n*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "N", j["n_gpus"], round(j["value"], 1), j["scaling"], "ms/step", round(j["ms_per_step"], 2), "| coll", (j.get("collective_ms") or {}).get("per_exchange_device"),
              "| render", j.get("ms_render_per_step"), "combine+wait", j.get("ms_combine_per_step_incl_wait"), "|", j.get("multi_gpu_check"),
              "| extras", {k: (round(v["value"], 1) if v.get("value") else v.get("error")) for k, v in (j.get("extras") or {}).items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
