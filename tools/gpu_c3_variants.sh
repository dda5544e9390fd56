#!/bin/bash
# C3 (and C4 / C5) on the default build and on every variant: bash tools/gpu_c3_variants.sh TAG "c3 c4 c5"
TAG=${1:-c3v}; CFGS=${2:-c3}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_configs.py -q -x -k "skip or advance or c3" > $OUT/${TAG}_tests.log 2>&1; tail -2 $OUT/${TAG}_tests.log
python tools/run_configs.py $CFGS > $OUT/${TAG}_default.jsonl 2>> $OUT/${TAG}.err
for so in voxeltoy_b200/variants/*.so; do
    [ -f "$so" ] || continue
    VT_LIB_PATH=$so python tools/run_configs.py $CFGS > $OUT/${TAG}_$(basename $so .so).jsonl 2>> $OUT/${TAG}.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.jsonl")):
    for l in open(f):
        try:
            j = json.loads(l); print(f, j["config"][:12], round(j["msamples_per_s"], 1), j.get("kernel_ms"))
        except Exception as e:
            print(f, "unreadable", e)
PY
