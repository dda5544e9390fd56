#!/bin/bash
# bench only (no tests): default build + every variant. usage: bash tools/gpu_bench_variants.sh TAG [steps]
TAG=${1:-run}; STEPS=${2:-6}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
for so in voxeltoy_b200/variants/*.so; do
    [ -f "$so" ] || continue
    name=$(basename $so .so)
    VT_LIB_PATH=$so timeout 600 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_${name}.json 2> $OUT/${TAG}_bench_${name}.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_bench*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), {k: round(v, 2) for k, v in j["roofline"]["kernel_ms_per_step"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
