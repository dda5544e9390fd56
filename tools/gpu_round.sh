#!/bin/bash
# One GPU-box session: the parity suite, then the bench on the default build and on every A/B variant in
# voxeltoy_b200/variants/. Usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG [steps]'
TAG=${1:-run}; STEPS=${2:-6}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py tests/test_gpu_edges.py tests/test_gpu_renderer.py \
    tests/test_gpu_env.py tests/test_gpu_group.py tests/test_gpu_headless.py tests/test_gpu_configs.py tests/test_gpu_fullsize.py -q -x --timeout 600 > $OUT/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_tests.log
tail -5 $OUT/${TAG}_tests.log
timeout 600 python bench.py --steps $STEPS --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench.err
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
