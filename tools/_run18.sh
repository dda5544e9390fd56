timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r02am_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02am_tests.log; tail -3 gpurun_out/r02am_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/r02am_bench.json 2> gpurun_out/r02am_bench.err; python - <<PY
import json
j=json.loads(open("gpurun_out/r02am_bench.json").read().strip().splitlines()[-1])
print(round(j["value"],1), round(j["e2e"]["value"],1), j["gpu_launches"], j["clocks"], j["cpu_baseline"]["value"])
PY
