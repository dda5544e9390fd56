#!/bin/bash
# Profiling recipe of B200_PROFILING.md for this repo, run under gpurun (1 GPU):
#   tools/profile_round.sh <tag>     -> gpurun_out/<tag>_{bench.json,launches.csv,wf_trace.ncu-rep,wf_shade.ncu-rep}
set -u
tag=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 1500 gpurun_out/${tag}_bench.json
# every launch of one 64-pass batch (--passes 64: a bench step is 4 such batches of 128 Mi paths) with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 120 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --passes 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
# one batch (5 launches) of the two big kernels, full set
for k in wf_generate wf_trace wf_shade; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 5 -o gpurun_out/${tag}_$k \
    python bench.py --passes 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_$k.log 2>&1
done
ls -la gpurun_out | grep ${tag}
