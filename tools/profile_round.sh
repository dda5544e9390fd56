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
# one batch of the big kernels, full set: 4 wf_trace + 5 wf_shade launches (iterations 1..4 / 0..4 of a 64-pass batch) and one wf_generate
# launch (which also traces the primary rays); the skip counts pass the warm-up batches
for k in wf_trace:16:4 wf_shade:20:5 wf_generate:4:1; do
n=${k%%:*}; r=${k#*:}; s=${r%%:*}; c=${r#*:}
ncu --set full --clock-control none --import-source on -k regex:$n -s $s -c $c -o gpurun_out/${tag}_$n \
    python bench.py --passes 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_$n.log 2>&1
done
ls -la gpurun_out | grep ${tag}
