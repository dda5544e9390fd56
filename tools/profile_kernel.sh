#!/bin/bash
# ncu --set full capture of one 64-pass batch (5 launches) of one wavefront kernel: tools/profile_kernel.sh <tag> <wf_trace|wf_shade|wf_generate>
set -u
tag=$1; k=$2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 5 -o gpurun_out/${tag}_${k} \
    python bench.py --passes 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_${k}.log 2>&1
ls -la gpurun_out/${tag}_${k}.ncu-rep
