#!/bin/bash
# ncu --set full of wf_trace on BASELINE config 3 (bunny 512^3, empty-space skip on): tools/profile_c3.sh <tag>
tag=${1:-r02_c3}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:wf_trace -s 8 -c 4 -o gpurun_out/${tag}_wf_trace \
    python tools/run_configs.py c3 > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/${tag}_wf_trace.ncu-rep
