bash tools/profile_round.sh r02_v23 > gpurun_out/r02_v23_profile.log 2>&1
bash tools/profile_c3.sh r02_c3_v23 >> gpurun_out/r02_v23_profile.log 2>&1
for k in wf_trace wf_shade wf_generate; do bash tools/ncu_summary.sh gpurun_out/r02_v23_$k.ncu-rep gpurun_out/r02_v23_$k >> gpurun_out/r02_v23_profile.log 2>&1; done
bash tools/ncu_summary.sh gpurun_out/r02_c3_v23_wf_trace.ncu-rep gpurun_out/r02_c3_v23_wf_trace >> gpurun_out/r02_v23_profile.log 2>&1
for k in wf_trace wf_shade wf_generate; do
ncu -i gpurun_out/r02_v23_$k.ncu-rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; rows=r[2:]
for i,n in enumerate(h):
    if 'smsp__pcsamp_warps_issue_stalled' in n and 'not_issued' not in n:
        print('%-60s %s' % (n.replace('smsp__pcsamp_warps_issue_stalled_',''), ' '.join('%9.0f' % float(x[i] or 0) for x in rows)))
" > gpurun_out/r02_v23_${k}_stalls.txt
done
rm -f gpurun_out/r02_v23_wf_generate.ncu-rep gpurun_out/r02_c3_v23_wf_trace.ncu-rep gpurun_out/r02_v23_wf_trace.ncu-rep gpurun_out/r02_v23_wf_shade.ncu-rep
tail -6 gpurun_out/r02_v23_profile.log
