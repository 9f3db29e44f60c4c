#!/bin/bash
# ncu --set full of the tcgen05 policy kernel (run under gpurun)
cd /root/repo; mkdir -p gpurun_out
timeout 300 python profiles/tools/time_policy.py > gpurun_out/policy_plain.log 2>&1 || { tail -5 gpurun_out/policy_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:policy_act -s 5 -c 1 -f -o gpurun_out/prof_policy timeout 600 python profiles/tools/time_policy.py > gpurun_out/ncu_policy.log 2>&1
ls -la gpurun_out/prof_policy.ncu-rep
