#!/bin/bash
# round 2, GPU call 17: the stand-alone rollout policy kernel alone: time + ncu
cd /root/repo; mkdir -p gpurun_out
L=marl-sortingenv_b200/csrc/libmsort.so
timeout 200 python profiles/tools/time_rollout_policy.py $L > gpurun_out/r02_17_plain.log 2>&1 && cat gpurun_out/r02_17_plain.log &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_policy_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02_rpolicy python profiles/tools/time_rollout_policy.py $L > gpurun_out/r02_17_ncu.log 2>&1
echo "ncu rc=$?"
