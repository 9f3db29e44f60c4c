#!/bin/bash
# round 2, GPU call 14: ncu capture of the fused rollout kernel
cd /root/repo; mkdir -p gpurun_out
L=marl-sortingenv_b200/csrc/libmsort.so
timeout 200 python profiles/tools/time_rollout_fused.py $L > gpurun_out/r02_14_plain.log 2>&1 && cat gpurun_out/r02_14_plain.log &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -f -o gpurun_out/prof_r02_fused python profiles/tools/time_rollout_fused.py $L > gpurun_out/r02_14_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_14_ncu.log
