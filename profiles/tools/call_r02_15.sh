#!/bin/bash
# round 2, GPU call 15: fused rollout kernel after the epilogue cuts: tests + timing + ncu
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_fused_gpu.py -x -q > gpurun_out/r02_15_fused_tests.log 2>&1; echo "fused tests rc=$?"; tail -15 gpurun_out/r02_15_fused_tests.log
L=marl-sortingenv_b200/csrc/libmsort.so
timeout 200 python profiles/tools/time_rollout_fused.py $L > gpurun_out/r02_15_plain.log 2>&1 && cat gpurun_out/r02_15_plain.log &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -f -o gpurun_out/prof_r02_fused2 python profiles/tools/time_rollout_fused.py $L > gpurun_out/r02_15_ncu.log 2>&1
echo "ncu rc=$?"
