#!/bin/bash
# round 2, GPU call 7: native PPO update kernels + graph rollout tests, then training throughput
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ppo_gpu.py -x -q > gpurun_out/r02_7_ppo_tests.log 2>&1; echo "ppo tests rc=$?"; tail -30 gpurun_out/r02_7_ppo_tests.log
timeout 300 python examples/train_ppo.py 2048 20000000 > gpurun_out/r02_7_train.txt 2>&1; echo "train rc=$?"; tail -8 gpurun_out/r02_7_train.txt
# fresh ncu capture of Env_2's tensor-core step kernel (1 048 576 envs, 6 CTAs/SM)
L=marl-sortingenv_b200/csrc/libmsort.so
N=1048576 TENSOR=1 timeout 200 python profiles/tools/time_variant.py $L press > gpurun_out/r02_7_press_plain.log 2>&1 &&
N=1048576 TENSOR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 100 -c 1 -f -o gpurun_out/prof_r02_press_tc python profiles/tools/time_variant.py $L press > gpurun_out/r02_7_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_7_press_plain.log
