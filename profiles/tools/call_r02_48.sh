#!/bin/bash
# round 2, GPU call 48: whole GPU suite + the two training runs with the faster update kernels
cd /root/repo; mkdir -p gpurun_out
date; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_48_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02_48_gpu_tests.log
timeout 300 python examples/train_ppo.py 2048 20000000 > gpurun_out/train_ppo_r02.txt 2>&1; tail -3 gpurun_out/train_ppo_r02.txt
bash profiles/tools/call_r02_38.sh 2>&1 | tail -7
