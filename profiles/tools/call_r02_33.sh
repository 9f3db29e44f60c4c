#!/bin/bash
# round 2, GPU call 33: new tests (fused statistics, PPO on Env_1 / Env_2) + the whole GPU suite on the final library
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_33_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -12 gpurun_out/r02_33_gpu_tests.log
