#!/bin/bash
# round 2, GPU call 35: PPO gradient kernel with five row buffers (two CTAs per SM): parity tests + timing
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ppo_gpu.py -x -q > gpurun_out/r02_35_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_35_tests.log
timeout 600 python profiles/tools/time_ppo_update.py 2>&1 | tee gpurun_out/r02_35_ppo_update.txt
timeout 300 python examples/train_ppo.py 2048 20000000 2>&1 | tail -3
