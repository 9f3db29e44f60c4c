#!/bin/bash
# round 2, GPU call 43/44: PPO gradient kernel variants: parity tests + timing
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ppo_gpu.py -x -q > gpurun_out/r02_43_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_43_tests.log
timeout 600 python profiles/tools/time_ppo_update.py 2>&1 | tee gpurun_out/r02_43_ppo_update.txt
