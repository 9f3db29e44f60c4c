#!/bin/bash
# round 2, GPU call 40: msort_policy_eval (strided rows) + modular agents on the kernel; policy / api tests
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_act_gpu.py tests/test_api_surfaces.py tests/test_ppo_gpu.py -x -q > gpurun_out/r02_40_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02_40_tests.log
