#!/bin/bash
# round 2, GPU call 42: whole GPU suite after the last host-side changes
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_42_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/r02_42_gpu_tests.log
