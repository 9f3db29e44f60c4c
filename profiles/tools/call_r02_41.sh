#!/bin/bash
# round 2, GPU call 41: archive round trip test + whole GPU suite
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_41_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -8 gpurun_out/r02_41_gpu_tests.log
