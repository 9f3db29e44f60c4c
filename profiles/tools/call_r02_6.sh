#!/bin/bash
# round 2, GPU call 6: new parity tests (PHILOX vs reference fixture, K3 streams), whole GPU suite
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_6_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -6 gpurun_out/r02_6_gpu_tests.log
