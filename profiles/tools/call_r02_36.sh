#!/bin/bash
# round 2, GPU call 36: Env_2 split form (policy kernel + policy-free step kernel) vs the fused tensor kernel: parity + timing
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_mlp_gpu.py tests/test_cuda_parity.py tests/test_cuda_edge_cases.py -x -q > gpurun_out/r02_36_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_36_tests.log
L=marl-sortingenv_b200/csrc/libmsort.so
{
for tp in 2 1; do for n in 1048576 262144; do N=$n TENSOR=$tp timeout 200 python profiles/tools/time_variant.py $L press 2>&1 | tail -1; done; done
timeout 600 python profiles/tools/time_step_streams.py press 2>&1 | head -2
} | tee gpurun_out/r02_36_split.txt
