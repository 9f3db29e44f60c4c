#!/bin/bash
# round 2, GPU call 22: Env_2 tensor kernel: actions loaded one tile ahead, weights by TMA: parity + timing
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_mlp_gpu.py tests/test_cuda_parity.py tests/test_cuda_edge_cases.py -x -q > gpurun_out/r02_22_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_22_tests.log
L=marl-sortingenv_b200/csrc/libmsort.so
for n in 1048576 262144; do N=$n TENSOR=1 timeout 200 python profiles/tools/time_variant.py $L press 2>&1 | tail -1; done | tee gpurun_out/r02_22_press.txt
