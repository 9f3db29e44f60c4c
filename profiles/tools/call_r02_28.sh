#!/bin/bash
# round 2, GPU call 28: one-CTA-per-tile Env_2 tensor kernel without the set-up barrier: parity + timing + bench entries
cd /root/repo; mkdir -p gpurun_out
L=marl-sortingenv_b200/csrc/libmsort.so
for n in 1048576 262144; do N=$n TENSOR=1 timeout 200 python profiles/tools/time_variant.py $L press 2>&1 | tail -1; done | tee gpurun_out/r02_28_press.txt
timeout 900 python -m pytest tests/test_tc_mlp_gpu.py tests/test_cuda_parity.py tests/test_cuda_edge_cases.py -x -q > gpurun_out/r02_28_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_28_tests.log
timeout 600 python profiles/tools/time_step_streams.py press 2>&1 | head -2 | tee -a gpurun_out/r02_28_press.txt
