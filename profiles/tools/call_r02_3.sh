#!/bin/bash
# round 2, GPU call 3: occupancy fix of the tensor-core Env_2 kernel (CTAs-per-SM sweep) + native step_host test
cd /root/repo; mkdir -p gpurun_out
L=marl-sortingenv_b200/csrc/libmsort.so
timeout 600 python -m pytest tests/test_api_surfaces.py tests/test_tc_mlp_gpu.py tests/test_cuda_edge_cases.py -x -q -m gpu > gpurun_out/r02_3_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_3_tests.log
{
N=1048576 TENSOR=1 timeout 200 python profiles/tools/time_variant.py $L press 2>&1 | tail -2
for c in 1 2 3 4 5; do N=1048576 TENSOR=1 CTAS=$c timeout 200 python profiles/tools/time_variant.py $L press 2>&1 | tail -1; done
N=262144 TENSOR=1 timeout 200 python profiles/tools/time_variant.py $L press 2>&1 | tail -1
N=1048576 TENSOR=0 timeout 200 python profiles/tools/time_variant.py $L press 2>&1 | tail -2
} | tee gpurun_out/r02_3_ctas.txt
