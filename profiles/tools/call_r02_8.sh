#!/bin/bash
# round 2, GPU call 8: Env_2 tensor kernel at 8 CTAs/SM (27 KB smem, 64 regs): parity + timing vs 7 / 6 CTA builds
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_mlp_gpu.py tests/test_cuda_parity.py tests/test_cuda_edge_cases.py -x -q > gpurun_out/r02_8_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_8_tests.log
V=marl-sortingenv_b200/csrc/variants; L=marl-sortingenv_b200/csrc/libmsort.so
{
for lib in $L $V/libmsort_tc7.so $V/libmsort_tc6.so; do
  for n in 1048576 262144; do N=$n TENSOR=1 timeout 200 python profiles/tools/time_variant.py $lib press 2>&1 | tail -2; done
done
} | tee gpurun_out/r02_8_press.txt
