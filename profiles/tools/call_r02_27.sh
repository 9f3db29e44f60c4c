#!/bin/bash
# round 2, GPU call 27: one-CTA-per-tile Env_2 tensor kernel at 8 / 7 / 6 CTAs per SM (64 / 72 / 80 registers); parity of the default
cd /root/repo; mkdir -p gpurun_out
V=marl-sortingenv_b200/csrc/variants; L=marl-sortingenv_b200/csrc/libmsort.so
{
for lib in $L $V/libmsort_tcnp7.so $V/libmsort_tcnp6.so; do
  for n in 1048576 262144; do N=$n TENSOR=1 timeout 200 python profiles/tools/time_variant.py $lib press 2>&1 | tail -1; done
done
} | tee gpurun_out/r02_27_tcnp_regs.txt
timeout 900 python -m pytest tests/test_tc_mlp_gpu.py tests/test_cuda_parity.py tests/test_cuda_edge_cases.py -x -q > gpurun_out/r02_27_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_27_tests.log
