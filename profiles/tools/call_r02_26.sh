#!/bin/bash
# round 2, GPU call 26: Env_2 tensor kernel as one CTA per tile (MSORT_TC_PERSIST=0) vs persistent
cd /root/repo; mkdir -p gpurun_out
V=marl-sortingenv_b200/csrc/variants; L=marl-sortingenv_b200/csrc/libmsort.so
{
for lib in $L $V/libmsort_tcnp.so; do
  for n in 1048576 262144; do N=$n TENSOR=1 timeout 200 python profiles/tools/time_variant.py $lib press 2>&1 | tail -1; done
done
MSORT_LIB=$V/libmsort_tcnp.so timeout 300 python -m pytest tests/test_tc_mlp_gpu.py -x -q 2>&1 | tail -2
} | tee gpurun_out/r02_26_tcnp.txt
