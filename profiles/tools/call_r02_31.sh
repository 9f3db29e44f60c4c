#!/bin/bash
# round 2, GPU call 31: Env_1's HOT kernel at 64 registers / 8 CTAs per SM (48 B of spills) vs 72 / 7
cd /root/repo; mkdir -p gpurun_out
V=marl-sortingenv_b200/csrc/variants; L=marl-sortingenv_b200/csrc/libmsort.so
{
for lib in $L $V/libmsort_sort8.so; do
  for n in 1048576 65536; do N=$n timeout 200 python profiles/tools/time_variant.py $lib sort 2>&1 | tail -1; done
done
} | tee gpurun_out/r02_31_sort8.txt
