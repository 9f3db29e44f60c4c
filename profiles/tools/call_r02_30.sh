#!/bin/bash
# round 2, GPU call 30: rolled MLP epilogues in the Env_2 tensor kernel (instruction footprint) vs unrolled (current libmsort.so)
cd /root/repo; mkdir -p gpurun_out
V=marl-sortingenv_b200/csrc/variants; L=marl-sortingenv_b200/csrc/libmsort.so
{
for lib in $L $V/libmsort_tcroll.so; do
  for n in 1048576 262144; do N=$n TENSOR=1 timeout 200 python profiles/tools/time_variant.py $lib press 2>&1 | tail -1; done
done
} | tee gpurun_out/r02_30_tcroll.txt
