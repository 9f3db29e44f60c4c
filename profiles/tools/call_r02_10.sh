#!/bin/bash
# round 2, GPU call 10: mbarrier wait flavour (suspend hint 4000 ns / 50 ns / test_wait spin) on the tensor-core MLP chain
cd /root/repo; mkdir -p gpurun_out
V=marl-sortingenv_b200/csrc/variants; L=marl-sortingenv_b200/csrc/libmsort.so
{
for lib in $L $V/libmsort_hint50.so $V/libmsort_hint0.so; do
  timeout 120 python profiles/tools/time_tc_logits.py $lib 2>&1 | tail -1
  N=1048576 TENSOR=1 timeout 200 python profiles/tools/time_variant.py $lib press 2>&1 | tail -1
done
} | tee gpurun_out/r02_10_hint.txt
