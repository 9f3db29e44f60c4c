#!/bin/bash
# round 2, GPU call 9: what saturates Env_2's tensor kernel?  MUFU microbenchmark + sigmoid ablations + no-policy floor
cd /root/repo; mkdir -p gpurun_out
./profiles/tools/mufu_bench | tee gpurun_out/r02_9_mufu.txt
V=marl-sortingenv_b200/csrc/variants; L=marl-sortingenv_b200/csrc/libmsort.so
{
for lib in $L $V/libmsort_sig1.so $V/libmsort_sig2.so; do
  N=1048576 TENSOR=1 timeout 200 python profiles/tools/time_variant.py $lib press 2>&1 | tail -1
done
N=1048576 NOPOLICY=1 timeout 200 python profiles/tools/time_variant.py $L press 2>&1 | tail -1
} | tee gpurun_out/r02_9_ablate.txt
