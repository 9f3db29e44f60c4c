#!/bin/bash
# round 2, GPU call 21: Env_2 as one CTA per tile (MSORT_HOT_PERSIST=0) vs the persistent loop, with and without the FFMA2 policy
cd /root/repo; mkdir -p gpurun_out
V=marl-sortingenv_b200/csrc/variants; L=marl-sortingenv_b200/csrc/libmsort.so
{
for lib in $L $V/libmsort_nopersist.so; do
  N=1048576 NOPOLICY=1 timeout 200 python profiles/tools/time_variant.py $lib press 2>&1 | tail -1
  N=262144 NOPOLICY=1 timeout 200 python profiles/tools/time_variant.py $lib press 2>&1 | tail -1
  N=1048576 TENSOR=0 timeout 200 python profiles/tools/time_variant.py $lib press 2>&1 | tail -1
done
} | tee gpurun_out/r02_21_nopersist.txt
