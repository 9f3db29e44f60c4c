#!/bin/bash
# round 2, GPU call 11: cycle stamps inside the tensor-core MLP chain at 1 / 4 / 8 CTAs per SM
cd /root/repo; mkdir -p gpurun_out
V=marl-sortingenv_b200/csrc/variants
for c in 1 4 8; do TCPROF=1 timeout 120 python profiles/tools/time_tc_logits.py $V/libmsort_prof$c.so 2>&1 | tail -2; done | tee gpurun_out/r02_11_tcprof.txt
