#!/bin/bash
# round 2, GPU call 45: how much of the PPO gradient kernel is the global atomics? (plain racing stores as a timing experiment)
cd /root/repo; mkdir -p gpurun_out
{
echo "== libmsort.so"; timeout 300 python profiles/tools/time_ppo_update.py 2>&1 | grep "gradient kernel"
echo "== no atomics (wrong sums)"; MSORT_LIB=marl-sortingenv_b200/csrc/variants/libmsort_noatomic.so timeout 300 python profiles/tools/time_ppo_update.py 2>&1 | grep "gradient kernel"
} | tee gpurun_out/r02_45_noatomic.txt
