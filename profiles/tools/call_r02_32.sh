#!/bin/bash
# round 2, GPU call 32: split rollout, free-running vs event-staggered chains
cd /root/repo; mkdir -p gpurun_out
timeout 600 python profiles/tools/time_rollout_split.py marl-sortingenv_b200/csrc/libmsort.so 2>&1 | tee gpurun_out/r02_32_stagger.txt
