#!/bin/bash
# round 2, GPU call 23: where a PPO iteration's time goes (rollout graph vs native update kernels)
cd /root/repo; mkdir -p gpurun_out
timeout 600 python profiles/tools/time_ppo_update.py 2>&1 | tee gpurun_out/r02_23_ppo_update.txt
