#!/bin/bash
# round 2, GPU call 34: ncu of the PPO forward+backward kernel (ppo_kernel<29,22,true>)
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppo_kernel -s 30 -c 1 -f -o gpurun_out/prof_r02_ppo python profiles/tools/time_ppo_update.py > gpurun_out/r02_34_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_34_ncu.log
