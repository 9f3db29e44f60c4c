#!/bin/bash
# round 2, GPU call 24: Env_2 (tensor-core policy) as S env ranges on S streams
cd /root/repo; mkdir -p gpurun_out
timeout 600 python profiles/tools/time_step_streams.py press 2>&1 | tee gpurun_out/r02_24_press_streams.txt
