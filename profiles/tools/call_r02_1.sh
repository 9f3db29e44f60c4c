#!/bin/bash
# round 2, GPU call 1: tensor-core embedded policy (Env_2) — correctness, then timing against the FFMA2 form
cd /root/repo; mkdir -p gpurun_out
L=marl-sortingenv_b200/csrc/libmsort.so
timeout 600 python -m pytest tests/test_tc_mlp_gpu.py -x -q -s > gpurun_out/r02_1_tc_tests.log 2>&1; echo "tc tests rc=$?"
tail -15 gpurun_out/r02_1_tc_tests.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_1_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/r02_1_gpu_tests.log
for n in 262144 1048576; do for t in 1 0; do
  N=$n TENSOR=$t timeout 300 python profiles/tools/time_variant.py $L press 2>&1 | tail -1
done; done | tee gpurun_out/r02_1_press_timing.txt
N=1048576 GAIN=1.0 TENSOR=1 timeout 300 python profiles/tools/time_variant.py $L press 2>&1 | tail -1 | tee -a gpurun_out/r02_1_press_timing.txt
timeout 300 python profiles/tools/time_variant.py $L mono 2>&1 | tail -1 | tee -a gpurun_out/r02_1_press_timing.txt
timeout 300 python profiles/tools/time_variant.py $L sort 2>&1 | tail -1 | tee -a gpurun_out/r02_1_press_timing.txt
