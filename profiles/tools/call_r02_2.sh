#!/bin/bash
# round 2, GPU call 2: why is the tensor-core Env_2 kernel slow?  MMA-count experiment + ncu capture
cd /root/repo; mkdir -p gpurun_out
V=marl-sortingenv_b200/csrc/variants; L=marl-sortingenv_b200/csrc/libmsort.so
{
for lib in $L $V/libmsort_terms3.so $V/libmsort_terms1.so; do timeout 120 python profiles/tools/time_tc_logits.py $lib 2>&1 | tail -1; done
for lib in $L $V/libmsort_terms3.so $V/libmsort_terms1.so; do N=262144 TENSOR=1 timeout 200 python profiles/tools/time_variant.py $lib press 2>&1 | tail -1; done
} | tee gpurun_out/r02_2_terms.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_2_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02_2_gpu_tests.log
export N=262144 TENSOR=1
timeout 200 python profiles/tools/time_variant.py $L press > gpurun_out/r02_2_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 20 -c 1 -f -o gpurun_out/prof_r02_tc python profiles/tools/time_variant.py $L press > gpurun_out/r02_2_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02_2_ncu.log; ls -la gpurun_out/*.ncu-rep
