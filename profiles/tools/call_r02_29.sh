#!/bin/bash
# round 2, GPU call 29: final evidence refresh after the Env_2 kernel change: whole GPU suite, bench, Env_2 ncu capture, workloads
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_29_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02_29_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_29_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_29_smoke.log
timeout 900 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r02.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_r02_reference_arm.json 2> gpurun_out/bench_r02_reference_arm.err; echo "bench ref rc=$?"
L=marl-sortingenv_b200/csrc/libmsort.so
N=1048576 TENSOR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 100 -c 1 -f -o gpurun_out/prof_r02_press_tc_final python profiles/tools/time_variant.py $L press > gpurun_out/r02_29_ncu_press.log 2>&1; echo "ncu press rc=$?"
{
for k in sort press mono; do for n in 65536 262144 1048576; do N=$n timeout 200 python profiles/tools/time_variant.py $L $k 2>&1 | tail -1; done; done
N=1048576 NOPOLICY=1 timeout 200 python profiles/tools/time_variant.py $L press 2>&1 | tail -1
N=1048576 TENSOR=0 timeout 200 python profiles/tools/time_variant.py $L press 2>&1 | tail -1
timeout 200 python profiles/tools/time_rollout_policy.py $L 2>&1 | tail -1
timeout 200 python profiles/tools/time_rollout_fused.py $L 2>&1 | tail -1
} | tee gpurun_out/other_workloads_r02.txt
