#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one `--set full` capture of the step kernel.
# usage: profiles/tools/capture.sh <tag> [bench args]
tag=${1:-rXX}; shift
cd /root/repo; mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-rollout $*"
$B > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu1_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 10 -c 1 -f -o gpurun_out/prof_$tag $B > gpurun_out/ncu2_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep
