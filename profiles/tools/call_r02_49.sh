#!/bin/bash
# round 2, GPU call 49: what do the per-CTA statistics atomics cost the Env_3 step kernel today? (two ranges on two streams)
cd /root/repo; mkdir -p gpurun_out
{
echo "== info_level=episode (statistics + episode arrays: the bench configuration)"; timeout 300 python profiles/tools/time_step_streams.py mono 2>&1 | head -2
echo "== info_level=none, track_stats off"; NOSTATS=1 timeout 300 python profiles/tools/time_step_streams.py mono 2>&1 | head -2
} | tee gpurun_out/r02_49_stats_cost.txt
