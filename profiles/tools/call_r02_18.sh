#!/bin/bash
# round 2, GPU call 18: persistent rollout policy kernel with next-tile prefetch: tests + timing
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_fused_gpu.py tests/test_ppo_gpu.py -x -q > gpurun_out/r02_18_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02_18_tests.log
L=marl-sortingenv_b200/csrc/libmsort.so
timeout 200 python profiles/tools/time_rollout_policy.py $L 2>&1 | tail -1
N=262144 timeout 200 python profiles/tools/time_rollout_policy.py $L 2>&1 | tail -1
timeout 200 python profiles/tools/time_rollout_fused.py $L 2>&1 | tail -1
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_18_bench.json 2> gpurun_out/r02_18_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('/root/repo/gpurun_out/r02_18_bench.json').read().strip().splitlines()[-1])
r = d.get("rollout")
def show(name, x): print(name, round(x["value"] / 1e9, 2), "G", round(x["ms_per_step"] * 1e3, 1), "us")
show("rollout headline", r)
for k in ("split", "fused_one_kernel", "two_kernels", "two_kernels_r01"):
    if k in r: show(k, r[k])
PY
