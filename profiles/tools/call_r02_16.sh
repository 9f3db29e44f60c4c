#!/bin/bash
# round 2, GPU call 16: stand-alone rollout policy kernel (the fused kernel's policy half), split vs fused rollout timing
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_fused_gpu.py tests/test_ppo_gpu.py -x -q > gpurun_out/r02_16_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02_16_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_16_bench.json 2> gpurun_out/r02_16_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('/root/repo/gpurun_out/r02_16_bench.json').read().strip().splitlines()[-1])
r = d.get("rollout")
def show(name, x): print(name, round(x["value"] / 1e9, 2), "G", round(x["ms_per_step"] * 1e3, 1), "us")
show("rollout headline", r)
for k in ("split", "fused_one_kernel", "two_kernels", "two_kernels_r01"):
    if k in r: show(k, r[k])
PY
tail -3 gpurun_out/r02_16_bench.err
