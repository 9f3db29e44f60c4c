#!/bin/bash
# round 2, GPU call 13: fused rollout kernel (step + next step's policy): parity tests, PPO tests, rollout timing
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_fused_gpu.py -x -q > gpurun_out/r02_13_fused_tests.log 2>&1; echo "fused tests rc=$?"; tail -25 gpurun_out/r02_13_fused_tests.log
timeout 900 python -m pytest tests/test_ppo_gpu.py tests/test_policy_act_gpu.py -x -q > gpurun_out/r02_13_ppo_tests.log 2>&1; echo "ppo tests rc=$?"; tail -5 gpurun_out/r02_13_ppo_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_13_bench.json 2> gpurun_out/r02_13_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('/root/repo/gpurun_out/r02_13_bench.json').read().strip().splitlines()[-1])
print("value", d["value"] / 1e9, "G; rollout", json.dumps(d.get("rollout"))[:900])
PY
tail -3 gpurun_out/r02_13_bench.err
