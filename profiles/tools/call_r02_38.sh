#!/bin/bash
# round 2, GPU call 38: MaskablePPO at scale (131 072 envs: split rollout on two streams + native update), 400 M timesteps
cd /root/repo; mkdir -p gpurun_out
timeout 900 python - <<'PY' 2>&1 | tee gpurun_out/train_ppo_131072_r02.txt
import sys, time, torch
sys.path.insert(0, "/root/repo")
import marl_sortingenv_b200 as ms
from marl_sortingenv_b200.ppo import MaskablePPO, evaluate_policy
n = 131072
env = ms.BatchedMonolithEnv(n, max_steps=200, seed=42, noise_sorting=0.0, info_level="none", track_stats=False)
model = MaskablePPO(env, n_steps=32, batch_size=262144, n_epochs=10)
print("rollout form:", "split (policy kernel + step kernel, two streams)" if model.split_rollout else ("fused" if model.fused_rollout else "r01"), flush=True)
t0 = time.time()
print("untrained:", evaluate_policy(model, ms.BatchedMonolithEnv, noise_sorting=0.0), flush=True)
for k in range(5):
    model.learn((k + 1) * 80_000_000)
    torch.cuda.synchronize()
    mean, std = evaluate_policy(model, ms.BatchedMonolithEnv, noise_sorting=0.0)
    print(f"{model.num_timesteps:>11d} timesteps {time.time() - t0:6.1f} s   return {mean:7.2f} +- {std:5.2f}   finite {bool(torch.isfinite(model.flat).all())}", flush=True)
PY
