#!/usr/bin/env python
"""Train a MaskablePPO-style agent on the batched Monolith env entirely on the GPU.

    python examples/train_ppo.py [num_envs] [total_timesteps]

Protocol of the reference (main.py:42-52): max_steps=200, noise 0, bale size 200; evaluation =
deterministic masked policy, 200 steps.  Published returns of the reference (CPU, 100 000
timesteps): PPO Monolith 32.77 +- 1.12, Rule-Based 44.03 +- 1.10 (utils/benchmark_plot_summary.py).
On one B200 2 048 envs x 20 M timesteps take ~2.5 s and reach ~85 (round 1: ~30 s): the rollout is one CUDA graph of fused
policy + step kernels, the update four hand-written kernels per minibatch (marl-sortingenv_b200/csrc/msort_ppo.cu).
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marl_sortingenv_b200 as ms                                   # noqa: E402
from marl_sortingenv_b200.ppo import MaskablePPO, evaluate_policy   # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    total = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20_000_000
    env = ms.BatchedMonolithEnv(n, max_steps=200, seed=42, noise_sorting=0.0, info_level="none", track_stats=False)
    model = MaskablePPO(env, n_steps=64, batch_size=16384, n_epochs=10)
    t0 = time.time()
    print("untrained:", evaluate_policy(model, ms.BatchedMonolithEnv, noise_sorting=0.0))
    for k in range(5):
        model.learn((k + 1) * total // 5)
        torch.cuda.synchronize()
        mean, std = evaluate_policy(model, ms.BatchedMonolithEnv, noise_sorting=0.0)
        print(f"{model.num_timesteps:>11d} timesteps {time.time() - t0:6.1f} s   return {mean:7.2f} +- {std:5.2f}")


if __name__ == "__main__":
    main()
