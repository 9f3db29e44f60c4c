#!/usr/bin/env python
"""The reference's policy benchmark (utils/benchmark_models.py:126-171, plotted by
utils/benchmark_plot_summary.py) on the device: cumulative reward of several policies on Env_3_Monolith,
200 steps, noise 0, bale size 200, with and without action masking — here over 4 096 seeds per row in
about a second each instead of 10 seeds on one CPU env.

    python examples/benchmark_policies.py [ppo_timesteps]

Rows (all five of the reference's table): Random and Rule-Based are the reference's own `mode='random'` /
`mode='rule_based'` action sources (env_monolith.py:152-184), evaluated by kernels; the three PPO rows are
MaskablePPO-style agents trained on the spot with the GPU-resident loop (ppo.py) for `ppo_timesteps` each
(default 10 M; the reference trains 100 k on one env): "PPO Monolith" on Env_3; "PPO Sort-Only" = a sort agent
trained on Env_1 driving Env_3's sensor while the press action falls back to the random draw (mode='model'
with no press agent, env_monolith.py:213-219); "PPO Modular" = that sort agent plus a press agent trained on
Env_2 with the sort agent embedded (evaluated in-kernel by the step kernel's FFMA2 MLP), composed by
`ppo.modular_actions` exactly as mode='model' does (env_monolith.py:186-221).  Published values
(benchmark_plot_summary.py:5-18) are printed next to the measured ones.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marl_sortingenv_b200 as ms                                   # noqa: E402
from marl_sortingenv_b200.ppo import MaskablePPO, modular_actions, sort_policy_weights   # noqa: E402

PUBLISHED = {("Random", True): (-84.28, 22.29), ("Rule-Based", True): (44.03, 1.10), ("PPO Monolith", True): (32.77, 1.12),
             ("PPO Sort-Only", True): (-70.22, 10.56), ("PPO Modular", True): (30.61, 0.87),
             ("Random", False): (-109.36, 6.29), ("Rule-Based", False): (43.20, 1.07), ("PPO Monolith", False): (-100.31, 1.02),
             ("PPO Sort-Only", False): (-83.52, 10.14), ("PPO Modular", False): (-64.98, 7.92)}
N, STEPS = 4096, 200


@torch.no_grad()
def episode_returns(policy, masking: bool, model=None, sort_model=None, press_model=None):
    # the reference never sanitises a rule-based action, with or without masking (env_monolith.py:166-184 then
    # :252-262 "Rule-based or masked random: actions should be valid, convert directly"): its two Rule-Based
    # rows run the same code and differ only by seeds
    # ... and neither is a mode='model' choice (:258-262); only the external / unmasked-random action is sanitised
    sanitise = (not masking) and policy in ("Random", "PPO Monolith")
    env = ms.BatchedMonolithEnv(N, max_steps=STEPS, seed=2024 + int(masking), noise_sorting=0.0, balesize=200, auto_reset=False,
                                use_action_masking=not sanitise, info_level="none", track_stats=False)
    obs, _ = env.reset()
    total = torch.zeros(N, dtype=torch.float64, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(1)
    for t in range(STEPS):
        if policy == "Random":           # masked: uniform over valid actions; unmasked: uniform over all 22 (sanitised in step)
            a = env.sample_actions(seed=3, t=t) if masking else torch.randint(0, env.A, (N,), device="cuda", generator=gen)
        elif policy == "Rule-Based":
            a = env.rule_based_actions()
        elif policy == "PPO Sort-Only":  # mode='model', press agent missing: random press action (valid ones under masking)
            a = modular_actions(env, sort_model, None, use_action_masking=masking, seed=5, t=t)
        elif policy == "PPO Modular":
            a = modular_actions(env, sort_model, press_model, use_action_masking=masking)
        else:                            # trained agent: deterministic, masked logits only when masking is on
            m = env.action_masks() if masking else torch.ones((N, env.A), dtype=torch.bool, device="cuda")
            a, _ = model.predict(obs, action_masks=m, deterministic=True)
        obs, r, _, _, _ = env.step(a)
        total += r.double()
    env.close()
    return total.mean().item(), total.std().item()


def main():
    ppo_steps = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    kw = dict(max_steps=STEPS, seed=42, noise_sorting=0.0, info_level="none", track_stats=False)
    t0 = time.time()
    model = MaskablePPO(ms.BatchedMonolithEnv(2048, **kw), n_steps=64, batch_size=16384, n_epochs=10).learn(ppo_steps)
    sort_model = MaskablePPO(ms.BatchedSortingEnv(2048, **kw), n_steps=64, batch_size=16384, n_epochs=10).learn(ppo_steps)
    press_env = ms.BatchedPressingEnv(2048, **kw)
    press_env.set_agents(sort_agent=sort_policy_weights(sort_model))    # env_2_press.py:39-40: the trained sort agent, embedded
    press_model = MaskablePPO(press_env, n_steps=64, batch_size=16384, n_epochs=10).learn(ppo_steps)
    torch.cuda.synchronize()
    print(f"trained PPO Monolith / Sort / Press for {model.num_timesteps} timesteps each in {time.time() - t0:.1f} s\n")
    print(f"{'policy':14s} {'masking':8s} {'measured (4096 seeds)':>24s} {'published (10 seeds)':>24s}")
    for masking in (True, False):
        for policy in ("Random", "Rule-Based", "PPO Sort-Only", "PPO Modular", "PPO Monolith"):
            t1 = time.time()
            mean, std = episode_returns(policy, masking, model, sort_model, press_model)
            pm, ps = PUBLISHED[(policy, masking)]
            print(f"{policy:14s} {str(masking):8s} {mean:12.2f} +- {std:6.2f} {pm:14.2f} +- {ps:6.2f}   ({time.time() - t1:.2f} s)")


if __name__ == "__main__":
    main()
