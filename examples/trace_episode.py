#!/usr/bin/env python
"""Record one episode of a few device envs with the telemetry recorder and print / plot it.

    python examples/trace_episode.py [path/to/MARL-SortingEnv]

The recorder snapshots the traced envs on the GPU after every step and rebuilds the reference's own
Python logs from the snapshots (`reward_data`, `press_actions_per_timestep`, `bale_count`;
env_super.py:928-946, 631-637, 661-687).  With the reference checkout given (and matplotlib/seaborn
installed) the episode is rendered with the reference's dashboard, `utils.plotting.plot_env`
(plotting.py:28), exactly as `test_env` does at the end of an episode (testing.py:64-68).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marl_sortingenv_b200 as ms   # noqa: E402


def main():
    steps = 200
    env = ms.BatchedMonolithEnv(4096, max_steps=steps, seed=42, noise_sorting=0.0, auto_reset=False, info_level="full")
    rec = env.attach_trace([0, 1, 2], capacity=steps)      # before reset(): the recorder starts at the episode boundary
    env.reset(seed=42)
    for t in range(steps):
        env.step(env.rule_based_actions())                 # the reference's heuristic, evaluated on the device
    torch.cuda.synchronize()
    logs = rec.reference_logs(0)
    rd = logs["reward_data"]
    bales = {m: len(b) for m, b in logs["bale_count"].items()}
    print(f"env 0: return {sum(rd['Total']):.3f} over {len(rd['Total'])} steps; bales per material {bales}")
    print("first press-log entries:", logs["press_actions_per_timestep"][:12])
    if len(sys.argv) > 1:
        sys.path.insert(0, sys.argv[1])
        from utils.plotting import plot_env                # the reference's own dashboard
        plot_env(rec.reference_view(0), save=True, show=False, log_dir="./", filename="msort_trace", steps_test=steps)
        print("wrote ./msort_trace.svg")


if __name__ == "__main__":
    main()
