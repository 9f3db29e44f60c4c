"""us/step of Env_3 at max_steps=200 (the reference's episode length: levels up to 20 000, HOT kernel without the
small-level shortcut) next to max_steps=50: python profiles/tools/time_maxsteps.py [libmsort.so]"""
import os, sys
if len(sys.argv) > 1: os.environ["MSORT_LIB"] = sys.argv[1]
sys.path.insert(0, "/root/repo")
import torch
import marl_sortingenv_b200 as ms
n = 1 << 20
for max_steps in (50, 200):
    env = ms.BatchedMonolithEnv(n, max_steps=max_steps, seed=42, info_level="episode")
    env.reset()
    T = 256
    acts = torch.zeros((T, n), dtype=torch.int64, device="cuda")
    for t in range(T):
        env.sample_actions(7, t, out=acts[t]); env.step(acts[t])
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for t in range(512):
            env.step(acts[t % T])
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 512 * 1e3)
    print(f"max_steps {max_steps:4d}: {env.step_variant:14s} us/step {best:.2f}  G/s {n / best / 1e3:.2f}", flush=True)
    env.close()
