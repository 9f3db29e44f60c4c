"""Experiment: K back-to-back step() launches of one handle as S env ranges on S streams (msort_step_range), so the
tail of one range's kernel overlaps the head of the other's.  python profiles/tools/time_step_streams.py [kind]"""
import sys
sys.path.insert(0, "/root/repo")
import torch
import marl_sortingenv_b200 as ms

kind = sys.argv[1] if len(sys.argv) > 1 else "mono"
n = 1 << 20
cls = {"mono": ms.BatchedMonolithEnv, "sort": ms.BatchedSortingEnv, "press": ms.BatchedPressingEnv}[kind]
import os
env = cls(n, max_steps=50, seed=42, info_level="none", track_stats=False) if os.environ.get("NOSTATS") else cls(n, max_steps=50, seed=42, info_level="episode")
if kind == "press":
    from marl_sortingenv_b200.policy import sb3_style_init
    env.set_sort_policy(sb3_style_init(0))
env.reset()
T = 128
acts = torch.zeros((T, n), dtype=torch.int64, device="cuda")
for t in range(T):
    env.sample_actions(7, t, out=acts[t]); env.step(acts[t])
torch.cuda.synchronize()
for S in (1, 2, 3, 4):
    per = -(-(n // S) // 128) * 128
    ranges = [(lo, min(n, lo + per)) for lo in range(0, n, per)]
    streams = [torch.cuda.Stream() for _ in ranges]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for t in range(512):
            for s, r in zip(streams, ranges):
                with torch.cuda.stream(s):
                    env.step(acts[t % T], env_range=r if S > 1 else None)
        for s in streams:
            cur.wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 512 * 1e3)
    print(f"{kind} {S} range(s)/stream(s): us/step {best:.2f}  G/s {n / best / 1e3:.2f}", flush=True)
    del g
