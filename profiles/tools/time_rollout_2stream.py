"""Experiment: the rollout loop (policy_act + step per env-step) on TWO streams over two half-batches, so that the
latency-bound tcgen05 policy kernel of one half shares the SMs with the ALU-bound step kernel of the other.
(A build with the persistent policy grid limited to 1..3 CTAs per SM, so that step CTAs fit beside it, was not faster.)
python profiles/tools/time_rollout_2stream.py"""
import os, sys
sys.path.insert(0, "/root/repo")
import torch
import marl_sortingenv_b200 as ms
from marl_sortingenv_b200.ppo import MaskableActorCritic, pack_actor_critic

N = 1 << 20
K = 64


def make(n, off):
    e = ms.BatchedMonolithEnv(n, max_steps=50, seed=42, info_level="none", track_stats=False, global_env_offset=off)
    e.reset()
    return e


def bufs(n):
    return (torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"))


def timed(g):
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K * 1e3)
    return best


torch.manual_seed(0)
pol = MaskableActorCritic(29, 22).cuda()
packed = pack_actor_critic(pol)

# one stream, whole batch
env = make(N, 0); out = bufs(N)
for t in range(5):
    a, _, _ = env.policy_act(packed, seed=1, t=t, out=out); env.step(a)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for t in range(K):
        a, _, _ = env.policy_act(packed, seed=1, t=t, out=out); env.step(a)
us = timed(g)
print(f"1 stream : {us:.1f} us/step  {N / us / 1e3:.2f} G env-steps/s", flush=True)
del g, env

# S streams, S sub-batches
for S in (2, 3, 4, 8):
    n = N // S // 128 * 128
    parts = [make(n, k * n) for k in range(S)]
    outs = [bufs(n) for _ in range(S)]
    streams = [torch.cuda.Stream() for _ in range(S)]
    for e, o in zip(parts, outs):
        for t in range(5):
            a, _, _ = e.policy_act(packed, seed=1, t=t, out=o); e.step(a)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for t in range(K):
            for s, e, o in zip(streams, parts, outs):
                with torch.cuda.stream(s):
                    a, _, _ = e.policy_act(packed, seed=1, t=t, out=o); e.step(a)
        for s in streams:
            cur.wait_stream(s)
    us = timed(g)
    print(f"{S} streams x {n} envs: {us:.1f} us/step  {S * n / us / 1e3:.2f} G env-steps/s", flush=True)
    del g, parts, outs
