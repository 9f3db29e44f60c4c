"""python profiles/tools/time_rollout_split.py <libmsort.so> — the rollout as msort_rollout_policy + msort_step over two env ranges on
two streams, (a) free-running chains, (b) staggered by events so that one range's policy kernel (XU-bound) always runs beside the
other range's step kernel (ALU-bound): us per env-step at N envs, CUDA graph of 64 steps."""
import os, sys
os.environ["MSORT_LIB"] = sys.argv[1]
sys.path.insert(0, "/root/repo")
import torch
import marl_sortingenv_b200 as ms
from marl_sortingenv_b200.ppo import MaskableActorCritic, flatten_parameters
n = int(os.environ.get("N", 1 << 20))
env = ms.BatchedMonolithEnv(n, max_steps=50, seed=42, info_level="episode")
torch.manual_seed(0)
pol = MaskableActorCritic(29, 22).cuda()
pf = env.rollout_pack(flatten_parameters(pol))
out = (torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"))
env.reset()
for t in range(8):
    env.rollout_policy(pf, seed=7, t=t, out=out); env.step(out[0])
torch.cuda.synchronize()
half = (n // 2 + 127) // 128 * 128
K = 64
for mode in ("free", "staggered", "staggered4"):
    nr = 4 if mode == "staggered4" else 2
    per = (n // nr + 127) // 128 * 128
    ranges = [(lo, min(n, lo + per)) for lo in range(0, n, per)]
    streams = [torch.cuda.Stream() for _ in ranges]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        prev_policy_done = None
        for t in range(K):
            for k, (s, r) in enumerate(zip(streams, ranges)):
                with torch.cuda.stream(s):
                    if mode != "free" and prev_policy_done is not None:
                        s.wait_event(prev_policy_done)            # policies never run side by side: each one beside another range's step
                    env.rollout_policy(pf, seed=7, t=8 + t, out=out, env_range=r)
                    if mode != "free":
                        prev_policy_done = torch.cuda.Event(); prev_policy_done.record(s)
                    env.step(out[0], env_range=r)
        for s in streams:
            cur.wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K * 1e3)
    print(f"{os.path.basename(sys.argv[1]):20s} split rollout n {n} [{mode}, {len(ranges)} ranges]  us/step {best:.2f}  G/s {n / best / 1e3:.2f}", flush=True)
    del g
