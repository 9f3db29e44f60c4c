"""python profiles/tools/time_rollout_policy.py <libmsort.so> — us per launch of msort_rollout_policy at N envs (default
1 048 576): the rollout's policy half alone (tiles from HBM); the ncu target for that kernel."""
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
env.reset()
a = torch.zeros(n, dtype=torch.int64, device="cuda")
for t in range(5):
    env.sample_actions(3, t, out=a); env.step(a)
out = (torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"))
for t in range(3):
    env.rollout_policy(pf, seed=7, t=t, out=out)
torch.cuda.synchronize()
K = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(K):
    env.rollout_policy(pf, seed=7, t=t, out=out)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / K * 1e3
print(f"{os.path.basename(sys.argv[1]):28s} rollout_policy n {n}: {us:.2f} us  ({n * 154 / us / 1e3:.0f} GB/s of 138 B in + 16 B out per env)", flush=True)
