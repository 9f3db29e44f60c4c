"""python profiles/tools/time_rollout_fused.py <libmsort.so> — us per env-step of the fused rollout kernel (msort_rollout_step)
at N envs (default 1 048 576), CUDA graph of 64 steps; used for A/B of kernel variants and as the ncu target."""
import os, sys
os.environ["MSORT_LIB"] = sys.argv[1]
sys.path.insert(0, "/root/repo")
import torch
import marl_sortingenv_b200 as ms
from marl_sortingenv_b200.ppo import MaskableActorCritic, flatten_parameters, pack_actor_critic
n = int(os.environ.get("N", 1 << 20))
env = ms.BatchedMonolithEnv(n, max_steps=50, seed=42, info_level="episode")
torch.manual_seed(0)
pol = MaskableActorCritic(29, 22).cuda()
flat = flatten_parameters(pol); packed = pack_actor_critic(pol); pf = env.rollout_pack(flat)
outs = [(torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda")) for _ in range(2)]
env.reset()
env.policy_act(packed, seed=7, t=0, out=outs[0])
for t in range(8):
    env.rollout_step(outs[t % 2][0], pf, 7, t + 1, outs[(t + 1) % 2])
torch.cuda.synchronize()
K = 64
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for t in range(K):
        env.rollout_step(outs[t % 2][0], pf, 7, 9 + t, outs[(t + 1) % 2])
g.replay(); torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / K * 1e3)
print(f"{os.path.basename(sys.argv[1]):28s} fused rollout n {n} [{env.step_variant}]  us/step {best:.2f}  G/s {n/best/1e3:.2f}", flush=True)
