import os, sys, torch, time
if len(sys.argv) > 1: os.environ["MSORT_LIB"] = sys.argv[1]
sys.path.insert(0, "/root/repo")
import marl_sortingenv_b200 as ms
from marl_sortingenv_b200.ppo import MaskablePPO, MaskableActorCritic, pack_actor_critic
n = 1 << 20
env = ms.BatchedMonolithEnv(n, max_steps=50, seed=42, info_level="none", track_stats=False)
env.reset()
pol = MaskableActorCritic(env.D, env.A).cuda()
packed = pack_actor_critic(pol)
for t in range(20):
    a, lp, v = env.policy_act(packed, seed=1, t=t); env.step(a)
torch.cuda.synchronize()
with torch.no_grad():
    a, lp, v = env.policy_act(packed, seed=1, t=99)
    ref = torch.log_softmax(pol.masked_logits(env.obs, env.mask), dim=-1).gather(1, a[:, None]).squeeze(1)
    print("max |logp - torch fp32|:", float((lp - ref).abs().max()), " max |value - torch fp32|:", float((v - pol.vf(env.obs).squeeze(1)).abs().max()),
          " mean |dlogp|:", float((lp - ref).abs().mean()))
def timeit(f, k=50):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(k): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3
out = (torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"))
print("policy_act only   us:", timeit(lambda: env.policy_act(packed, seed=1, t=3, out=out)))
def loop_fused():
    a, _, _ = env.policy_act(packed, seed=1, t=3, out=out); env.step(a)
print("policy_act + step us:", timeit(loop_fused))
@torch.no_grad()
def loop_torch():
    a, lp, v = pol.act(env.obs, env.mask); env.step(a)
print("torch act + step  us:", timeit(loop_torch, 20))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(50): loop_fused()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 50 * 1e3
print(f"graph(policy_act + step) us/step: {us:.1f}  -> {n/us/1e3:.2f} G env-steps/s")
