"""python profiles/tools/time_tc_logits.py <libmsort.so> — us per 1 048 576 envs of the tensor-core embedded policy alone
(tc_logits_kernel through msort_debug_policy_logits): isolates the MMA round trips from the rest of the step kernel."""
import os, sys
os.environ["MSORT_LIB"] = sys.argv[1]
sys.path.insert(0, "/root/repo")
import torch
import marl_sortingenv_b200 as ms
from marl_sortingenv_b200.policy import sb3_style_init
n = int(os.environ.get("N", 1 << 20))
env = ms.BatchedPressingEnv(256, max_steps=50, seed=1)
env.set_sort_policy(sb3_style_init(0, action_gain=1.0))
x = torch.rand((n, 13), device="cuda")
for _ in range(3):
    env.policy_logits_tensor(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    env.policy_logits_tensor(x)
e1.record(); torch.cuda.synchronize()
print(f"{os.path.basename(sys.argv[1]):28s} tc_logits n {n}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
if os.environ.get("TCPROF"):
    out = env.policy_logits_tensor(x)
    torch.cuda.synchronize()
    names = ["start", "A1 built", "sync1", "issued1", "mma1 done", "epilogue1 done", "(A2 built)", "sync2", "issued2", "mma2 done", "-", "epilogue2 done"]
    v = out.reshape(-1)[:12].tolist()
    print("   cycles since tile start (CTA 0, last tile): " + ", ".join(f"{n} {int(c)}" for n, c in zip(names, v)), flush=True)
