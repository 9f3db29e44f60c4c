"""python profiles/tools/time_e2e.py [kind] — ms per step of BatchedEnv.step_host (msort_step_host) for several chunk
counts, 1 048 576 envs, pinned uint8 actions: finds the chunking that keeps the D2H engine busiest."""
import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import marl_sortingenv_b200 as ms
kind = sys.argv[1] if len(sys.argv) > 1 else "mono"
n = int(os.environ.get("N", 1 << 20))
cls = {"mono": ms.BatchedMonolithEnv, "sort": ms.BatchedSortingEnv, "press": ms.BatchedPressingEnv}[kind]
env = cls(n, max_steps=50, seed=42, info_level="episode")
env.reset()
T = 24
acts = torch.zeros((T, n), dtype=torch.uint8).pin_memory()
tmp = torch.zeros(n, dtype=torch.int64, device="cuda")
for t in range(T):
    env.sample_actions(7, t, out=tmp); env.step(tmp); acts[t].copy_(tmp.to(torch.uint8).cpu())
bytes_per_step = None
for chunks in (1, 2, 3, 4, 6, 8, 12, 16, 32):
    env.reset(seed=42)
    env.step_host(acts[0], chunks=chunks); env.reset(seed=42)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(T):
        env.step_host(acts[t], chunks=chunks)
    dt = (time.perf_counter() - t0) / T
    b = env.h2d_bytes + env.d2h_bytes
    print(f"{kind} chunks {chunks:3d}: {dt * 1e3:7.3f} ms/step  {n / dt / 1e9:.3f} G env-steps/s  {b / dt / 1e9:.1f} GB/s over PCIe", flush=True)
