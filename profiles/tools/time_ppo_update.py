"""python profiles/tools/time_ppo_update.py — ms per MaskablePPO update (n_epochs x minibatches of the native kernels) and
per rollout for a few (num_envs, n_steps, batch_size) shapes; splits the update into its three kernels with CUDA events."""
import ctypes as C, os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import marl_sortingenv_b200 as ms
from marl_sortingenv_b200 import _abi
from marl_sortingenv_b200.ppo import MaskablePPO

def ev():
    return torch.cuda.Event(enable_timing=True)

for n, T, bs in ((2048, 64, 16384), (16384, 64, 65536), (131072, 32, 262144)):
    env = ms.BatchedMonolithEnv(n, max_steps=200, seed=1, noise_sorting=0.0, info_level="none", track_stats=False)
    m = MaskablePPO(env, n_steps=T, batch_size=bs, n_epochs=10)
    for _ in range(3):
        adv, ret = m.collect_rollout(); m.update(adv, ret)
    torch.cuda.synchronize()
    e = [ev() for _ in range(4)]
    e[0].record(); adv, ret = m.collect_rollout(); e[1].record(); m.update(adv, ret); e[2].record(); torch.cuda.synchronize()
    rows = n * T
    print(f"n {n} T {T} batch {bs}: rollout {e[0].elapsed_time(e[1]):.2f} ms ({rows / e[0].elapsed_time(e[1]) / 1e3:.1f} M steps/s), "
          f"update {e[1].elapsed_time(e[2]):.2f} ms = {10 * -(-rows // bs)} minibatches of {bs} rows "
          f"({e[1].elapsed_time(e[2]) * 1e3 / (10 * -(-rows // bs)):.1f} us each, {10 * rows / e[1].elapsed_time(e[2]) / 1e3:.1f} M rows/s)", flush=True)
    # one gradient kernel alone
    lib = m.lib
    batch = m._batch(rows, m.buf["obs"], m.buf["mask"], m.buf["act"], m.buf["logp"], adv, ret)
    perm = torch.randperm(rows, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    for _ in range(3):
        lib.msort_ppo_gradient(C.byref(batch), C.byref(m._hp), p(m.flat), p(m._grads), p(perm), 0, min(bs, rows), p(m._scratch), p(m._stats), st)
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(20):
        lib.msort_ppo_gradient(C.byref(batch), C.byref(m._hp), p(m.flat), p(m._grads), p(perm), 0, min(bs, rows), p(m._scratch), p(m._stats), st)
    b.record(); torch.cuda.synchronize()
    m._grads.zero_()
    print(f"      gradient kernel pair (adv stats + forward/backward) on {min(bs, rows)} rows: {a.elapsed_time(b) / 20 * 1e3:.1f} us", flush=True)
    env.close()
