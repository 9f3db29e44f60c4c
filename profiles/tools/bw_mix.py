"""HBM bandwidth of plain torch kernels at different read:write mixes (python profiles/tools/bw_mix.py).
The step kernel moves 72 B in and 207 B out per env-step (26 % reads / 74 % writes); MEASURED_PEAKS.json's
hbm_gbs is a 50/50 copy.  This prints what the same GPU sustains for pure-write, copy, read-heavy and pure-read
streams (measured round 1: fill 7.2, copy 6.4, add 6.9, sum 5.7 TB/s), to say that a write-heavy mix is not what
keeps the kernel's real DRAM traffic (4.6 TB/s) below the copy figure."""
import torch

dev = "cuda"
n = 1 << 28                                   # 256 Mi floats = 1 GiB


def timeit(fn, nbytes, label, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{label:44s} {nbytes / best / 1e6:8.1f} GB/s  ({best:.3f} ms)", flush=True)


a = torch.ones(n, device=dev)
b = torch.ones(n, device=dev)
c = torch.empty(n, device=dev)
timeit(lambda: c.fill_(1.0), 4 * n, "fill_            (0 % read / 100 % write)")
timeit(lambda: c.copy_(a), 4 * n * 2, "copy_            (50 % read / 50 % write)")
timeit(lambda: torch.add(a, b, out=c), 4 * n * 3, "add              (67 % read / 33 % write)")
timeit(lambda: a.sum(), 4 * n, "sum              (100 % read)")
