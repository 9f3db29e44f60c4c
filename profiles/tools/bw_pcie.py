"""PCIe table: pinned host <-> device copy bandwidth per rank, alone and with all ranks copying at once.

    python profiles/tools/bw_pcie.py                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/tools/bw_pcie.py

Per rank: D2H, H2D and both directions at once for 16 / 128 MB buffers (CUDA events, best of 5), first one rank at a
time (the others idle), then all ranks together (barrier before every measurement, max over ranks = the job's rate).
Names the limiter of bench.py's 8-GPU e2e number: if the all-ranks D2H rate per GPU drops against the solo rate, the
host side (root ports / memory) — not the GPUs — is the ceiling."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
numa = None
if os.environ.get("BIND", "1") == "1":
    from marl_sortingenv_b200.sharding import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def rate(mb, mode, active):
    n = mb << 20
    h1, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d1, d2 = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = 0.0
    for _ in range(6):
        barrier()
        if not active:
            barrier()
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s1):
                h1.copy_(d1, non_blocking=True)
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s2):
                d2.copy_(h2, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, n * (2 if mode == "both" else 1) / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        barrier()
    return best


out = {"world": world, "numa": numa, "rows": []}
for mb in (16, 128):
    for mode in ("d2h", "h2d", "both"):
        solo = [0.0] * world
        for r in range(world):
            v = rate(mb, mode, rank == r)
            t = torch.tensor([v], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            solo[r] = float(t.item())
        v = rate(mb, mode, True)
        t = torch.tensor([v], device=dev)
        allr = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(allr, t)
        else:
            allr = [t]
        together = [float(x.item()) for x in allr]
        if rank == 0:
            out["rows"].append({"MB": mb, "mode": mode, "solo_GBs": [round(x, 1) for x in solo],
                                "together_GBs": [round(x, 1) for x in together], "together_sum_GBs": round(sum(together), 1)})
            print(f"{mb:4d} MB {mode:5s} solo {[round(x, 1) for x in solo]}  together {[round(x, 1) for x in together]}  sum {sum(together):.1f} GB/s", flush=True)
nodes = [None] * world
if world > 1:
    dist.all_gather_object(nodes, numa)
else:
    nodes = [numa]
if rank == 0:
    out["numa_per_rank"] = nodes
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
