"""Multi-GPU sharding: independent envs are split into contiguous blocks of the GLOBAL env
index, one process per GPU.  There is no exchange step in the dynamics, so the only collective
is one all-reduce (SUM) of the 16-element f64 episode-statistics vector per rollout
(NCCL over NVLink on GPUs; gloo in the CPU tests).  Philox counters use the global env id, so a
trajectory does not depend on how many GPUs the batch is sharded over."""
from __future__ import annotations

import os

import torch


def shard_bounds(global_envs: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the global env ids owned by `rank`: contiguous, sizes differ by at most 1."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(int(global_envs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_offset(global_envs: int, rank: int, world: int) -> int:
    return shard_bounds(global_envs, rank, world)[0]


def env_rank_world() -> tuple[int, int, int]:
    """(rank, local_rank, world) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def allreduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """SUM-all-reduce of the episode statistics vector in place (no-op without a process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def make_sharded_env(kind: str, global_envs: int, **kwargs):
    """This rank's shard of a `global_envs`-wide batch (one process per GPU)."""
    from .batched import ENV_CLASSES
    rank, local_rank, world = env_rank_world()
    lo, hi = shard_bounds(global_envs, rank, world)
    kwargs.setdefault("device", f"cuda:{local_rank}")
    return ENV_CLASSES[kind](hi - lo, global_env_offset=lo, **kwargs)


class StatsReducer:
    """The one collective of the path — SUM-all-reduce of the 16 x f64 episode-statistics vector per rollout
     — taken OFF the stepping stream: `start(stats)` snapshots the vector into one of two
    buffers and issues the all-reduce on a side stream ordered after everything queued so far; the stepping stream
    goes on with the next rollout at once.  `result()` joins (the consumer's stream waits for the side stream) and
    returns the reduced vector of the rollout before.  Without a process group the snapshot is the result."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.side = torch.cuda.Stream(device=self.device)
        self.buf = [torch.zeros(16, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.k = 0
        self.done = torch.cuda.Event()
        self.pending = None

    def start(self, stats: torch.Tensor) -> None:
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        b = self.buf[self.k]
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            b.copy_(stats)
            allreduce_stats(b)
            self.done.record(self.side)
        self.pending, self.k = b, self.k ^ 1

    def result(self) -> torch.Tensor:
        if self.pending is None:
            raise RuntimeError("StatsReducer.result() before start()")
        torch.cuda.current_stream(self.device).wait_event(self.done)
        return self.pending


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process to the CPU cores of the NUMA node the GPU hangs off and prefer that node's memory, so the
    pinned host buffers of `step_host` are allocated next to the GPU's PCIe root port (with 8 ranks on a two-socket
    box, half of them otherwise stream 50 GB/s across the socket link).  Best effort: returns what it found and did."""
    import ctypes
    out = {"numa_node": None, "cpus": None, "mempolicy": False}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        out["pci"] = bdf
        if node < 0:
            return out
        out["numa_node"] = node
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            out["cpus"] = len(allowed)
        mask = ctypes.c_ulong(1 << node)
        libc = ctypes.CDLL(None, use_errno=True)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238                       # x86_64
        if libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64)) == 0:
            out["mempolicy"] = True
    except Exception as e:                                                # containers may hide sysfs / forbid the syscall
        out["error"] = f"{type(e).__name__}: {e}"
    return out
