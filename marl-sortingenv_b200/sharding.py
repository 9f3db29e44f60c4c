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
