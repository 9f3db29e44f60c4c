"""N>1 host logic on CPU: contiguous sharding by global env id, the episode-statistics
all-reduce over gloo (world_size 2), and invariance of trajectories to the shard count
(Philox counters use the GLOBAL env id) — checked with the CPU oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marl_sortingenv_b200.sharding import allreduce_stats, shard_bounds
from parity_util import config_for, state_rows

META = dict(kind="mono", max_steps=20, noise=0.05, balesize=200, use_action_masking=True,
            check_overflow=False, auto_reset=True)
GLOBAL_N, T = 301, 45


def test_shard_bounds_cover_and_are_contiguous():
    for n, w in ((10, 4), (8, 8), (7, 3), (1 << 23, 8)):
        b = [shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _run_shard(lo, hi):
    from oracle.cpu_oracle import OracleEnv
    env = OracleEnv(config_for(META, hi - lo, rng_mode="philox", seed=11, global_env_offset=lo))
    env.reset()
    rewards = []
    for t in range(T):
        a = env.sample_masked_actions(5, t)
        _, r, *_ = env.step(a)
        rewards.append(r.copy())
    return state_rows(env.state), np.stack(rewards, 0), env.stats.copy()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(GLOBAL_N, rank, world)
    st, rew, stats = _run_shard(lo, hi)
    total = allreduce_stats(torch.from_numpy(stats.copy()))
    q.put((rank, lo, hi, st, rew, total.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    st1, rew1, stats1 = _run_shard(0, GLOBAL_N)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    st2 = np.concatenate([o[3] for o in out], 0)
    rew2 = np.concatenate([o[4] for o in out], 1)
    assert np.array_equal(st1, st2), "trajectories depend on the shard count"
    assert np.array_equal(rew1, rew2)
    for o in out:                                    # both ranks hold the all-reduced statistics
        assert np.allclose(o[5], stats1, rtol=1e-12, atol=1e-9)
    assert stats1[0] == GLOBAL_N * (T // META["max_steps"]) and stats1[3] == GLOBAL_N * T
