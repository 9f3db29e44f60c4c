#!/usr/bin/env python
"""TEST INFRASTRUCTURE — CPU throughput of the UNMODIFIED reference under a SubprocVecEnv-style pool
(BASELINE.md §3): one worker process per host core, each owning a slice of reference envs, driven
over pipes with the command protocol of SB3's SubprocVecEnv (`step(actions)` → stacked obs / rewards /
dones / infos with `terminal_observation` and an unseeded `reset()` on done; `env_method("action_masks")`).
stable-baselines3 itself is not installed (and not installable offline), hence the restatement.

Needs a copy of the reference: /root/reference in the build container, or oracle/_ref, which oracle/make_ref.sh
stages from it and which travels to the GPU box with the snapshot (git-ignored).  bench.py's reference arm calls
run_steps(); the round-1 build-container result is profiles/cpu_reference_subproc_r01.json.

    python oracle/ref_subproc_bench.py [--seconds 6] [--envs-per-worker 8] [--out profiles/...json]
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(conn, kind, n_envs, seed0, max_steps, noise):
    from oracle.ref_loader import make_reference_env
    envs = [make_reference_env(kind, max_steps=max_steps, seed=seed0 + i, noise_sorting=noise, balesize=200)
            for i in range(n_envs)]
    while True:
        cmd, data = conn.recv()
        if cmd == "reset":
            conn.send(np.stack([e.reset(seed=seed0 + i)[0] for i, e in enumerate(envs)]))
        elif cmd == "masks":                                   # env_method("action_masks")
            conn.send(np.stack([np.asarray(e.action_masks(), dtype=bool) for e in envs]))
        elif cmd == "step":
            obs, rew, done, infos = [], [], [], []
            for e, a in zip(envs, data):
                o, r, term, trunc, info = e.step(int(a))
                if term or trunc:                              # VecEnv auto-reset semantics
                    info = dict(info, terminal_observation=o)
                    o, _ = e.reset()
                obs.append(o); rew.append(r); done.append(term); infos.append(info)
            conn.send((np.stack(obs), np.asarray(rew), np.asarray(done), infos))
        elif cmd == "close":
            conn.close()
            return


def run_steps(kind: str, steps: int, warmup: int, envs_per_worker: int, workers: int, max_steps=50, noise=0.05):
    """`warmup` untimed and `steps` timed batched steps (bench.py --impl reference); (env-steps/s, env-steps, seconds)."""
    return run(kind, None, envs_per_worker, workers, max_steps=max_steps, noise=noise, steps=steps, warmup=warmup)


def run(kind: str, seconds, envs_per_worker: int, workers: int, max_steps=50, noise=0.05, steps=None, warmup=5):
    ctx = mp.get_context("fork")
    pipes, procs = [], []
    for w in range(workers):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_worker, args=(b, kind, envs_per_worker, 1 + w * envs_per_worker, max_steps, noise), daemon=True)
        p.start(); b.close()
        pipes.append(a); procs.append(p)
    for c in pipes:
        c.send(("reset", None))
    for c in pipes:
        c.recv()
    rng = np.random.default_rng(0)

    def one_step():
        for c in pipes:
            c.send(("masks", None))
        masks = [c.recv() for c in pipes]
        for c, m in zip(pipes, masks):                         # uniform over the valid mask, as MaskablePPO's env_method + sample
            acts = [int(rng.choice(np.flatnonzero(row))) for row in m]
            c.send(("step", acts))
        for c in pipes:
            c.recv()
    for _ in range(warmup):
        one_step()
    n_steps, t0 = 0, time.perf_counter()
    while (n_steps < steps) if steps is not None else (time.perf_counter() - t0 < seconds):
        one_step()
        n_steps += 1
    dt = time.perf_counter() - t0
    for c in pipes:
        c.send(("close", None))
    for p in procs:
        p.join(timeout=5)
    total = n_steps * workers * envs_per_worker
    return total / dt, total, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=6.0)
    ap.add_argument("--envs-per-worker", type=int, default=8)
    ap.add_argument("--workers", type=int, default=os.cpu_count())
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = {"what": "unmodified reference envs under a SubprocVecEnv-style pool (pipes, one process per core, VecEnv "
                   "auto-reset, env_method('action_masks') before every step, uniform valid actions)",
           "host": platform.processor() or platform.machine(), "cores": a.workers, "numpy": np.__version__,
           "python": platform.python_version(), "envs_per_worker": a.envs_per_worker, "max_steps": 50,
           "noise_sorting": 0.05, "results": {}}
    for kind in ("sort", "press", "mono"):
        sps, total, dt = run(kind, a.seconds, a.envs_per_worker, a.workers)
        res["results"][kind] = {"env_steps_per_sec": sps, "env_steps": total, "seconds": dt}
        print(f"{kind}: {sps:,.0f} env-steps/s ({total} steps in {dt:.1f} s, {a.workers} processes)", flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
