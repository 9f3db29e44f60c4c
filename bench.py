#!/usr/bin/env python
"""bench.py — env-steps/s of the fused step() hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]           # product arm (CUDA)
    python bench.py --impl reference [--steps K] [--warmup W]     # CPU arm (oracle port)

Workload (config.workload): Env_3_Monolith with action masking, 1 048 576 envs per GPU
(BASELINE configs[3]; configs[4] = the same per-GPU shard on 2/4/8 GPUs, weak scaling),
PHILOX generator, auto-reset, max_steps=50.  A "step" is ONE fused step() launch over the
whole batch with valid masked-random actions already resident in HBM.  The actions are
pre-recorded by an identical seeded run (the dynamics are deterministic per seed), so the
timed region contains only the hot path.  Working set per step (state 218 MB + obs 122 MB +
mask 23 MB + actions/reward/done) exceeds the 126 MB L2, so no L2 flush is needed between
iterations ("inputs larger than L2").

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import socket
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ENVS_PER_GPU = 1 << 20
MAX_STEPS = 50
SEED = 42
ACTION_SEED = 7
ALGO_BYTES_PER_STEP = {"sort": 223, "press": 244, "mono": 307}   # SURVEY.md §8d / DESIGN.md
# dram__bytes_read.sum + dram__bytes_write.sum per step-kernel launch at this workload, from the
# `ncu --set full` capture summarised in profiles/ (None until a capture exists).
NCU_TRAFFIC_BYTES_PER_LAUNCH = {("mono", 1 << 20): 232.4e6}   # profiles/ncu_r01_step_kernel.md (kernel v17: 75.5 MB read + 156.8 MB written)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="msort", choices=["msort", "reference"])
    ap.add_argument("--step-streams", type=int, default=2, help="env ranges / CUDA streams each timed step is launched as (1 = one whole-batch launch)")
    ap.add_argument("--rollout-streams", type=int, default=2, help="env ranges / CUDA streams of the rollout-loop measurement")
    ap.add_argument("--kind", default="mono", choices=["sort", "press", "mono"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-rollout", action="store_true", help="skip the policy-in-the-loop rollout measurement")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch the K timed steps one by one instead of replaying one CUDA graph of them")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clock/throttle sampling DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for (ts, l) in self.lines if t0 - 0.05 <= ts <= t1 + 0.15] or [l for _, l in self.lines]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
                for nme, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_rollout(kind: str, budget_s: float, steps: int | None = None, warmup: int = 0):
    """Times the CPU oracle (C port of the reference algorithm, oracle/msort_oracle.c) on all host
    cores on a bounded sample of the same workload.  Returns (env_steps_per_s, cores, sample, K)."""
    from oracle.cpu_oracle import OracleEnv
    from parity_util import config_for
    cores = os.cpu_count() or 1
    meta = dict(kind=kind, max_steps=MAX_STEPS, noise=0.05, balesize=200, use_action_masking=True,
                check_overflow=False, auto_reset=True)
    n_cal = 4096 * cores
    env = OracleEnv(config_for(meta, n_cal, rng_mode="philox", seed=SEED), nthreads=cores)
    env.reset()
    t = time.perf_counter(); env.rollout(5, ACTION_SEED); per_env_step = (time.perf_counter() - t) / (5 * n_cal)
    if steps is None:
        n, K = n_cal * 4, None
        K = max(10, int(budget_s / (per_env_step * n)))
    else:
        K = steps
        n = int(budget_s / (per_env_step * max(1, K + warmup)))
        n = max(cores * 256, min(ENVS_PER_GPU, n // (cores * 64) * (cores * 64)))
    env = OracleEnv(config_for(meta, n, rng_mode="philox", seed=SEED), nthreads=cores)
    env.reset()
    if warmup:
        env.rollout(warmup, ACTION_SEED)
    t = time.perf_counter()
    done, _ = env.rollout(K, ACTION_SEED + 1)
    dt = time.perf_counter() - t
    return done / dt, cores, f"{n} envs x {K} steps ({done} env-steps, {dt:.1f} s)", K, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, cores, sample, K, dt = cpu_rollout(args.kind, budget_s=25.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, K),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": ("CPU arm = oracle/msort_oracle.c (C port of the reference's step/reset, pinned to recorded "
                 "reference trajectories) on all host cores; the Python reference itself cannot travel to the box "
                 "(measured in the build container under a SubprocVecEnv-style pool: 18.9e3 env-steps/s on 8 cores, "
                 "profiles/cpu_reference_subproc_r01.json)"),
    }
    emit(line)


def workload_config(args, world):
    kind_name = {"sort": "Env_1_Sorting", "press": "Env_2_Pressing", "mono": "Env_3_Monolith"}[args.kind]
    return {"workload": f"{kind_name} with action masking, {args.envs_per_gpu} envs per GPU "
                        f"({args.envs_per_gpu * world} total), PHILOX generator, auto-reset, max_steps={MAX_STEPS}, "
                        f"masked-random actions pre-recorded in HBM",
            "env": kind_name, "envs_per_gpu": args.envs_per_gpu, "global_envs": args.envs_per_gpu * world,
            "max_steps": MAX_STEPS, "rng": "philox4x32-10", "action_masking": True, "auto_reset": True,
            "l2": "inputs larger than L2 (no flush needed)", "parallelism": f"env-sharded x{world}",
            "launch": ("eager, one whole-batch step kernel per step" if args.no_graph else
                       "one CUDA graph of the K steps" + (", each step = 2 env-range launches of the step kernel on 2 streams"
                                                          if args.step_streams >= 2 else ", one step kernel per step")),
            "stats_allreduce": "once per K-step rollout (NCCL, 128 B)"}


# ----------------------------------------------------------------------------- the one JSON line
_REAL_STDOUT = None


def claim_stdout():
    """Native libraries print banners on fd 1 (NCCL's "NCCL version ..." at communicator creation): point fd 1
    at stderr for the rest of the process and keep the real stdout for the ONE JSON line the driver parses."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ----------------------------------------------------------------------------- CUDA arm
def run_msort(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:           # convenience: re-launch ourselves under torchrun
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    claim_stdout()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from marl_sortingenv_b200.batched import ENV_CLASSES
    from marl_sortingenv_b200.sharding import allreduce_stats, shard_offset

    n, K, W = args.envs_per_gpu, args.steps, args.warmup
    W = max(W, 3)
    env = ENV_CLASSES[args.kind](n, device=dev, max_steps=MAX_STEPS, seed=SEED, info_level="episode",
                                 global_env_offset=shard_offset(n * world, rank, world))
    if args.kind == "press":
        from marl_sortingenv_b200.policy import sb3_style_init
        env.set_sort_policy(sb3_style_init(0))

    # ---- record valid masked-random actions with an identical seeded run (not timed)
    T = W + K
    actions = torch.empty((T, n), dtype=torch.int64, device=dev)
    env.reset(seed=SEED)
    for t in range(T):
        env.sample_actions(ACTION_SEED, t, out=actions[t])
        env.step(actions[t])
    env.sync_check()
    final_ref = env.state.clone()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- timed region: K fused step() launches back to back (one CUDA graph of K kernel nodes by
    #      default, so the measurement is not throttled by Python launch overhead)
    use_graph = not args.no_graph
    env.reset(seed=SEED)
    env.stats.zero_()
    for t in range(W):
        env.step(actions[t])
    if world > 1:                                  # NCCL communicators are created lazily: do it before timing
        allreduce_stats(torch.zeros(16, dtype=torch.float64, device=dev))
    # Each step is launched as `step_streams` env ranges of the one handle on as many CUDA streams (msort_step_range;
    # ranges are independent, the result is bit-identical — asserted below): the tail of one range's kernel overlaps
    # the head of the other's, which at 1 M envs recovers the ramp / tail a single 8192-CTA launch pays (62 -> 58 us).
    # --no-graph keeps one whole-batch launch per step (per-launch events, the ncu launch list).
    half = (n // 2 + 127) // 128 * 128
    ranges = [(0, n)] if (args.step_streams < 2 or half >= n or not use_graph) else [(0, half), (half, n)]
    streams = [torch.cuda.Stream(device=dev) for _ in ranges] if len(ranges) > 1 else []
    graph = None
    if use_graph:
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            if not streams:
                for t in range(K):
                    env.step(actions[W + t])
            else:
                cur = torch.cuda.current_stream(dev)
                for st_ in streams:
                    st_.wait_stream(cur)
                for t in range(K):
                    for st_, r in zip(streams, ranges):
                        with torch.cuda.stream(st_):
                            env.step(actions[W + t], env_range=r)
                for st_ in streams:
                    cur.wait_stream(st_)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)] if not use_graph else None
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.25)
    barrier()
    t0 = time.time()
    ev0.record()
    if use_graph:
        graph.replay()
    else:
        ev[0].record()
        for t in range(K):
            env.step(actions[W + t])
            ev[t + 1].record()
    if world > 1:
        allreduce_stats(env.stats)            # the only collective: 128 B of episode stats per rollout
    ev1.record()
    barrier()
    t1 = time.time()
    gpu_launches = K * len(ranges)
    clocks = sampler.stop(t0, t1) if sampler else None
    total_ms = ev0.elapsed_time(ev1)
    per_kernel_ms = [ev[t].elapsed_time(ev[t + 1]) for t in range(K)] if not use_graph else [total_ms / K]
    assert torch.equal(env.state, final_ref), "timed run diverged from the recorded run (non-determinism)"
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    value = n * world * K / (total_ms * 1e-3)
    stats = env.stats.clone()
    stats = stats.cpu().tolist()

    # ---- e2e: the same steps through the host-buffer API (pinned H2D of actions, D2H of results)
    e2e = None
    if not args.no_e2e:
        Ke = min(K, 50)
        host_actions = actions[W:W + Ke].cpu().pin_memory()
        env.reset(seed=SEED)
        for t in range(W):
            env.step(actions[t])
        env.step_host(host_actions[0]); env.reset(seed=SEED)      # allocate pinned buffers outside the timing
        for t in range(W):
            env.step(actions[t])
        barrier()
        te0 = time.perf_counter()
        for t in range(Ke):
            env.step_host(host_actions[t])
        torch.cuda.synchronize(dev)
        te = torch.tensor([time.perf_counter() - te0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": n * world * Ke / float(te.item()), "unit": "env-steps/s",
               "h2d_bytes_per_step": env.h2d_bytes * world, "d2h_bytes_per_step": env.d2h_bytes * world,
               "steps": Ke, "api": "BatchedEnv.step_host (pinned host actions in; obs, reward, terminated, mask out)"}

    # ---- rollout loop (BASELINE configs[3] "MaskablePPO rollout loop"): every env-step = fused actor-critic
    #      inference + masked categorical draw (msort_policy_act, tcgen05) followed by the fused step(),
    #      fresh SB3-style random-init towers 29->32->32->{22|1}; one CUDA graph of Kr such pairs
    rollout = None
    if not args.no_rollout:
        from marl_sortingenv_b200.ppo import MaskableActorCritic, pack_actor_critic
        torch.manual_seed(0)
        packed = pack_actor_critic(MaskableActorCritic(env.D, env.A).to(dev))
        out = (torch.empty(n, dtype=torch.int64, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev))
        Kr = 64
        env.reset(seed=SEED)
        for t in range(W):
            env.policy_act(packed, seed=ACTION_SEED, t=t, out=out); env.step(out[0])
        torch.cuda.synchronize(dev)
        # two env ranges of the one handle on two streams (msort_*_range), as ppo.MaskablePPO's rollout does: the
        # latency-bound policy kernel of one range shares the SMs with the step kernel of the other
        half = (n // 2 + 127) // 128 * 128
        ranges = [(0, n)] if args.rollout_streams < 2 or half >= n else [(0, half), (half, n)]
        streams = [torch.cuda.Stream(device=dev) for _ in ranges]
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            cur = torch.cuda.current_stream(dev)
            for s in streams:
                s.wait_stream(cur)
            for t in range(Kr):
                for s, r in zip(streams, ranges):
                    with torch.cuda.stream(s):
                        env.policy_act(packed, seed=ACTION_SEED, t=W + t, out=out, env_range=r); env.step(out[0], env_range=r)
            for s in streams:
                cur.wait_stream(s)
        gr.replay()
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(); gr.replay(); r1.record()
        barrier()
        tr = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        rollout = {"value": n * world * Kr / (float(tr.item()) * 1e-3), "unit": "env-steps/s", "steps": Kr,
                   "ms_per_step": float(tr.item()) / Kr, "launches_per_step": 2 * len(ranges), "streams": len(ranges),
                   "what": "per env-step: msort_policy_act (actor-critic 29-32-32-{22|1} on tcgen05, masked categorical "
                           "draw) + fused step(), env ranges on separate streams; obs/mask never leave HBM"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    kern_ms = statistics.mean(per_kernel_ms)
    algo_bytes = ALGO_BYTES_PER_STEP[args.kind] * n
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    line = {
        "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH.get((args.kind, n)), "kernel": f"msort::step_kernel<{args.kind.upper()},PHILOX> [{env.step_variant} instantiation]",
                     "kernel_ms": kern_ms, "launches_per_step": len(ranges), "algorithmic_bytes_per_launch": algo_bytes,
                     "bytes_per_env_step": ALGO_BYTES_PER_STEP[args.kind], "peak_source": peak_src},
        "episode_stats": {"episodes": stats[0], "mean_return": stats[1] / max(1.0, stats[0]),
                          "mean_length": stats[2] / max(1.0, stats[0]), "env_steps": stats[3],
                          "bales": stats[6]},
    }
    if rollout is not None:
        line["rollout"] = rollout
    if not args.no_cpu_baseline:
        v, cores, sample, _, _ = cpu_rollout(args.kind, budget_s=15.0)
        line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_msort(args)


if __name__ == "__main__":
    main()
