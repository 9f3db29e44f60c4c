#!/usr/bin/env python
"""bench.py — env-steps/s of the fused step() hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]           # product arm (CUDA)
    python bench.py --impl reference [--steps K] [--warmup W]     # CPU arm: the unmodified reference (+ the C port)

Headline workload (config.workload): Env_3_Monolith with action masking, 1 048 576 envs per GPU
(BASELINE configs[3]; configs[4] = the same per-GPU shard on 2/4/8 GPUs, weak scaling),
PHILOX generator, auto-reset, max_steps=50.  A "step" is ONE fused step() over the whole batch with
valid masked-random actions already resident in HBM.  The actions are pre-recorded by an identical
seeded run (the dynamics are deterministic per seed), so the timed region contains only the hot path.
Working set per step (state 218 MB + obs 122 MB + mask 23 MB + actions/reward/done) exceeds the
126 MB L2, so no L2 flush is needed between iterations ("inputs larger than L2").

The line also carries, each measured in the same run:
  e2e        the same steps through the host-buffer call (msort_step_host: pinned host actions in, obs / reward /
             flag words out, chunk-pipelined) — the headline against the reference arm
  rollout    policy inference + masked draw + step per env-step (BASELINE configs[3] "rollout loop")
  configs    the other BASELINE configurations: Env_1 65 536 envs (PHILOX and REPLAY), Env_2 262 144 envs with the
             embedded policy, Env_3 8 388 608 envs in total sharded over the ranks (strong scaling)
  shard_invariance   the same 65 536 global env ids stepped on 1 GPU and on N GPUs end in the same state
  allreduce_us       the one collective (16 x f64 episode statistics), issued on a side stream

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import re
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
STRONG_GLOBAL_ENVS = 1 << 23          # BASELINE configs[4]: 8 388 608 envs in total
MAX_STEPS = 50
SEED = 42
ACTION_SEED = 7
ALGO_BYTES_PER_STEP = {"sort": 223, "press": 244, "mono": 307}   # SURVEY.md §8d / DESIGN.md
REPLAY_EXTRA_BYTES = 132                                         # REPLAY mode: consumed stream bytes per env-step
KIND_NAME = {"sort": "Env_1_Sorting", "press": "Env_2_Pressing", "mono": "Env_3_Monolith"}


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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --global-envs in total, split over the ranks (BASELINE configs[4] as the survey sized it)")
    ap.add_argument("--global-envs", type=int, default=STRONG_GLOBAL_ENVS)
    ap.add_argument("--e2e-chunks", type=int, default=0, help="env ranges msort_step_host pipelines (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-rollout", action="store_true", help="skip the policy-in-the-loop rollout measurement")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
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


def ncu_traffic(kind: str, n: int):
    """DRAM bytes per step-kernel launch from the committed `ncu --set full` summary of this workload (None if there is
    none): the newest profiles/ncu_r*_step_kernel*.md whose kernel line names this env kind and whose grid matches."""
    want_grid = (n + 127) // 128
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(pdir)):
        if not re.match(r"ncu_r\d+_step_kernel.*\.md$", f):
            continue
        txt = open(os.path.join(pdir, f)).read()
        k = re.search(r"\| kernel \| \| (.*?) \|", txt)
        g = re.search(r"\| launch__grid_size \|[^|]*\| *([\d.]+)", txt)
        t = re.search(r"DRAM traffic = [\d.]+ \+ [\d.]+ = ([\d.]+) (\w+)", txt)
        if not (k and g and t):
            continue
        num = {"sort": 1, "press": 2, "mono": 3}[kind]                     # first template argument = the env kind
        if not re.search(rf"step_kernel<(\(int\))?{num},", k.group(1)):
            continue
        if int(float(g.group(1))) != want_grid:
            continue
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(t.group(2), 1e6)
        best = {"bytes": float(t.group(1)) * scale, "source": f"profiles/{f}", "kernel": k.group(1)[:96]}
    return best


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


# ----------------------------------------------------------------------------- CPU arms
def cpu_rollout(kind: str, budget_s: float, steps: int | None = None, warmup: int = 0):
    """Times the CPU oracle (C port of the reference algorithm, oracle/msort_oracle.c) on all host
    cores on a bounded sample of the same workload.  Returns (env_steps_per_s, cores, sample, K, dt)."""
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


def python_reference(kind: str, steps: int, warmup: int, envs_per_worker: int = 64):
    """The UNMODIFIED reference envs (oracle/_ref, staged by oracle/make_ref.sh; /root/reference in the build
    container) under a SubprocVecEnv-style pool, one worker process per host core (oracle/ref_subproc_bench.py):
    `warmup` + `steps` batched steps with env_method('action_masks') + masked-random actions + VecEnv auto-reset.
    Returns None where no copy of the reference exists."""
    from oracle.ref_loader import REFERENCE_ROOT, reference_available
    if not reference_available():
        return None
    from oracle.ref_subproc_bench import run_steps
    cores = os.cpu_count() or 1
    sps, total, dt = run_steps(kind, steps=steps, warmup=warmup, envs_per_worker=envs_per_worker, workers=cores,
                               max_steps=MAX_STEPS, noise=0.05)
    return {"value": sps, "unit": "env-steps/s", "cores": cores, "kind": "reference",
            "sample": f"{cores} processes x {envs_per_worker} envs x {steps} steps ({total} env-steps, {dt:.1f} s)",
            "seconds": dt, "steps": steps, "root": os.path.relpath(REFERENCE_ROOT, ROOT) if REFERENCE_ROOT.startswith(ROOT) else REFERENCE_ROOT}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pv, cores, psample, K, pdt = cpu_rollout(args.kind, budget_s=20.0, steps=args.steps, warmup=args.warmup)
    port = {"value": pv, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": psample}
    ref = python_reference(args.kind, steps=max(args.steps, 20), warmup=max(args.warmup, 3))
    base = ref if ref is not None else port
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": base["value"], "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (ref["seconds"] / ref["steps"] if ref is not None else pdt / max(1, K)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "cpu_port": port,
        "e2e": {"value": base["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": ("value = the UNMODIFIED Python reference (Env_3_Monolith.step / reset / action_masks from oracle/_ref) under a "
                 "SubprocVecEnv-style pool, one process per host core, masked-random actions, VecEnv auto-reset; cpu_port = "
                 "oracle/msort_oracle.c (the C restatement pinned to recorded reference trajectories) on all host cores, the "
                 "conservative comparison" if ref is not None else
                 "no copy of the reference on this box (oracle/_ref missing): value = oracle/msort_oracle.c, the C port"),
    }
    emit(line)


def workload_config(args, world):
    kind_name = KIND_NAME[args.kind]
    per = envs_per_rank(args, world)
    return {"workload": f"{kind_name} with action masking, {per} envs per GPU "
                        f"({per * world} total), PHILOX generator, auto-reset, max_steps={MAX_STEPS}, "
                        f"masked-random actions pre-recorded in HBM",
            "env": kind_name, "envs_per_gpu": per, "global_envs": per * world,
            "max_steps": MAX_STEPS, "rng": "philox4x32-10", "action_masking": True, "auto_reset": True,
            "l2": "inputs larger than L2 (no flush needed)", "parallelism": f"env-sharded x{world}",
            "launch": ("eager, one whole-batch step kernel per step" if args.no_graph else
                       "one CUDA graph of the K steps" + (", each step = 2 env-range launches of the step kernel on 2 streams"
                                                          if args.step_streams >= 2 else ", one step kernel per step")),
            "stats_allreduce": "once per K-step rollout (NCCL, 128 B), on a side stream: off the stepping stream's critical path"}


def envs_per_rank(args, world):
    if args.scaling == "strong":
        return max(128, args.global_envs // world // 128 * 128)
    return args.envs_per_gpu


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


# ----------------------------------------------------------------------------- CUDA arm: building blocks
class Ctx:
    """Per-process CUDA / torch.distributed context of the product arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa = None
        if not args.no_numa:
            from marl_sortingenv_b200.sharding import bind_to_gpu_numa_node
            self.numa = bind_to_gpu_numa_node(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def make_env(ctx, kind, n, global_envs=None, **kw):
    from marl_sortingenv_b200.batched import ENV_CLASSES
    from marl_sortingenv_b200.sharding import shard_offset
    total = n * ctx.world if global_envs is None else global_envs
    env = ENV_CLASSES[kind](n, device=ctx.dev, max_steps=MAX_STEPS, seed=SEED, info_level="episode",
                            global_env_offset=shard_offset(total, ctx.rank, ctx.world), **kw)
    if kind == "press":
        from marl_sortingenv_b200.policy import sb3_style_init
        env.set_sort_policy(sb3_style_init(0))
    return env


def record_actions(env, T):
    """Valid masked-random actions of an identical seeded run (not timed); returns (actions [T, n], final state)."""
    torch = env.state.new_empty(0).__class__ and __import__("torch")
    actions = torch.empty((T, env.num_envs), dtype=torch.int64, device=env.device)
    env.reset(seed=SEED)
    for t in range(T):
        env.sample_actions(ACTION_SEED, t, out=actions[t])
        env.step(actions[t])
    env.sync_check()
    return actions, env.state.clone()


def timed_steps(ctx, env, actions, W, K, step_streams=2, use_graph=True, after=None):
    """W warm-up steps, then K fused step()s timed with CUDA events on the launch stream (one CUDA graph of the K steps,
    each step as `step_streams` env ranges on as many streams, unless `use_graph` is off).  `after()` runs on the stream
    right after the closing event (the side-stream statistics all-reduce).  Returns (total ms max over ranks, per-step ms
    list, launches per step)."""
    torch = ctx.torch
    n, dev = env.num_envs, ctx.dev
    env.reset(seed=SEED)
    if env.stats is not None:
        env.stats.zero_()
    for t in range(W):
        env.step(actions[t])
    half = (n // 2 + 127) // 128 * 128
    ranges = [(0, n)] if (step_streams < 2 or half >= n or not use_graph) else [(0, half), (half, n)]
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
    ctx.barrier()
    t0 = time.time()
    ev0.record()
    if use_graph:
        graph.replay()
    else:
        ev[0].record()
        for t in range(K):
            env.step(actions[W + t])
            ev[t + 1].record()
    ev1.record()
    if after is not None:
        after()
    ctx.barrier()
    t1 = time.time()
    total_ms = ctx.max_over_ranks(ev0.elapsed_time(ev1))
    per = [ev[t].elapsed_time(ev[t + 1]) for t in range(K)] if not use_graph else [total_ms / K]
    return total_ms, per, len(ranges), (t0, t1)


def roofline(kind, n, kern_ms, peak, peak_src, variant, launches_per_step, extra_bytes=0):
    algo = (ALGO_BYTES_PER_STEP[kind] + extra_bytes) * n
    achieved = algo / (kern_ms * 1e-3) / 1e9
    tr = ncu_traffic(kind, n)
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": tr["bytes"] if tr else None, "traffic_source": tr["source"] if tr else None,
            "traffic_note": ("dram__bytes_read.sum + dram__bytes_write.sum of ONE isolated whole-batch launch of this kernel (ncu --set full, "
                             "committed summary); the timed step issues the same work as `launches_per_step` env-range launches") if tr else None,
            "kernel": f"msort::step_kernel<{kind.upper()},PHILOX> [{variant} instantiation]",
            "kernel_ms": kern_ms, "launches_per_step": launches_per_step, "algorithmic_bytes_per_launch": algo,
            "bytes_per_env_step": ALGO_BYTES_PER_STEP[kind] + extra_bytes, "peak_source": peak_src}


def measure_e2e(ctx, args, env, actions, W, K):
    """The same steps through the host-buffer call: every step copies its actions from pinned host memory (1 byte per
    env) and brings obs / reward / flag words back to pinned host memory (msort_step_host, chunk-pipelined)."""
    import numpy as np
    torch = ctx.torch
    Ke = min(K, 50)
    host_actions = actions[W:W + Ke].to(torch.uint8).cpu().pin_memory()
    env.reset(seed=SEED)
    env.step_host(host_actions[0], chunks=args.e2e_chunks)          # allocate the pinned buffers outside the timing
    env.reset(seed=SEED)
    for t in range(W):
        env.step(actions[t])
    ctx.barrier()
    te0 = time.perf_counter()
    for t in range(Ke):
        out = env.step_host(host_actions[t], chunks=args.e2e_chunks)
    torch.cuda.synchronize(ctx.dev)
    dt = ctx.max_over_ranks(time.perf_counter() - te0)
    # the host copies are the device results of the last step, and the trajectory is the recorded one
    assert np.array_equal(out[0], env.obs.cpu().numpy()) and np.array_equal(np.asarray(out[4]), env.mask.cpu().numpy())
    assert np.array_equal(np.asarray(out[2]), env.terminated.cpu().numpy())
    n, world = env.num_envs, ctx.world
    return {"value": n * world * Ke / dt, "unit": "env-steps/s", "h2d_bytes_per_step": env.h2d_bytes * world,
            "d2h_bytes_per_step": env.d2h_bytes * world, "steps": Ke, "ms_per_step": 1e3 * dt / Ke,
            "pcie_gbs_per_gpu": (env.h2d_bytes + env.d2h_bytes) * Ke / dt / 1e9,
            "api": "BatchedEnv.step_host -> msort_step_host: pinned host uint8 actions in; obs f32, reward f32 and one 16-bit flag "
                   "word per env (11 mask bits + done) out; env ranges pipelined on the library's streams (H2D | kernel | D2H overlap)"}


def measure_rollout(ctx, args, env, W):
    """BASELINE configs[3] "MaskablePPO rollout loop": every env-step = fused actor-critic inference + masked categorical
    draw (msort_policy_act, tcgen05) followed by the fused step(); fresh SB3-style random-init towers; one CUDA graph."""
    torch = ctx.torch
    from marl_sortingenv_b200.ppo import MaskableActorCritic, pack_actor_critic
    n, dev, world = env.num_envs, ctx.dev, ctx.world
    torch.manual_seed(0)
    packed = pack_actor_critic(MaskableActorCritic(env.D, env.A).to(dev))
    out = (torch.empty(n, dtype=torch.int64, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev))
    Kr = 64
    env.reset(seed=SEED)
    for t in range(W):
        env.policy_act(packed, seed=ACTION_SEED, t=t, out=out); env.step(out[0])
    torch.cuda.synchronize(dev)
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
    ctx.barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(); gr.replay(); r1.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(r0.elapsed_time(r1))
    two = {"value": n * world * Kr / (ms * 1e-3), "unit": "env-steps/s", "steps": Kr, "ms_per_step": ms / Kr,
           "launches_per_step": 2 * len(ranges), "streams": len(ranges),
           "what": "per env-step: msort_policy_act (actor-critic 29-32-32-{22|1} on tcgen05, masked categorical draw) + fused "
                   "step(), env ranges on separate streams; obs/mask never leave HBM"}
    if env.kind != "mono":
        return two
    # the same loop as ONE kernel per env-step: msort_rollout_step = step(a_t) + the policy forward / masked draw for step
    # t+1 on the observation tile still in shared memory (the first action of the rollout comes from msort_policy_act)
    from marl_sortingenv_b200.ppo import flatten_parameters
    torch.manual_seed(0)
    pol = MaskableActorCritic(env.D, env.A).to(dev)
    flat = flatten_parameters(pol)
    packed = pack_actor_critic(pol)
    pf = env.rollout_pack(flat)
    outs = [out, (torch.empty(n, dtype=torch.int64, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev))]
    env.reset(seed=SEED)
    env.policy_act(packed, seed=ACTION_SEED, t=0, out=outs[0])
    for t in range(W):
        env.rollout_step(outs[t % 2][0], pf, ACTION_SEED, t + 1, outs[(t + 1) % 2])
    torch.cuda.synchronize(dev)
    assert env.step_variant == "hot_fused", env.step_variant
    gf = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gf):
        env.policy_act(packed, seed=ACTION_SEED, t=W, out=outs[W % 2])
        for t in range(Kr):
            env.rollout_step(outs[(W + t) % 2][0], pf, ACTION_SEED, W + t + 1, outs[(W + t + 1) % 2])
    gf.replay()
    ctx.barrier()
    r0.record(); gf.replay(); r1.record()
    ctx.barrier()
    msf = ctx.max_over_ranks(r0.elapsed_time(r1))
    # ... and as two kernels with the same arithmetic: msort_rollout_policy (128 threads per tile, 8 CTAs per SM) + msort_step
    gs = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gs):
        cur = torch.cuda.current_stream(dev)
        for s_ in streams:
            s_.wait_stream(cur)
        for t in range(Kr):
            for s_, r in zip(streams, ranges):
                with torch.cuda.stream(s_):
                    env.rollout_policy(pf, seed=ACTION_SEED, t=W + t, out=out, env_range=r if len(ranges) > 1 else None)
                    env.step(out[0], env_range=r if len(ranges) > 1 else None)
        for s_ in streams:
            cur.wait_stream(s_)
    gs.replay()
    ctx.barrier()
    r0.record(); gs.replay(); r1.record()
    ctx.barrier()
    mss = ctx.max_over_ranks(r0.elapsed_time(r1))
    split = {"value": n * world * Kr / (mss * 1e-3), "unit": "env-steps/s", "steps": Kr, "ms_per_step": mss / Kr,
             "launches_per_step": 2 * len(ranges), "streams": len(ranges),
             "what": "per env-step: msort_rollout_policy (the fused kernel's policy half as its own kernel, tiles by TMA) + fused step()"}
    fused = {"value": n * world * Kr / (msf * 1e-3), "unit": "env-steps/s", "steps": Kr, "ms_per_step": msf / Kr, "launches_per_step": 1}
    if mss < msf:
        return dict(split, variant="split", fused_one_kernel=fused, two_kernels_r01=two)
    return {"value": n * world * Kr / (msf * 1e-3), "unit": "env-steps/s", "steps": Kr, "ms_per_step": msf / Kr,
            "launches_per_step": 1, "streams": 1, "variant": env.step_variant, "split": split,
            "what": "per env-step ONE kernel (msort_rollout_step): fused step() + the next step's actor-critic forward "
                    "(29-32-32-{22|1}, tcgen05) and masked categorical draw on the observation tile still in shared memory; "
                    "one msort_policy_act per rollout for the first action",
            "two_kernels": two}


def measure_shard_invariance(ctx):
    """SURVEY.md section 8e on the hardware: the same 65 536 GLOBAL env ids, stepped 60 times under the rule-based
    action kernel (a pure function of the state), end in the same state whether one GPU holds them all or the
    ranks hold contiguous shards.  Per 8 192-env block a checksum of the state planes; all-gathered; rank 0 compares with
    its own single-GPU run of all 65 536."""
    torch, dist = ctx.torch, ctx.dist
    from marl_sortingenv_b200.batched import ENV_CLASSES
    G, B, T = 65536, 8192, 60

    def run(n, offset):
        env = ENV_CLASSES["mono"](n, device=ctx.dev, max_steps=25, seed=SEED, info_level="episode", global_env_offset=offset)
        env.reset(seed=SEED)
        act = torch.empty(n, dtype=torch.int64, device=ctx.dev)
        for _ in range(T):
            env.rule_based_actions(after_shift=True, out=act)
            env.step(act)
        planes = env.state.view(torch.int64).reshape(13, -1, 2)[:, :n]            # [plane, env, 2 words]
        w = torch.arange(1, 27, dtype=torch.int64, device=ctx.dev).reshape(13, 1, 2) * 0x9E3779B97F4A7C15 % (1 << 61)
        per_env = (planes * w).sum(dim=(0, 2)) + (planes >> 7).sum(dim=(0, 2))    # wrap-around int64 arithmetic
        pos = torch.arange(n, dtype=torch.int64, device=ctx.dev) % B + 1          # position inside the block matters too
        cs = (per_env * pos).reshape(-1, B).sum(dim=1)
        env.close()
        return cs

    per_rank = G // ctx.world
    mine = run(per_rank, ctx.rank * per_rank)
    if ctx.world > 1:
        parts = [torch.empty_like(mine) for _ in range(ctx.world)]
        dist.all_gather(parts, mine)
        sharded = torch.cat(parts)
    else:
        sharded = mine
    ok = None
    if ctx.rank == 0:
        single = run(G, 0) if ctx.world > 1 else run(G // 2, 0).new_empty(0)
        if ctx.world > 1:
            ok = bool(torch.equal(single, sharded))
        else:   # one GPU: the two halves stepped as separate handles must reproduce the whole
            a, b = run(G // 2, 0), run(G // 2, G // 2)
            ok = bool(torch.equal(torch.cat([a, b]), sharded))
        assert ok, "shard invariance violated: the state of a global env id depends on how the batch is sharded"
    return {"global_envs": G, "steps": T, "blocks": G // B, "shards": ctx.world if ctx.world > 1 else 2,
            "policy": "rule-based action kernel", "ok": ok}


def measure_configs(ctx, args, peak, peak_src):
    """The BASELINE configurations the headline does not cover, each device-timed in this run:
    configs[1] Env_1 65 536 envs (PHILOX hot kernel and the REPLAY instantiation on device-resident streams),
    configs[2] Env_2 262 144 envs with the embedded policy, configs[4] Env_3 8 388 608 envs in total over the ranks."""
    torch = ctx.torch
    out = {}
    W, K = 5, 200
    if ctx.rank == 0:
        # ---- Env_1, 65 536 envs
        env = make_env(ctx_single(ctx), "sort", 65536)
        acts, _ = record_actions(env, W + K)
        ms, per, lps, _ = timed_steps(ctx_single(ctx), env, acts, W, K, step_streams=1)
        out["env1_65536_philox"] = cfg_entry("sort", 65536, 1, ms, K, peak, peak_src, env.step_variant, lps,
                                             note="fits L2 (13 MB of state): not an HBM measurement")
        env.close()
        out["env1_65536_replay"] = replay_config(ctx, peak, peak_src)
        # ---- Env_2, 262 144 envs, embedded policy
        for label, tensor in (("env2_262144_mlp", 1), ("env2_262144_mlp_ffma2", 0)):
            from marl_sortingenv_b200 import _abi
            env = make_env(ctx_single(ctx), "press", 262144)
            env.set_option(_abi.OPT_TENSOR_POLICY, tensor)
            acts, _ = record_actions(env, W + K)
            out[label] = env2_entry(ctx, env, acts, 262144, W, K, peak, peak_src)
            env.close()
        env = make_env(ctx_single(ctx), "press", 1 << 20)
        acts, _ = record_actions(env, W + 100)
        out["env2_1048576_mlp"] = env2_entry(ctx, env, acts, 1 << 20, W, 100, peak, peak_src)
        env.close()
        del acts
        torch.cuda.empty_cache()
    # ---- Env_3, 8 388 608 envs in total, sharded over the ranks (strong scaling)
    ctx.barrier()
    n = STRONG_GLOBAL_ENVS // ctx.world
    Ks = 40
    env = make_env(ctx, "mono", n, global_envs=STRONG_GLOBAL_ENVS)
    acts, _ = record_actions(env, W + Ks)
    ms, per, lps, _ = timed_steps(ctx, env, acts, W, Ks, step_streams=2)
    if ctx.rank == 0:
        e = cfg_entry("mono", n, ctx.world, ms, Ks, peak, peak_src, env.step_variant, lps)
        e["global_envs"] = STRONG_GLOBAL_ENVS
        out["env3_8388608_strong"] = e
    env.close()
    return out


def env2_entry(ctx, env, acts, n, W, K, peak, peak_src):
    """Env_2 timed the way the headline is — each step as two env-range launches on two streams (`msort_step_range`) — and as
    one whole-batch launch per step; the entry is the two-range run, the single launch rides along as `one_launch`."""
    ms1, _, lps1, _ = timed_steps(ctx_single(ctx), env, acts, W, K, step_streams=1)
    one = cfg_entry("press", n, 1, ms1, K, peak, peak_src, env.step_variant, lps1)
    ms2, _, lps2, _ = timed_steps(ctx_single(ctx), env, acts, W, K, step_streams=2)
    e = cfg_entry("press", n, 1, ms2, K, peak, peak_src, env.step_variant, lps2)
    e["one_launch"] = {k: one[k] for k in ("value", "us_per_step", "frac", "launches_per_step")}
    return e


class _Single:
    """A view of the context that behaves like a one-rank job (rank-0-only measurements)."""

    def __init__(self, ctx):
        self.torch, self.dev, self.world, self.rank = ctx.torch, ctx.dev, 1, 0

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        return float(x)


def ctx_single(ctx):
    return _Single(ctx)


def cfg_entry(kind, n, world, total_ms, K, peak, peak_src, variant, lps, extra_bytes=0, note=None):
    ms = total_ms / K
    bytes_per = ALGO_BYTES_PER_STEP[kind] + extra_bytes
    e = {"env": KIND_NAME[kind], "envs_per_gpu": n, "n_gpus": world, "value": n * world / (ms * 1e-3), "unit": "env-steps/s",
         "us_per_step": ms * 1e3, "steps": K, "variant": variant, "launches_per_step": lps,
         "bytes_per_env_step": bytes_per, "frac": bytes_per * n / (ms * 1e-3) / 1e9 / peak, "peak": peak}
    if note:
        e["note"] = note
    return e


def replay_config(ctx, peak, peak_src):
    """BASELINE configs[1]: Env_1, 65 536 envs, random actions, REPLAY instantiation consuming device-resident input
    streams (uniforms of the shape the reference's generators produce: 4 noise doubles per step, one double per
    redistribution draw; bit-exactness of this path is tests/test_cuda_parity.py::test_config2_env1_65536_replay_bit_exact)."""
    torch = ctx.torch
    from marl_sortingenv_b200.batched import ENV_CLASSES
    n, T, L = 65536, 50, 23 * 50
    g = torch.Generator(device=ctx.dev).manual_seed(1234)
    env = ENV_CLASSES["sort"](n, device=ctx.dev, max_steps=T, seed=SEED, info_level="episode", rng_mode="replay")
    noise = torch.rand((T, n, 4), dtype=torch.float64, device=ctx.dev, generator=g)
    redis = torch.rand((n, L), dtype=torch.float64, device=ctx.dev, generator=g)
    press = torch.randint(0, 11, (T, n), dtype=torch.uint8, device=ctx.dev, generator=g)
    acts = torch.randint(0, 2, (T, n), dtype=torch.int64, device=ctx.dev, generator=g)
    first = torch.randint(1, 3, (n,), dtype=torch.uint8, device=ctx.dev, generator=g)

    def episode():
        env.reset(seed=SEED, first_pattern=first)
        for t in range(T):
            env.step(acts[t], replay=dict(noise_u=noise[t], redis_u=redis, press_choice=press[t]))
    episode()
    torch.cuda.synchronize(ctx.dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        episode()
    graph.replay()
    torch.cuda.synchronize(ctx.dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); graph.replay(); e1.record()
    torch.cuda.synchronize(ctx.dev)
    env.sync_check()
    ms = e0.elapsed_time(e1)
    e = cfg_entry("sort", n, 1, ms, T, peak, peak_src, env.step_variant, 1, extra_bytes=REPLAY_EXTRA_BYTES,
                  note="one 50-step episode incl. its reset kernel; REPLAY streams (708 MB) resident in HBM")
    env.close()
    return e


# ----------------------------------------------------------------------------- CUDA arm
def run_msort(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:           # convenience: re-launch ourselves under torchrun
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    claim_stdout()
    ctx = Ctx(args)
    torch, dev, rank = ctx.torch, ctx.dev, ctx.rank
    from marl_sortingenv_b200.sharding import StatsReducer

    n, K, W = envs_per_rank(args, world), args.steps, max(args.warmup, 3)
    env = make_env(ctx, args.kind, n)
    actions, final_ref = record_actions(env, W + K)

    # ---- timed region: K fused step() launches back to back; the statistics all-reduce (the only collective) is issued
    #      on a side stream right after the closing event and joined later — it is not on the stepping stream
    reducer = StatsReducer(dev)
    if world > 1:                                  # NCCL communicators are created lazily: do it before timing
        reducer.start(torch.zeros(16, dtype=torch.float64, device=dev)); reducer.result(); torch.cuda.synchronize(dev)
    ar = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]

    def after():
        reducer.start(env.stats)
        with torch.cuda.stream(reducer.side):
            pass
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.25)
    total_ms, per_kernel_ms, lps, (t0, t1) = timed_steps(ctx, env, actions, W, K, step_streams=args.step_streams,
                                                         use_graph=not args.no_graph, after=after)
    clocks = sampler.stop(t0, t1) if sampler else None
    assert torch.equal(env.state, final_ref), "timed run diverged from the recorded run (non-determinism)"
    value = n * world * K / (total_ms * 1e-3)
    stats = reducer.result().clone()
    # the collective alone, device-timed on its side stream (10 back-to-back reductions)
    allreduce_us = None
    if world > 1:
        ctx.barrier()
        with torch.cuda.stream(reducer.side):
            ar[0].record()
            for _ in range(10):
                ctx.dist.all_reduce(reducer.buf[0])
            ar[1].record()
        torch.cuda.synchronize(dev)
        allreduce_us = ctx.max_over_ranks(ar[0].elapsed_time(ar[1]) * 100.0)
        stats = stats                                   # (buf[0] was scratch for the timing; `stats` is a clone)
    stats = stats.cpu().tolist()
    gpu_launches = K * lps
    variant = env.step_variant

    shard_inv = measure_shard_invariance(ctx)
    e2e = None if args.no_e2e else measure_e2e(ctx, args, env, actions, W, K)
    rollout = None if args.no_rollout else measure_rollout(ctx, args, env, W)
    env.close()
    del actions, final_ref
    torch.cuda.empty_cache()
    peak, peak_src = measured_peak()
    configs = None if args.no_configs else measure_configs(ctx, args, peak, peak_src)

    if rank != 0:
        ctx.close()
        return
    kern_ms = statistics.mean(per_kernel_ms)
    line = {
        "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches),
        "roofline": roofline(args.kind, n, kern_ms, peak, peak_src, variant, lps),
        "episode_stats": {"episodes": stats[0], "mean_return": stats[1] / max(1.0, stats[0]),
                          "mean_length": stats[2] / max(1.0, stats[0]), "env_steps": stats[3],
                          "bales": stats[6], "reduced_over_ranks": world},
        "shard_invariance": shard_inv, "numa": ctx.numa,
    }
    if allreduce_us is not None:
        line["allreduce_us"] = allreduce_us
    if rollout is not None:
        line["rollout"] = rollout
    if configs is not None:
        line["configs"] = configs
    if not args.no_cpu_baseline:
        v, cores, sample, _, _ = cpu_rollout(args.kind, budget_s=12.0)
        line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample}
        ref = python_reference(args.kind, steps=20, warmup=3)
        if ref is not None:
            line["cpu_reference"] = {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")}
    emit(line)
    ctx.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_msort(args)


if __name__ == "__main__":
    main()
