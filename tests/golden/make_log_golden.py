#!/usr/bin/env python
"""Regenerates tests/golden/reference_logs.npz from the UNMODIFIED reference: single episodes with
the replay inputs needed to re-run them AND the reference's own per-step Python logs —
`reward_data` (env_super.py:402-408, 928-946), `press_actions_per_timestep` (env_super.py:631-637,
729-736; env_2_press.py:127-131; env_monolith.py:132-138) and `bale_count` (env_super.py:661-687) —
which the telemetry recorder (marl-sortingenv_b200/telemetry.py) must reproduce from device snapshots.

Runs only in the build container (needs /root/reference).
    python tests/golden/make_log_golden.py            # rewrite the fixture
    python tests/golden/make_log_golden.py --check    # regenerate in memory and diff
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.ref_record import ALL5, record  # noqa: E402

OUT = os.path.join(HERE, "reference_logs.npz")
NAME_BASE = 100   # a material NAME in a press-log tuple is stored as NAME_BASE + index; None as -1


def encode_logs(env):
    rd = env.reward_data
    T = len(rd["Reward"])
    out = {
        "log_reward": np.asarray(rd["Reward"], dtype=np.float64).reshape(T, 2),
        "log_total": np.asarray(rd["Total"], dtype=np.float64),
        "log_setting": np.asarray(rd["Setting"], dtype=np.int8),
        "log_belt_occ": np.asarray(rd["Belt_Occupancy"], dtype=np.float64),
        "log_belt_prop": np.asarray([[bp[m] for m in ALL5[:4]] for bp in rd["Belt_Proportions"]],
                                    dtype=np.float64).reshape(T, 4),   # dicts {material: share} (env_super.py:199-210)
        "log_true": np.asarray([[rd[f"{m}_True"][t] for m in ALL5] for t in range(T)], dtype=np.int32),
        "log_false": np.asarray([[rd[f"{m}_False"][t] for m in ALL5] for t in range(T)], dtype=np.int32),
        "log_accuracy": np.asarray(rd["Accuracy"], dtype=np.float64),
    }
    pl = []
    for code, mat in env.press_actions_per_timestep:
        if mat is None:
            m = -1
        elif isinstance(mat, str):
            m = NAME_BASE + ALL5.index(mat)
        else:
            m = int(mat)
        pl.append((0 if code is None else int(code), m))
    out["press_log"] = np.asarray(pl, dtype=np.int16).reshape(len(pl), 2)
    bm, bs, bq = [], [], []
    for mi, m in enumerate(ALL5):
        for size, q in env.bale_count[m]:
            bm.append(mi); bs.append(int(size)); bq.append(int(q))
    out["bale_mat"] = np.asarray(bm, dtype=np.int8)
    out["bale_size"] = np.asarray(bs, dtype=np.int32)
    out["bale_q"] = np.asarray(bq, dtype=np.int16)
    return out


def groups():
    g = []
    for kind in ("sort", "press", "mono"):
        for masking in (True, False):
            for seed in (5, 6):
                g.append((f"log_{kind}_m{int(masking)}_s{seed}",
                          dict(kind=kind, max_steps=120, noise=0.05, balesize=200, use_action_masking=masking,
                               check_overflow=False, auto_reset=False, steps=120, seed=seed,
                               policy="masked_random" if masking else "uniform")))
    # overflow termination ends the episode early and logs the split penalty
    g.append(("log_mono_overflow", dict(kind="mono", max_steps=120, noise=0.05, balesize=200, use_action_masking=False,
                                        check_overflow=True, auto_reset=False, steps=120, seed=9, policy="uniform")))
    return g


def build():
    out = {"numpy_version": np.asarray(np.__version__)}
    names = []
    for name, meta in groups():
        r = record(meta["kind"], seed=meta["seed"], steps=meta["steps"], max_steps=meta["max_steps"],
                   noise=meta["noise"], balesize=meta["balesize"], policy=meta["policy"],
                   action_seed=2000 + meta["seed"], use_action_masking=meta["use_action_masking"],
                   check_overflow=meta["check_overflow"], auto_reset=False, keep_env=True)
        env = r.pop("env")
        # the run stops being meaningful after the episode ends: keep the steps up to termination
        term = np.flatnonzero(r["terminated"])
        T = int(term[0]) + 1 if term.size else meta["steps"]
        n_draws = int(r["n_draws"][:T].sum())
        for k in ("action", "noise_u", "input_counts", "press_choice", "sort_mode", "reward", "terminated",
                  "overflow", "state", "obs"):
            out[f"{name}/{k}"] = r[k][:T]
        out[f"{name}/redis_u"] = r["redis_u"][:n_draws]
        out[f"{name}/first_pattern0"] = np.asarray(r["first_pattern0"], dtype=np.uint8)
        logs = encode_logs(env)
        # the reference kept stepping (and logging) after termination when steps > T: cut the per-step lists
        for k, v in logs.items():
            if k.startswith("log_"):
                v = v[:T]
            out[f"{name}/{k}"] = v
        if T < meta["steps"]:   # press log / bale lists of a cut run would include later steps: re-record exactly T steps
            r2 = record(meta["kind"], seed=meta["seed"], steps=T, max_steps=meta["max_steps"], noise=meta["noise"],
                        balesize=meta["balesize"], policy=meta["policy"], action_seed=2000 + meta["seed"],
                        use_action_masking=meta["use_action_masking"], check_overflow=meta["check_overflow"],
                        auto_reset=False, keep_env=True)
            for k, v in encode_logs(r2["env"]).items():
                out[f"{name}/{k}"] = v
        meta = dict(meta, T=T)
        out[f"{name}/meta"] = np.asarray(json.dumps(meta))
        names.append(name)
    out["groups"] = np.asarray(json.dumps(names))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    data = build()
    if a.check:
        old = np.load(OUT, allow_pickle=False)
        bad = [k for k in data if k != "numpy_version" and not np.array_equal(np.asarray(old[k]), data[k])]
        print("MISMATCH: " + ", ".join(bad) if bad else f"fixture reproduces ({len(data)} arrays)")
        return 1 if bad else 0
    np.savez_compressed(OUT, **data)
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1e6:.2f} MB, {len(data)} arrays")
    return 0


if __name__ == "__main__":
    sys.exit(main())
