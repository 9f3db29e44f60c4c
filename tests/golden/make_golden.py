#!/usr/bin/env python
"""Regenerates tests/golden/reference_trajectories.npz from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference; see oracle/ref_loader.py for how
the reference is imported without gymnasium/matplotlib).  The fixture holds, for a grid of
(env kind × action masking × overflow check × noise × seeds), the recorded replay inputs
(actions, numpy RNG uniforms, generator batches, Env_1 press choices, sort modes) and the
reference's outputs after every step (obs, reward, flags, masks, full integer state).

    python tests/golden/make_golden.py [--set r01|r02]            # rewrite the fixture(s)
    python tests/golden/make_golden.py [--set r01|r02] --check    # regenerate in memory and diff against the file

Two fixture files: `reference_trajectories.npz` (set r01, round 1) and `reference_trajectories_r02.npz` (set r02:
non-whole-percent quality thresholds, a 16-seed grid, 200-step episodes without action masking).

Reproducibility.  A recording is a deterministic function of its seeds UNTIL the episode's first unseeded reset:
`Env_Super.reset(seed=None)` rebuilds the input generator from OS entropy (env_super.py:375,
`SeasonalInputGenerator(seed=None)`), so the pattern order of every later episode — and everything downstream of
it — differs from run to run (the fixture records the order that occurred, `first_pattern`, which is why replay
still works).  `--check` therefore compares every recording up to and including its first terminated step, plus
the replay inputs that do not depend on the pattern order (actions are drawn from the mask, so they do), and
fails on any difference there.

numpy version used for the committed file: see the `numpy_version` entry inside it.
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

from oracle.ref_record import record, sb3_style_mlp_weights  # noqa: E402
from parity_util import stack_recordings  # noqa: E402

OUT = {"r01": os.path.join(HERE, "reference_trajectories.npz"),
       "r02": os.path.join(HERE, "reference_trajectories_r02.npz")}

SMALL = {"action": np.int16, "input_counts": np.uint8, "press_choice": np.uint8,
         "sort_mode": np.uint8, "terminated": np.bool_, "overflow": np.bool_,
         "overflow_material": np.int8, "mask": np.bool_, "state": np.int32,
         "n_draws": np.int16, "first_pattern0": np.uint8}


def groups_r02():
    g = []
    # --- quality thresholds that are not whole percents: an EMPTY container reports Python's round(threshold, 2)
    #     (env_super.py:788-789) and the purity-difference observation round(purity - threshold, 2) (:222-225)
    for kind in ("sort", "mono"):
        g.append((f"thr_{kind}", dict(kind=kind, max_steps=50, noise=0.05, balesize=200, use_action_masking=True,
                                      check_overflow=False, auto_reset=True, steps=110, seeds=[21, 22, 23, 24],
                                      policy="masked_random", thresholds=[0.875, 0.905, 0.9, 0.815])))
    # --- the grid again on 16 seeds
    for kind in ("sort", "press", "mono"):
        for masking in (True, False):
            g.append((f"grid16_{kind}_m{int(masking)}", dict(
                kind=kind, max_steps=50, noise=0.05, balesize=200, use_action_masking=masking, check_overflow=False,
                auto_reset=True, steps=110, seeds=list(range(101, 117)),
                policy="masked_random" if masking else "uniform")))
    # --- 200-step episodes (the reference's benchmark length, main.py:42-51) without masking and without a reset:
    #     invalid actions are sanitised every few steps (Env_3 then skips the press timers, env_monolith.py:133-138)
    for kind in ("sort", "press", "mono"):
        g.append((f"noreset_{kind}_m0", dict(kind=kind, max_steps=200, noise=0.05, balesize=200,
                                             use_action_masking=False, check_overflow=False, auto_reset=False,
                                             steps=200, seeds=[31, 32, 33, 34], policy="uniform")))
    return g


def groups():
    g = []
    # --- Appendix C known-answer trajectories: seed 42, 200 steps, noise 0
    for kind, pol in (("sort", "rule"), ("press", "first_valid"), ("mono", "second_valid")):
        g.append((f"kat_{kind}", dict(kind=kind, max_steps=200, noise=0.0, balesize=200,
                                      use_action_masking=True, check_overflow=False,
                                      auto_reset=False, steps=200, seeds=[42], policy=pol)))
    # --- the grid: masking x overflow x noise, auto-reset across 2+ episodes
    for kind in ("sort", "press", "mono"):
        for masking in (True, False):
            for ovf in (False, True):
                for noise in (0.05, 0.0):
                    name = f"grid_{kind}_m{int(masking)}_o{int(ovf)}_n{int(noise * 100):02d}"
                    g.append((name, dict(kind=kind, max_steps=50, noise=noise, balesize=200,
                                         use_action_masking=masking, check_overflow=ovf,
                                         auto_reset=True, steps=110, seeds=[1, 2, 3],
                                         policy="masked_random" if masking else "uniform")))
    # --- parameter variations (ctor arguments the reference exposes)
    for kind in ("sort", "press", "mono"):
        g.append((f"var_{kind}", dict(kind=kind, max_steps=30, noise=0.1, balesize=150,
                                      use_action_masking=True, check_overflow=False,
                                      auto_reset=True, steps=70, seeds=[7, 8],
                                      policy="masked_random")))
    # --- generator rule (no recorded batches): one seeded episode, no auto-reset
    for kind in ("sort", "mono"):
        g.append((f"gen_{kind}", dict(kind=kind, max_steps=90, noise=0.05, balesize=200,
                                      use_action_masking=True, check_overflow=False,
                                      auto_reset=False, steps=90, seeds=[0, 3, 4, 42],
                                      policy="masked_random")))
    # --- Env_3.step(mode='rule_based'): the reference's own heuristic picks the action inside step()
    g.append(("rule_mono", dict(kind="mono", max_steps=200, noise=0.0, balesize=200, use_action_masking=True,
                                check_overflow=False, auto_reset=False, steps=200, seeds=[1, 2, 3],
                                policy="mode_rule_based")))
    # --- Env_2 with an embedded sort-policy MLP (both modes occur with action gain 1.0)
    g.append(("mlp_press", dict(kind="press", max_steps=50, noise=0.05, balesize=200,
                                use_action_masking=True, check_overflow=False, auto_reset=True,
                                steps=110, seeds=[11, 12, 13], policy="masked_random", mlp=True,
                                mlp_seed=19, mlp_gain=1.0)))
    return g


def build(which="r01"):
    out = {"numpy_version": np.asarray(np.__version__)}
    names = []
    for name, meta in (groups() if which == "r01" else groups_r02()):
        weights = None
        if meta.get("mlp"):
            weights = sb3_style_mlp_weights(meta["mlp_seed"], meta["mlp_gain"])
            out[f"{name}/mlp_weights"] = weights
        recs = [record(meta["kind"], seed=s, steps=meta["steps"], max_steps=meta["max_steps"],
                       noise=meta["noise"], balesize=meta["balesize"], policy=meta["policy"],
                       action_seed=1000 + s, use_action_masking=meta["use_action_masking"],
                       check_overflow=meta["check_overflow"], auto_reset=meta["auto_reset"],
                       mlp_weights=weights, thresholds=meta.get("thresholds")) for s in meta["seeds"]]
        b = stack_recordings(recs)
        T = b.pop("T")
        assert T == meta["steps"]
        for k, v in b.items():
            out[f"{name}/{k}"] = v.astype(SMALL[k]) if k in SMALL else v
        out[f"{name}/meta"] = np.asarray(json.dumps(meta))
        names.append(name)
    out["groups"] = np.asarray(json.dumps(names))
    return out


PER_STEP = ("action", "noise_u", "input_counts", "press_choice", "sort_mode", "obs", "reward", "terminated",
            "overflow", "overflow_material", "mask", "state", "acc_belt", "n_draws", "mlp_margin")


def check(which):
    """Regenerate in memory and compare with the committed file on everything a re-run can reproduce."""
    data, old = build(which), np.load(OUT[which], allow_pickle=False)
    names = json.loads(str(data["groups"]))
    bad = []
    if json.loads(str(old["groups"])) != names:
        bad.append("groups")
    for name in names:
        for k in (k for k in data if k.startswith(name + "/")):
            new, ref = data[k], np.asarray(old[k]) if k in old.files else None
            kk = k[len(name) + 1:]
            if ref is None:
                bad.append(k)
            elif kk in PER_STEP:
                term = np.asarray(old[f"{name}/terminated"])            # [T, R]
                for r in range(term.shape[1]):
                    done = np.flatnonzero(term[:, r])
                    t_end = (int(done[0]) + 1) if done.size else term.shape[0]   # up to and including the first terminated step
                    if new.shape != ref.shape or not np.array_equal(new[:t_end, r], ref[:t_end, r]):
                        bad.append(f"{k}[:{t_end}, {r}]")
                        break
            elif kk == "redis_u":                                         # the stream's prefix consumed by the first episode
                nd_new, nd_old = data[f"{name}/n_draws"], np.asarray(old[f"{name}/n_draws"])
                term = np.asarray(old[f"{name}/terminated"])
                for r in range(term.shape[1]):
                    done = np.flatnonzero(term[:, r])
                    t_end = (int(done[0]) + 1) if done.size else term.shape[0]
                    L = int(nd_old[:t_end, r].sum())
                    if int(nd_new[:t_end, r].sum()) != L or not np.array_equal(new[r, :L], ref[r, :L]):
                        bad.append(f"{k}[{r}, :{L}]")
                        break
            elif not np.array_equal(new, ref):                            # meta, mlp_weights, obs0, first_pattern0
                bad.append(k)
    print(f"[{which}] " + ("MISMATCH: " + ", ".join(bad) if bad else
                           f"fixture reproduces up to each recording's first unseeded reset ({len(data)} arrays, {len(names)} groups)"))
    return 1 if bad else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--set", default="all", choices=["all", "r01", "r02"])
    a = ap.parse_args()
    rc = 0
    for which in (["r01", "r02"] if a.set == "all" else [a.set]):
        if a.check:
            rc |= check(which)
            continue
        data = build(which)
        np.savez_compressed(OUT[which], **data)
        print(f"wrote {OUT[which]}: {os.path.getsize(OUT[which]) / 1e6:.2f} MB, {len(data)} arrays")
    return rc


if __name__ == "__main__":
    sys.exit(main())
