#!/usr/bin/env python
"""Regenerates tests/golden/reference_trajectories.npz from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference; see oracle/ref_loader.py for how
the reference is imported without gymnasium/matplotlib).  The fixture holds, for a grid of
(env kind × action masking × overflow check × noise × seeds), the recorded replay inputs
(actions, numpy RNG uniforms, generator batches, Env_1 press choices, sort modes) and the
reference's outputs after every step (obs, reward, flags, masks, full integer state).

    python tests/golden/make_golden.py            # rewrite the fixture
    python tests/golden/make_golden.py --check    # regenerate in memory and diff against the file

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

OUT = os.path.join(HERE, "reference_trajectories.npz")

SMALL = {"action": np.int16, "input_counts": np.uint8, "press_choice": np.uint8,
         "sort_mode": np.uint8, "terminated": np.bool_, "overflow": np.bool_,
         "overflow_material": np.int8, "mask": np.bool_, "state": np.int32,
         "n_draws": np.int16, "first_pattern0": np.uint8}


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


def build():
    out = {"numpy_version": np.asarray(np.__version__)}
    names = []
    for name, meta in groups():
        weights = None
        if meta.get("mlp"):
            weights = sb3_style_mlp_weights(meta["mlp_seed"], meta["mlp_gain"])
            out[f"{name}/mlp_weights"] = weights
        recs = [record(meta["kind"], seed=s, steps=meta["steps"], max_steps=meta["max_steps"],
                       noise=meta["noise"], balesize=meta["balesize"], policy=meta["policy"],
                       action_seed=1000 + s, use_action_masking=meta["use_action_masking"],
                       check_overflow=meta["check_overflow"], auto_reset=meta["auto_reset"],
                       mlp_weights=weights) for s in meta["seeds"]]
        b = stack_recordings(recs)
        T = b.pop("T")
        assert T == meta["steps"]
        for k, v in b.items():
            out[f"{name}/{k}"] = v.astype(SMALL[k]) if k in SMALL else v
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
