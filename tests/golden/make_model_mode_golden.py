#!/usr/bin/env python
"""Writes tests/golden/reference_model_mode.npz from the UNMODIFIED reference (build container only; see
oracle/ref_loader.py): Env_3_Monolith.step(action=None, mode='model') (env_monolith.py:186-221) with a sort
agent and a maskable press agent assigned — deterministic stand-ins (oracle/ref_record.py SortStubAgent /
MaskablePressStubAgent) that also remember the observations the reference showed them.  Per step: those
observations (`agent_obs` = sort obs | press obs, taken after update_environment's shift), the action the
reference composed from the two answers, and the usual replay inputs / outputs.

    python tests/golden/make_model_mode_golden.py [--check]
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.ref_record import record  # noqa: E402
from parity_util import stack_recordings  # noqa: E402

OUT = os.path.join(HERE, "reference_model_mode.npz")
META = dict(kind="mono", max_steps=200, noise=0.05, balesize=200, use_action_masking=True, check_overflow=False,
            auto_reset=False, steps=200, seeds=[5, 6, 7, 8], policy="mode_model")


def build():
    recs = [record("mono", seed=s, steps=META["steps"], max_steps=META["max_steps"], noise=META["noise"],
                   balesize=META["balesize"], policy=META["policy"], action_seed=1000 + s,
                   use_action_masking=True, check_overflow=False, auto_reset=False) for s in META["seeds"]]
    b = stack_recordings(recs)
    b.pop("T")
    b["agent_obs"] = np.stack([r["agent_obs"] for r in recs], axis=1)
    out = {f"model_mono/{k}": v for k, v in b.items()}
    out["model_mono/meta"] = np.asarray(json.dumps(META))
    out["numpy_version"] = np.asarray(np.__version__)
    return out


def main():
    data = build()
    if "--check" in sys.argv:
        old = np.load(OUT, allow_pickle=False)
        bad = [k for k in data if k != "numpy_version" and not np.array_equal(np.asarray(old[k]), data[k])]
        print("MISMATCH: " + ", ".join(bad) if bad else f"fixture reproduces ({len(data)} arrays)")
        return 1 if bad else 0
    np.savez_compressed(OUT, **data)
    a = data["model_mono/action"]
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1e3:.0f} kB; sort mode 1 on {np.mean(a >= 11):.2f} of the steps, "
          f"{np.mean(a % 11 != 0):.2f} press actions, {len(np.unique(a))} distinct actions")
    return 0


if __name__ == "__main__":
    sys.exit(main())
