#!/usr/bin/env python
"""Writes tests/golden/reference_philox.npz: trajectories of the UNMODIFIED reference on the random inputs of the
PRODUCTION (Philox) generator — the fixture that pins the PHILOX / FAST / HOT instantiations of the step kernel to the
reference itself, bit for bit, not only to the oracle.

How.  The oracle's PHILOX step records what it drew in REPLAY form (noise uniforms; one uniform per redistribution
`choice` call, chosen so that numpy's cdf search lands on the unit the Philox draw took; Env_1's press choice; the
pattern order of every episode) — msort_oracle.c, mso_step_out_t rec_*.  oracle/ref_drive.py feeds exactly these to
reference envs (stand-ins for the env's three generator attributes; the reference source is untouched).  This script
asserts that reference and oracle then agree on EVERY step (integer state, masks, flags bit-exact; obs f32
bit-identical; rewards to 1e-12) and stores the reference's outputs.  On the GPU box the CUDA kernels, keyed with the
same seed / global env ids / actions, must reproduce them (tests/test_cuda_parity.py::test_cuda_philox_matches_reference_*).

    python tests/golden/make_philox_golden.py [--check]
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

from oracle.cpu_oracle import OracleEnv  # noqa: E402
from oracle.ref_drive import DrivenReferenceEnv  # noqa: E402
from oracle.ref_record import snapshot  # noqa: E402
from parity_util import _col, config_for, state_rows  # noqa: E402

OUT = os.path.join(HERE, "reference_philox.npz")

GROUPS = [
    # name, meta
    ("philox_sort", dict(kind="sort", max_steps=50, noise=0.05, balesize=200, use_action_masking=True, check_overflow=False,
                         auto_reset=True, steps=130, envs=24, seed=77, offset=1000)),
    ("philox_press", dict(kind="press", max_steps=50, noise=0.05, balesize=200, use_action_masking=True, check_overflow=False,
                          auto_reset=True, steps=130, envs=24, seed=78, offset=0)),
    ("philox_mono", dict(kind="mono", max_steps=50, noise=0.05, balesize=200, use_action_masking=True, check_overflow=False,
                         auto_reset=True, steps=130, envs=32, seed=79, offset=123456789012)),
    ("philox_mono_unmasked", dict(kind="mono", max_steps=40, noise=0.08, balesize=150, use_action_masking=False,
                                  check_overflow=True, auto_reset=True, steps=100, envs=24, seed=80, offset=5)),
]


def build_group(meta):
    n, T, kind = meta["envs"], meta["steps"], meta["kind"]
    cfg = config_for(meta, n, rng_mode="philox", seed=meta["seed"], global_env_offset=meta["offset"])
    ora = OracleEnv(cfg, nthreads=4)
    obs0, mask0 = ora.reset()
    refs = [DrivenReferenceEnv(kind, max_steps=meta["max_steps"], noise=meta["noise"], balesize=meta["balesize"]) for _ in range(n)]
    for i, r in enumerate(refs):
        o = r.reset(int(ora.state["gen_first"][i]), seed=0)
        assert np.array_equal(o, obs0[i])
    A = mask0.shape[1]
    rng = np.random.default_rng(meta["seed"] + 1)
    out = {k: [] for k in ("action", "obs", "reward", "terminated", "overflow", "mask", "state", "n_draws")}
    for t in range(T):
        a = ora.sample_masked_actions(3, t) if meta["use_action_masking"] else rng.integers(0, A, size=n)
        oo, orw, ot, om, oi = ora.step(a, record=True)
        ro, rr, rt, rm, rs, rov = [], [], [], [], [], []
        for i, r in enumerate(refs):
            k = int(oi["rec_n_draws"][i])
            o, rew, term, info = r.step(int(a[i]), oi["rec_noise_u"][i], oi["rec_redis_u"][i, :k],
                                        press_choice=int(oi["rec_press_choice"][i]),
                                        use_action_masking=meta["use_action_masking"], check_overflow=meta["check_overflow"])
            ro.append(o); rr.append(rew); rt.append(term); rov.append(bool(info.get("overflow", False)))
            rm.append(np.asarray(r.env.action_masks(), dtype=bool)); rs.append(snapshot(r.env))
            if term:                                  # VecEnv auto-reset: unseeded reset; the pattern order of the new episode as the oracle drew it
                r.reset(int(ora.state["gen_first"][i]))
        ro, rr, rt, rm, rs = np.stack(ro), np.asarray(rr), np.asarray(rt), np.stack(rm), np.stack(rs)
        # ---- the oracle's PHILOX step == the reference on the same random inputs
        assert np.array_equal(rt, ot), f"step {t}: terminated"
        live = ~rt
        assert np.array_equal(state_rows(ora.state)[live], rs[live]), f"step {t}: integer state"
        assert np.array_equal(om[live], rm[live]), f"step {t}: mask"
        assert np.array_equal(oo[live], ro[live]), f"step {t}: obs not bit-identical"
        if rt.any():
            assert np.array_equal(oi["terminal_obs"][rt], ro[rt]), f"step {t}: terminal obs"
            assert np.array_equal(oi["episode_length"][rt], rs[rt][:, _col("step")]), f"step {t}: episode length"
        assert np.allclose(orw, rr, rtol=0, atol=1e-12), f"step {t}: reward"
        assert np.array_equal(oi["overflow"].astype(bool), np.asarray(rov)), f"step {t}: overflow"
        out["action"].append(np.asarray(a, dtype=np.int16)); out["obs"].append(ro); out["reward"].append(rr)
        out["terminated"].append(rt); out["overflow"].append(np.asarray(rov)); out["mask"].append(rm)
        out["state"].append(rs.astype(np.int32)); out["n_draws"].append(oi["rec_n_draws"].astype(np.int16))
    res = {k: np.stack(v) for k, v in out.items()}
    res["obs0"] = obs0
    return res


def build():
    data, names = {"numpy_version": np.asarray(np.__version__)}, []
    for name, meta in GROUPS:
        g = build_group(meta)
        for k, v in g.items():
            data[f"{name}/{k}"] = v
        data[f"{name}/meta"] = np.asarray(json.dumps(meta))
        names.append(name)
        print(f"{name}: {meta['envs']} envs x {meta['steps']} steps, {int(g['terminated'].sum())} episodes ended, "
              f"{float(g['n_draws'].mean()):.1f} redistribution draws per env-step — reference == oracle on every step", flush=True)
    data["groups"] = np.asarray(json.dumps(names))
    return data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    data = build()
    if a.check:      # fully deterministic: every random input is prescribed, unseeded resets included
        old = np.load(OUT, allow_pickle=False)
        bad = [k for k in data if k != "numpy_version" and not np.array_equal(np.asarray(old[k]), data[k])]
        print("MISMATCH: " + ", ".join(bad) if bad else f"fixture reproduces ({len(data)} arrays)")
        return 1 if bad else 0
    np.savez_compressed(OUT, **data)
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1e6:.2f} MB")
    return 0


if __name__ == "__main__":
    sys.exit(main())
