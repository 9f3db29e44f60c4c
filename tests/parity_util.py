"""Shared helpers for the parity tests: drive a backend (CPU oracle or the CUDA library, both
exposing reset()/step() over numpy arrays) with the replay inputs of recorded reference
trajectories and compare every step against the recorded outputs."""
from __future__ import annotations

import numpy as np

from marl_sortingenv_b200 import _abi, make_config

# layout of the `state` rows stored in the fixtures (oracle/ref_record.py STATE_FIELDS)
STATE_FIELDS = [
    ("input", 4), ("belt", 4), ("sorting", 4),
    ("cont_true", 4), ("cont_false", 4), ("cont_e", 1),
    ("press_timer", 2), ("press_mat", 2), ("press_n", 2), ("press_q", 2),
    ("last_press_started", 1), ("last_press_amount", 1),
    ("step", 1),
    ("bale_n", 5), ("bale_last_size", 5), ("bale_last_q", 5), ("bale_sum", 5),
]

FLOAT_RTOL = 1e-5     # BASELINE.json north_star: floats within 1e-5 relative
FLOAT_ATOL = 1e-7     # absolute floor for values that are exactly 0 in the reference


def state_rows(state_struct: np.ndarray) -> np.ndarray:
    """msort_env_state_t[N] structured array → int64 [N, W] rows in fixture layout."""
    cols = []
    for name, w in STATE_FIELDS:
        v = np.asarray(state_struct[name]).reshape(len(state_struct), w)
        cols.append(v.astype(np.int64))
    return np.concatenate(cols, axis=1)


def stack_recordings(recs):
    """Stack R single-env recordings (same T) into batched replay inputs/outputs [T, R, ...]."""
    T = len(recs[0]["action"])
    L = max(1, max(len(r["redis_u"]) for r in recs))
    redis = np.zeros((len(recs), L), dtype=np.float64)
    for i, r in enumerate(recs):
        redis[i, :len(r["redis_u"])] = r["redis_u"]
    out = {"redis_u": redis, "T": T}
    for k in ("action", "noise_u", "input_counts", "press_choice", "sort_mode", "obs", "reward",
              "terminated", "overflow", "overflow_material", "mask", "state", "acc_belt",
              "n_draws", "mlp_margin"):
        out[k] = np.stack([r[k] for r in recs], axis=1)
    out["first_pattern0"] = np.asarray([r["first_pattern0"] for r in recs], dtype=np.uint8)
    out["obs0"] = np.stack([r["obs0"] for r in recs], axis=0)
    return out


def pack_counts(counts) -> np.ndarray:
    c = np.asarray(counts, dtype=np.uint32)
    return (c[..., 0] | (c[..., 1] << 8) | (c[..., 2] << 16) | (c[..., 3] << 24)).astype(np.uint32)


def config_dict_for(meta: dict):
    """The config.yml dict a fixture group was recorded under (None = the reference's shipped config.yml)."""
    if not meta.get("thresholds"):
        return None
    from marl_sortingenv_b200.config import load_config
    cfg = load_config(None)
    cfg["pressing_station"]["bale_quality_thresholds"] = dict(zip("ABCD", [float(x) for x in meta["thresholds"]]))
    return cfg


def config_for(meta: dict, num_envs: int, rng_mode="replay", **over) -> _abi.MsortConfig:
    if "config" not in over and config_dict_for(meta) is not None:
        over["config"] = config_dict_for(meta)
    kw = dict(max_steps=meta["max_steps"], seed=0, noise_sorting=meta["noise"],
              balesize=meta["balesize"], use_action_masking=meta["use_action_masking"],
              check_overflow=meta["check_overflow"], auto_reset=meta["auto_reset"],
              rng_mode=rng_mode, sort_policy_mlp=bool(meta.get("mlp", False)))
    kw.update(over)
    return make_config(meta["kind"], num_envs, **kw)


def assert_float_close(got, want, what, rtol=FLOAT_RTOL, atol=FLOAT_ATOL):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    err = np.abs(got - want)
    tol = atol + rtol * np.abs(want)
    if not np.all(err <= tol):
        idx = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{what}: max violation at {idx}: got {got[idx]!r} want {want[idx]!r}")


def replay_and_compare(backend, batch: dict, meta: dict, *, use_input_counts=True,
                       use_sort_mode=None, reward_rtol=FLOAT_RTOL, exact_obs=False,
                       check_mlp=False):
    """Drive `backend` through the recorded trajectory; assert parity at every step.

    Integer state, masks, terminated, overflow flags: bit-exact.  Obs / reward: `rtol`.
    Returns the number of env-steps compared."""
    T = batch["T"]
    n = batch["action"].shape[1]
    kind = meta["kind"]
    obs0, mask0 = backend.reset(first_pattern=batch["first_pattern0"])
    assert_float_close(obs0, batch["obs0"], "reset obs")
    if kind != "sort":
        assert mask0[:, 0].all() and not mask0[:, 1:11].any(), "reset mask: only the no-op is valid"
    if use_sort_mode is None:
        use_sort_mode = kind == "press" and not check_mlp and bool(meta.get("mlp", False))
    for t in range(T):
        kw = dict(noise_u=batch["noise_u"][t], redis_u=batch["redis_u"])
        if use_input_counts:
            kw["input_counts"] = pack_counts(batch["input_counts"][t])
        if kind == "sort":
            kw["press_choice"] = batch["press_choice"][t]
        if use_sort_mode:
            kw["sort_mode"] = batch["sort_mode"][t]
        obs, rew, term, mask, info = backend.step(batch["action"][t], **kw)
        st = state_rows(backend.export_state() if hasattr(backend, "export_state") else backend.state)
        ok = np.ones(n, dtype=bool)
        if check_mlp:
            # the embedded policy may legitimately flip its argmax only on a numerical tie
            flipped = info["sort_mode"].astype(np.int64) != batch["sort_mode"][t]
            assert np.all(batch["mlp_margin"][t][flipped] < 1e-4), \
                f"step {t}: sort-policy argmax differs away from a tie"
            if flipped.any():
                raise AssertionError("MLP tie flip encountered; pick fixture weights with wider margins")
        want_state = batch["state"][t]
        # after an auto-reset the recorded state is the post-step PRE-reset state of the reference
        # (the reference env itself does not auto-reset; ref_record resets at the next step), so
        # compare terminal rows through the info outputs instead.
        done_rows = batch["terminated"][t] & bool(meta["auto_reset"])
        live = ~done_rows
        if not np.array_equal(st[live], want_state[live]):
            bad = np.argwhere(st[live] != want_state[live])[0]
            raise AssertionError(f"step {t}: integer state mismatch at live-row {bad[0]} col {bad[1]}: "
                                 f"got {st[live][bad[0]]} want {want_state[live][bad[0]]}")
        assert np.array_equal(term, batch["terminated"][t]), f"step {t}: terminated mismatch"
        assert np.array_equal(info["overflow"].astype(bool), batch["overflow"][t]), f"step {t}: overflow"
        assert np.array_equal(info["overflow_material"].astype(np.int64)[batch["overflow"][t]],
                              batch["overflow_material"][t][batch["overflow"][t]]), f"step {t}: overflow mat"
        if kind == "sort":
            assert np.array_equal(info["press_action"], batch["press_choice"][t]), f"step {t}: press"
        assert np.array_equal(info["sort_mode"].astype(np.int64)[ok], batch["sort_mode"][t][ok]), \
            f"step {t}: sort mode"
        assert np.array_equal(mask[live], batch["mask"][t][live]), f"step {t}: mask mismatch"
        assert_float_close(rew, batch["reward"][t], f"step {t}: reward", rtol=reward_rtol)
        if done_rows.any():
            assert_float_close(info["terminal_obs"][done_rows], batch["obs"][t][done_rows],
                               f"step {t}: terminal obs")
            assert_float_close(obs[done_rows], batch["obs0"][done_rows], f"step {t}: reset obs after done")
            assert np.array_equal(info["episode_length"][done_rows],
                                  want_state[done_rows][:, _col("step")]), f"step {t}: episode length"
        if exact_obs:
            assert np.array_equal(obs[live], batch["obs"][t][live]), f"step {t}: obs not bit-identical"
        else:
            assert_float_close(obs[live], batch["obs"][t][live], f"step {t}: obs")
    return T * n


def _col(name):
    o = 0
    for nme, w in STATE_FIELDS:
        if nme == name:
            return o
        o += w
    raise KeyError(name)


# --------------------------------------------------------------------------- fixtures
import json as _json
import os as _os

_GOLDEN_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")
GOLDEN_PATH = _os.path.join(_GOLDEN_DIR, "reference_trajectories.npz")
# round-2 additions (tests/golden/make_golden.py --set r02): non-whole-percent quality thresholds, a 16-seed grid,
# 200-step episodes without masking
GOLDEN_PATHS = [GOLDEN_PATH, _os.path.join(_GOLDEN_DIR, "reference_trajectories_r02.npz")]


class _Golden(dict):
    @property
    def files(self):
        return list(self.keys())


_golden_cache = None


def golden():
    global _golden_cache
    if _golden_cache is None:
        g, names = _Golden(), []
        for p in GOLDEN_PATHS:
            if not _os.path.isfile(p):
                continue
            d = np.load(p, allow_pickle=False)
            names += _json.loads(str(d["groups"]))
            for k in d.files:
                if k not in ("groups", "numpy_version"):
                    g[k] = d[k]
        g["groups"] = np.asarray(_json.dumps(names))
        _golden_cache = g
    return _golden_cache


def golden_group_names():
    return _json.loads(str(golden()["groups"]))


def golden_group(name):
    """→ (meta dict, batch dict [T,R,...]) of one fixture group, widened to working dtypes."""
    d = golden()
    meta = _json.loads(str(d[f"{name}/meta"]))
    wide = {"action": np.int64, "input_counts": np.int64, "press_choice": np.uint8,
            "sort_mode": np.int64, "overflow_material": np.int64, "state": np.int64,
            "n_draws": np.int64}
    batch = {}
    for k in d.files:
        if k.startswith(name + "/") and not k.endswith("/meta"):
            kk = k[len(name) + 1:]
            v = np.asarray(d[k])
            batch[kk] = v.astype(wide[kk]) if kk in wide else v
    batch["T"] = meta["steps"]
    return meta, batch


# ---------------------------------------------------------------- Env_3 mode='model' fixture and stand-in agents
def model_mode_golden():
    """tests/golden/reference_model_mode.npz (tests/golden/make_model_mode_golden.py): the reference's
    Env_3.step(mode='model') with deterministic stand-in agents -> (meta, batch [T,R,...])."""
    d = np.load(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "reference_model_mode.npz"),
                allow_pickle=False)
    meta = _json.loads(str(d["model_mono/meta"]))
    batch = {k[len("model_mono/"):]: np.asarray(d[k]) for k in d.files if k.startswith("model_mono/") and not k.endswith("/meta")}
    for k in ("action", "input_counts", "sort_mode", "state", "n_draws"):
        batch[k] = batch[k].astype(np.int64)
    batch["T"] = meta["steps"]
    return meta, batch


class TorchSortStub:
    """Batched twin of oracle/ref_record.py SortStubAgent (same rule, torch tensors)."""

    def predict(self, obs, deterministic=True, **_):
        import torch
        return torch.where(obs[:, 1] + obs[:, 3] > obs[:, 2] + obs[:, 4], 0, 1), None


class TorchPressStub:
    """Batched twin of oracle/ref_record.py MaskablePressStubAgent."""

    def predict(self, obs, deterministic=True, action_masks=None, **_):
        import torch
        n = obs.shape[0]
        m = torch.ones((n, 11), dtype=torch.bool, device=obs.device) if action_masks is None else action_masks
        level = obs[:, :5].repeat(1, 2)                                   # level of action a's container, a = 1..10
        score = torch.where(m[:, 1:], level, torch.full_like(level, -1.0))
        best = score.argmax(dim=1) + 1                                    # first maximum = lowest action on ties
        return torch.where(m[:, 1:].any(dim=1), best, torch.zeros_like(best)), None


# ---------------------------------------------------------------- the reference on the production generator's random inputs
def philox_golden():
    """tests/golden/reference_philox.npz (tests/golden/make_philox_golden.py): outputs of the UNMODIFIED reference driven by
    the random inputs the Philox generator produces for (seed, global env id, episode, step) -> {name: (meta, batch)}."""
    d = np.load(_os.path.join(_GOLDEN_DIR, "reference_philox.npz"), allow_pickle=False)
    out = {}
    for name in _json.loads(str(d["groups"])):
        meta = _json.loads(str(d[f"{name}/meta"]))
        batch = {k[len(name) + 1:]: np.asarray(d[k]) for k in d.files if k.startswith(name + "/") and not k.endswith("/meta")}
        out[name] = (meta, batch)
    return out


def compare_with_philox_reference(backend, meta, batch):
    """Step `backend` (PHILOX mode, same seed / global env ids) with the recorded actions and require the reference's
    outputs: integer state, masks, flags bit-exact on live envs; obs / reward within the float tolerance; terminal
    observation and episode length where an episode ends (the backend auto-resets, the reference row is pre-reset)."""
    T = meta["steps"]
    obs0, _ = backend.reset()
    assert_float_close(obs0, batch["obs0"], "reset obs")
    for t in range(T):
        obs, rew, term, mask, info = backend.step(batch["action"][t].astype(np.int64))
        want_term = batch["terminated"][t]
        assert np.array_equal(term, want_term), f"step {t}: terminated"
        live = ~want_term
        st = state_rows(backend.export_state() if hasattr(backend, "export_state") else backend.state)
        if not np.array_equal(st[live], batch["state"][t][live]):
            bad = np.argwhere(st[live] != batch["state"][t][live])[0]
            raise AssertionError(f"step {t}: integer state differs from the reference at live-row {bad[0]} col {bad[1]}: "
                                 f"got {st[live][bad[0]]} want {batch['state'][t][live][bad[0]]}")
        assert np.array_equal(mask[live], batch["mask"][t][live]), f"step {t}: mask"
        assert_float_close(rew, batch["reward"][t], f"step {t}: reward")
        assert_float_close(obs[live], batch["obs"][t][live], f"step {t}: obs")
        if want_term.any():
            assert_float_close(info["terminal_obs"][want_term], batch["obs"][t][want_term], f"step {t}: terminal obs")
            assert np.array_equal(info["episode_length"][want_term], batch["state"][t][want_term][:, _col("step")]), f"step {t}: episode length"
        if "overflow" in info:
            assert np.array_equal(np.asarray(info["overflow"]).astype(bool), batch["overflow"][t]), f"step {t}: overflow"
    return T * batch["action"].shape[1]
