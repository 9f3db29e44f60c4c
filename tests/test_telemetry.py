"""Telemetry: the logs rebuilt from per-step snapshots must equal the Python logs
the UNMODIFIED reference kept for the same episode (tests/golden/reference_logs.npz, written by
tests/golden/make_log_golden.py): reward_data, press_actions_per_timestep, bale_count.

CPU: snapshots come from the C oracle driven in REPLAY mode.  GPU: from the device TraceRecorder."""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

from marl_sortingenv_b200.telemetry import ALL5, logs_from_arrays, reference_view
from parity_util import config_for, pack_counts

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_logs.npz")
DATA = np.load(FIX, allow_pickle=False)
NAMES = json.loads(str(DATA["groups"]))
NAME_BASE = 100


def _group(name):
    meta = json.loads(str(DATA[f"{name}/meta"]))
    g = {k.split("/", 1)[1]: DATA[k] for k in DATA.files if k.startswith(name + "/") and not k.endswith("/meta")}
    return meta, g


def _step_kwargs(meta, g, t):
    kw = dict(noise_u=g["noise_u"][t][None], redis_u=g["redis_u"][None] if g["redis_u"].size else np.zeros((1, 1)),
              input_counts=pack_counts(g["input_counts"][t])[None])
    if meta["kind"] == "sort":
        kw["press_choice"] = g["press_choice"][t][None]
    return kw


def _decode_press_log(arr):
    out = []
    for code, m in arr.tolist():
        out.append((code, None if m < 0 else (ALL5[m - NAME_BASE] if m >= NAME_BASE else m)))
    return out


def _check_logs(logs, meta, g):
    T = meta["T"]
    rd = logs["reward_data"]
    assert len(rd["Reward"]) == T
    assert np.array_equal(np.asarray(rd["Accuracy"], dtype=np.float64), g["log_accuracy"]), "Accuracy (mean purity)"
    got = np.asarray(rd["Reward"], dtype=np.float64)
    assert np.all(np.abs(got - g["log_reward"]) <= 1e-7 + 1e-5 * np.abs(g["log_reward"])), "Reward terms"
    assert np.allclose(rd["Total"], g["log_total"], rtol=1e-5, atol=2e-7)
    assert np.array_equal(np.asarray(rd["Setting"]), g["log_setting"].astype(np.int64))
    assert np.array_equal(np.asarray(rd["Belt_Occupancy"], dtype=np.float64), g["log_belt_occ"])
    bp = np.asarray([[d[m] for m in ALL5[:4]] for d in rd["Belt_Proportions"]], dtype=np.float64)
    assert np.array_equal(bp, g["log_belt_prop"])          # same float64 division
    for mi, m in enumerate(ALL5):
        assert np.array_equal(np.asarray(rd[f"{m}_True"]), g["log_true"][:, mi]), f"{m}_True"
        assert np.array_equal(np.asarray(rd[f"{m}_False"]), g["log_false"][:, mi]), f"{m}_False"
    assert logs["press_actions_per_timestep"] == _decode_press_log(g["press_log"])
    want = {m: [] for m in ALL5}
    for mi, s, q in zip(g["bale_mat"].tolist(), g["bale_size"].tolist(), g["bale_q"].tolist()):
        want[ALL5[mi]].append((s, q))
    assert logs["bale_count"] == want


@pytest.mark.parametrize("name", NAMES)
def test_logs_from_oracle_snapshots_match_reference_logs(name):
    from oracle.cpu_oracle import OracleEnv
    meta, g = _group(name)
    cfg = config_for(meta, 1)
    env = OracleEnv(cfg, nthreads=1)
    env.reset(first_pattern=np.asarray([int(g["first_pattern0"])], dtype=np.uint8))
    snaps = [env.state.copy()[0]]
    info = {k: [] for k in ("action", "sort_mode", "press_action", "invalid_action", "reward_sort", "reward_press", "sorted_true")}
    for t in range(meta["T"]):
        obs, rew, term, mask, inf = env.step(g["action"][t][None], **_step_kwargs(meta, g, t))
        assert abs(rew[0] - g["reward"][t]) <= 1e-9 + 1e-9 * abs(g["reward"][t])
        snaps.append(env.state.copy()[0])
        for k in info:
            info[k].append(inf[k][0])
    snaps = np.asarray(snaps, dtype=env.state.dtype)
    logs = logs_from_arrays(meta["kind"], snaps, {k: np.asarray(v) for k, v in info.items()},
                            bale_size=int(cfg.bale_size), bale_remainder_threshold=float(cfg.bale_remainder_threshold))
    _check_logs(logs, meta, g)
    view = reference_view(meta["kind"], snaps, {k: np.asarray(v) for k, v in info.items()}, cfg, seed=meta["seed"])
    # the attributes utils/plotting.plot_env unpacks (plotting.py:32-48)
    for attr in ("current_material_input", "current_material_belt", "current_material_sorting", "container_materials",
                 "accuracy_belt", "accuracy_sorter", "sensor_current_setting", "reward_data", "belt_occupancy",
                 "press_state", "bale_count", "bale_standard_size", "quality_thresholds",
                 "press_actions_per_timestep", "container_global_max", "press_times", "seed"):
        assert hasattr(view, attr), attr
    last = g["state"][meta["T"] - 1]
    assert [view.container_materials[m] for m in "ABCD"] == last[12:16].tolist()
    assert view.container_materials["E"] == int(last[20])


def test_recorder_requires_episode_start():
    meta, g = _group(NAMES[0])
    from oracle.cpu_oracle import OracleEnv
    env = OracleEnv(config_for(meta, 1), nthreads=1)
    env.reset(first_pattern=np.asarray([int(g["first_pattern0"])], dtype=np.uint8))
    s = env.state.copy()
    s["bale_n"][0, 1] = 3
    snaps = np.asarray([s[0], s[0]], dtype=s.dtype)
    info = {k: np.zeros(1) for k in ("action", "sort_mode", "press_action", "invalid_action", "reward_sort", "reward_press", "sorted_true")}
    with pytest.raises(ValueError):
        logs_from_arrays(meta["kind"], snaps, info, bale_size=200, bale_remainder_threshold=0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_trace_recorder_matches_reference_logs(name):
    """The device path: BatchedEnv (REPLAY mode, info_level='full') with a TraceRecorder attached; the
    traced env sits among untraced ones."""
    import torch
    from marl_sortingenv_b200.batched import ENV_CLASSES
    meta, g = _group(name)
    n, j = 5, 3                                   # 5 identical envs, trace number 3 (and 0)
    env = ENV_CLASSES[meta["kind"]](n, max_steps=meta["max_steps"], seed=0, noise_sorting=meta["noise"],
                                    balesize=meta["balesize"], use_action_masking=meta["use_action_masking"],
                                    check_overflow=meta["check_overflow"], auto_reset=False, rng_mode="replay",
                                    info_level="full")
    rec = env.attach_trace([j, 0], capacity=meta["T"])
    env.reset(first_pattern=np.full(n, int(g["first_pattern0"]), dtype=np.uint8))
    redis = torch.as_tensor(np.tile(g["redis_u"] if g["redis_u"].size else np.zeros(1), (n, 1))).cuda().contiguous()
    for t in range(meta["T"]):
        kw = _step_kwargs(meta, g, t)
        rp = dict(noise_u=np.tile(kw["noise_u"], (n, 1)), redis_u=redis, input_counts=np.tile(kw["input_counts"], n))
        if "press_choice" in kw:
            rp["press_choice"] = np.tile(kw["press_choice"], n)
        env.step(torch.full((n,), int(g["action"][t]), dtype=torch.int64, device="cuda"), replay=rp)
    env.sync_check()
    assert rec.t == meta["T"]
    for slot in (0, 1):
        _check_logs(rec.reference_logs(slot), meta, g)
    view = rec.reference_view(0)
    assert view.current_step == meta["T"] and view.bale_standard_size == meta["balesize"]
    with pytest.raises(RuntimeError):
        env.step(torch.zeros(n, dtype=torch.int64, device="cuda"), replay=rp)   # capacity exhausted
