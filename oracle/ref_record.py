"""TEST INFRASTRUCTURE — records trajectories (inputs, RNG streams, outputs, integer
state) from the UNMODIFIED reference envs.  Used by tests/golden/make_golden.py to
produce the committed fixtures and by tests that run only where /root/reference
exists.  Nothing in the product package imports this file.

Recorded per step:
  * replay inputs  : external action, accuracy-noise uniforms (twin of
    `rng_noise = default_rng(seed+4)`, env_super.py:173,508), redistribution uniforms
    (twin of `rng = default_rng(seed+99)`, env_super.py:174,563 — one double per
    `choice` call), the input-batch counts the generator emitted
    (env_super.py:445-453), Env_1's internally sampled press action
    (env_super.py:291-300, read back from `press_actions_per_timestep`), and the sort
    mode that was applied (Env_2 with an embedded sort agent, env_2_press.py:106-112).
  * outputs        : obs (f32), reward (f64), terminated, info flags, action mask after
    the step, and the full integer plant state.
"""
from __future__ import annotations

import numpy as np

from .ref_loader import make_reference_env

MATS = ["A", "B", "C", "D"]
ALL5 = MATS + ["E"]
KIND_DIMS = {"sort": (13, 2), "press": (16, 11), "mono": (29, 22)}

STATE_FIELDS = [
    # name, width
    ("input", 4), ("belt", 4), ("sorting", 4),
    ("cont_true", 4), ("cont_false", 4), ("cont_e", 1),
    ("press_timer", 2), ("press_mat", 2), ("press_n", 2), ("press_q", 2),
    ("last_press_started", 1), ("last_press_amount", 1),
    ("step", 1),
    ("bale_n", 5), ("bale_last_size", 5), ("bale_last_q", 5), ("bale_sum", 5),
]
STATE_WIDTH = sum(w for _, w in STATE_FIELDS)


class _CountingRng:
    """Delegating proxy around `env.rng` that counts `choice` calls (env_super.py:563)."""

    def __init__(self, gen):
        self._gen = gen
        self.n_choice = 0

    def choice(self, *a, **k):
        self.n_choice += 1
        return self._gen.choice(*a, **k)

    def __getattr__(self, name):
        return getattr(self._gen, name)


class NumpyMlpSortAgent:
    """Duck-typed stand-in for an SB3 `MlpPolicy` sort agent (training.py:115 arch
    pi=[32,32], SB3 default tanh): Flatten → Linear(13,32)·tanh → Linear(32,32)·tanh →
    Linear(32,2) → argmax.  fp32 throughout.  `weights` is the flat 1570-float vector
    [W1(32x13) b1(32) W2(32x32) b2(32) W3(2x32) b3(2)], rows = output units."""

    def __init__(self, weights):
        w = np.asarray(weights, dtype=np.float32)
        assert w.size == 1570
        o = 0
        self.W1 = w[o:o + 416].reshape(32, 13); o += 416
        self.b1 = w[o:o + 32]; o += 32
        self.W2 = w[o:o + 1024].reshape(32, 32); o += 1024
        self.b2 = w[o:o + 32]; o += 32
        self.W3 = w[o:o + 64].reshape(2, 32); o += 64
        self.b3 = w[o:o + 2]
        self.last_logits = None

    def logits(self, obs):
        x = np.asarray(obs, dtype=np.float32)
        h = np.tanh(self.W1 @ x + self.b1).astype(np.float32)
        h = np.tanh(self.W2 @ h + self.b2).astype(np.float32)
        return (self.W3 @ h + self.b3).astype(np.float32)

    def predict(self, obs, deterministic=True, **_):
        lg = self.logits(obs)
        self.last_logits = lg
        return np.int64(np.argmax(lg)), None


def sb3_style_mlp_weights(seed: int, action_gain: float = 0.01) -> np.ndarray:
    """Orthogonal init with SB3's gains (sqrt2, sqrt2, action_gain), zero bias."""
    rng = np.random.default_rng(seed)

    def ortho(rows, cols, gain):
        a = rng.standard_normal((max(rows, cols), min(rows, cols)))
        q, r = np.linalg.qr(a)
        q = q * np.sign(np.diag(r))
        if rows < cols:
            q = q.T
        return (gain * q[:rows, :cols]).astype(np.float32)

    parts = [ortho(32, 13, np.sqrt(2)).ravel(), np.zeros(32, np.float32),
             ortho(32, 32, np.sqrt(2)).ravel(), np.zeros(32, np.float32),
             ortho(2, 32, action_gain).ravel(), np.zeros(2, np.float32)]
    return np.concatenate(parts).astype(np.float32)


def snapshot(env) -> np.ndarray:
    """Integer plant state of a reference env as one int64 row (layout STATE_FIELDS)."""
    cm, ps = env.container_materials, env.press_state
    row = []
    row += [int(x) for x in env.current_material_input]
    row += [int(x) for x in env.current_material_belt]
    row += [int(x) for x in env.current_material_sorting]
    row += [int(cm[m]) for m in MATS]
    row += [int(cm[m + "_False"]) for m in MATS]
    row += [int(cm["E"])]
    row += [int(ps["press_1"]), int(ps["press_2"])]
    for i in (1, 2):
        m = ps[f"material_{i}"]
        row.append(ALL5.index(m) if isinstance(m, str) else 0)
    row += [int(ps["n_1"]), int(ps["n_2"])]
    row += [int(np.rint(float(ps["q_1"]) * 100)), int(np.rint(float(ps["q_2"]) * 100))]
    row += [int(bool(env._last_press_started)), int(env._last_press_amount)]
    row += [int(env.current_step)]
    for f in (lambda b: len(b), lambda b: b[-1][0] if b else 0,
              lambda b: b[-1][1] if b else 0, lambda b: sum(x[0] for x in b)):
        row += [int(f(env.bale_count[m])) for m in ALL5]
    assert len(row) == STATE_WIDTH
    return np.asarray(row, dtype=np.int64)


def _press_log_to_discrete(entry) -> int:
    p, m = entry
    if p in (0, None):
        return 0
    assert p in (1, 2), f"unexpected press log entry {entry}"
    return (p - 1) * 5 + int(m) + 1


class SortStubAgent:
    """Deterministic stand-in for a trained sort agent in Env_3.step(mode='model') (env_monolith.py:189-193):
    remembers the observation the reference hands it and answers with a fixed rule on it (belt share of
    A+C against B+D, observation entries 1..4)."""

    def __init__(self):
        self.seen = None

    def predict(self, obs, deterministic=True, **_):
        x = np.asarray(obs, dtype=np.float32)
        self.seen = x.copy()
        return (0 if float(x[1]) + float(x[3]) > float(x[2]) + float(x[4]) else 1), None


class MaskablePressStubAgent:
    """Deterministic stand-in for a MaskablePPO press agent (the reference passes `action_masks` when the
    type name contains 'Maskable' and the object has `.policy`, env_monolith.py:199-206): the valid press
    action whose container is fullest by the observation (entries 0..4; ties -> lowest action), no-op when
    only the no-op is valid."""
    policy = object()

    def __init__(self):
        self.seen = None
        self.seen_mask = None

    def predict(self, obs, deterministic=True, action_masks=None, **_):
        x = np.asarray(obs, dtype=np.float32)
        self.seen = x.copy()
        m = np.ones(11, dtype=bool) if action_masks is None else np.asarray(action_masks, dtype=bool)
        self.seen_mask = m.copy()
        best, best_level = 0, -1.0
        for a in range(1, 11):
            if m[a] and float(x[(a - 1) % 5]) > best_level:
                best, best_level = a, float(x[(a - 1) % 5])
        return best, None


def record(kind: str, *, seed: int, steps: int, max_steps: int = 50, noise: float = 0.05,
           balesize: int = 200, policy="masked_random", action_seed: int = 0,
           use_action_masking: bool = True, check_overflow: bool = False,
           auto_reset: bool = True, mlp_weights=None, actions=None, keep_env: bool = False,
           thresholds=None) -> dict:
    """Run the reference for `steps` env-steps (auto-resetting unseeded like SB3's VecEnv when
    an episode ends, if `auto_reset`) and return everything needed to replay and compare."""
    D, A = KIND_DIMS[kind]
    env = make_reference_env(kind, max_steps=max_steps, seed=seed, noise_sorting=noise,
                             balesize=balesize)
    if thresholds is not None:
        # a user config.yml's pressing_station.bale_quality_thresholds (env_super.py:98): plain Python floats, set on
        # the instance exactly as Env_Super.__init__ leaves them — the reference source and its config.yml stay untouched
        env.quality_thresholds = {m: float(v) for m, v in zip(MATS, thresholds)}
    agent = None
    if kind == "press" and mlp_weights is not None:
        agent = NumpyMlpSortAgent(mlp_weights)
        env.set_agents(sort_agent=agent)
    stubs = None
    if policy == "mode_model":          # Env_3.step(action=None, mode='model') with both agents assigned
        stubs = (SortStubAgent(), MaskablePressStubAgent())
        env.set_agents(sort_agent=stubs[0], press_agent=stubs[1])
    obs0, _ = env.reset(seed=seed)
    env.rng = _CountingRng(env.rng)
    twin_noise = np.random.default_rng(seed + 4)
    twin_redis = np.random.default_rng(seed + 99)
    arng = np.random.default_rng(action_seed)

    out = {k: [] for k in ("action", "noise_u", "n_draws", "input_counts", "press_choice",
                           "sort_mode", "obs", "reward", "terminated", "overflow",
                           "overflow_material", "mask", "state", "acc_belt", "reset_before",
                           "first_pattern", "mlp_margin", "agent_obs")}
    first_pattern0 = int(env.input_generator.pattern_sequence[0])
    reset_next = False
    for t in range(steps):
        did_reset = False
        if reset_next and auto_reset:
            env.reset()            # unseeded, like SB3's VecEnv (streams run on; env_super.py:377)
            did_reset = True
        reset_next = False
        mask = np.asarray(env.action_masks(), dtype=bool)
        if actions is not None:
            a = int(actions[t])
        elif policy == "masked_random":
            v = np.flatnonzero(mask)
            a = int(v[arng.integers(0, v.size)])
        elif policy == "uniform":
            a = int(arng.integers(0, A))
        elif policy == "rule":              # config 1: Env_1 with sorting_rules() before the step
            a = int(env.sorting_rules())
        elif policy == "first_valid":       # Appendix C, Env_2 row
            v = np.flatnonzero(mask)
            a = int(v[1]) if v.size > 1 else 0
        elif policy == "second_valid":      # Appendix C, Env_3 row
            v = np.flatnonzero(mask)
            a = int(v[1]) if v.size > 1 else 0
        elif policy in ("mode_rule_based", "mode_model"):   # Env_3.step(action=None, mode=...): chosen inside step()
            a = None
        else:
            raise ValueError(policy)
        n_before = env.rng.n_choice
        kw = dict(use_action_masking=use_action_masking, check_overflow=check_overflow)
        if a is None:
            obs, reward, term, trunc, info = env.step(None, mode="model" if stubs else "rule_based", **kw)
            a = int(info["action"])
            if stubs:                   # what the two agents were shown inside step(): sort obs (13) | press obs (16)
                out["agent_obs"].append(np.concatenate([stubs[0].seen, stubs[1].seen]))
        else:
            obs, reward, term, trunc, info = env.step(a, **kw)
        assert trunc is False
        out["action"].append(a)
        out["noise_u"].append(twin_noise.random(4))
        out["n_draws"].append(env.rng.n_choice - n_before)
        out["input_counts"].append([int(x) for x in env.current_material_input])
        out["press_choice"].append(_press_log_to_discrete(env.press_actions_per_timestep[-1])
                                   if kind == "sort" else 0)
        out["sort_mode"].append(int(env.sensor_current_setting))
        out["obs"].append(np.asarray(obs, dtype=np.float32))
        out["reward"].append(float(reward))
        out["terminated"].append(bool(term))
        out["overflow"].append(bool(info.get("overflow", False)))
        om = info.get("overflow_material", None)
        out["overflow_material"].append(ALL5.index(om) if om is not None else -1)
        out["mask"].append(np.asarray(env.action_masks(), dtype=bool))
        out["state"].append(snapshot(env))
        out["acc_belt"].append(np.asarray(env.accuracy_belt, dtype=np.float64))
        out["reset_before"].append(did_reset)
        out["first_pattern"].append(int(env.input_generator.pattern_sequence[0]))
        if agent is not None and agent.last_logits is not None:
            out["mlp_margin"].append(float(abs(agent.last_logits[0] - agent.last_logits[1])))
        else:
            out["mlp_margin"].append(np.inf)
        if term:
            reset_next = True

    total_draws = int(np.sum(out["n_draws"]))
    res = {
        "obs0": np.asarray(obs0, dtype=np.float32),
        "first_pattern0": np.int64(first_pattern0),
        "action": np.asarray(out["action"], dtype=np.int64),
        "noise_u": np.asarray(out["noise_u"], dtype=np.float64).reshape(steps, 4),
        "n_draws": np.asarray(out["n_draws"], dtype=np.int64),
        "redis_u": twin_redis.random(total_draws) if total_draws else np.zeros(0),
        "input_counts": np.asarray(out["input_counts"], dtype=np.int64).reshape(steps, 4),
        "press_choice": np.asarray(out["press_choice"], dtype=np.int64),
        "sort_mode": np.asarray(out["sort_mode"], dtype=np.int64),
        "obs": np.asarray(out["obs"], dtype=np.float32).reshape(steps, D),
        "reward": np.asarray(out["reward"], dtype=np.float64),
        "terminated": np.asarray(out["terminated"], dtype=bool),
        "overflow": np.asarray(out["overflow"], dtype=bool),
        "overflow_material": np.asarray(out["overflow_material"], dtype=np.int64),
        "mask": np.asarray(out["mask"], dtype=bool).reshape(steps, A),
        "state": np.asarray(out["state"], dtype=np.int64).reshape(steps, STATE_WIDTH),
        "acc_belt": np.asarray(out["acc_belt"], dtype=np.float64).reshape(steps, 4),
        "reset_before": np.asarray(out["reset_before"], dtype=bool),
        "first_pattern": np.asarray(out["first_pattern"], dtype=np.int64),
        "mlp_margin": np.asarray(out["mlp_margin"], dtype=np.float64),
    }
    if stubs:
        res["agent_obs"] = np.asarray(out["agent_obs"], dtype=np.float32).reshape(steps, D)
    if keep_env:
        res["env"] = env          # the reference env itself (its Python logs; tests/golden/make_log_golden.py)
    return res
