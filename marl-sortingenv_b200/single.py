"""Single-environment classes with the reference's exact constructor / reset / step signatures.

`Env_1_Sorting`, `Env_2_Pressing`, `Env_3_Monolith` here are N=1 views of the batched device
simulator (ref: env_1_sort.py:12-154, env_2_press.py:12-165, env_monolith.py:12-284): numpy
observations, Python float rewards, bool flags and an `info` dict with the reference's keys.
They exist so that code written against the reference (testing.py's episode loop, SB3's
`check_env`, ActionMasker) runs unchanged; for throughput use the batched classes.

The reference's Python-side logs — `reward_data`, `press_actions_per_timestep`, `bale_count` (the
lists `testing.test_env` sums and `utils/plotting.plot_env` renders; env_super.py:928-946, 631-637,
661-687) — are available under the same names: a telemetry recorder (telemetry.py) snapshots the env
on the device after every step and the lists are rebuilt from the snapshots on access.

Inherent difference (documented in DESIGN.md): the random plant is the counter-based Philox generator
(same distributions, different bits than numpy's PCG64 streams).
"""
from __future__ import annotations

import numpy as np
import torch

from .batched import ENV_CLASSES

MATERIALS = ["A", "B", "C", "D"]
ALL5 = MATERIALS + ["E"]


class _SingleEnv:
    kind = "mono"

    def __init__(self, max_steps: int = 50, seed: int = None, noise_sorting: float = 0.05,
                 balesize: int = 200, simulation=False, config_path=None, device="cuda:0"):
        self._b = ENV_CLASSES[self.kind](1, device=device, max_steps=max_steps, seed=seed,
                                         noise_sorting=noise_sorting, balesize=balesize,
                                         config_path=config_path, auto_reset=False, info_level="full")
        self.name = self.kind
        self.max_steps = max_steps
        self.seed = seed
        self.material_names = list(MATERIALS)
        self.observation_space = self._b.observation_space
        self.action_space = self._b.action_space
        self.sort_agent = None
        self.press_agent = None
        self.mono_agent = None
        self.bale_standard_size = int(self._b.cfg.bale_size)
        self.container_global_max = int(self._b.cfg.container_capacity)
        self._act = torch.zeros(1, dtype=torch.int64, device=self._b.device)
        self._masking, self._overflow = True, False
        self._state_cache = None          # exported plain state of the current step (one export kernel + copy per step at most)
        self._set_host_streams(seed)
        from .telemetry import TraceRecorder
        self._b._trace = TraceRecorder(self._b, [0], capacity=max(int(max_steps), 1), growable=True)

    # ------------------------------------------------------------------ gym surface
    @property
    def unwrapped(self):
        return self

    def _set_host_streams(self, seed):
        """The host-side random streams of the reference's action fallbacks, seeded like Env_Super.set_seed
        (env_super.py:165-175): rng_sorting = default_rng(seed + 2), rng_pressing = default_rng(seed + 3).  (The plant's
        own draws — noise, redistribution, pattern order — are Philox streams on the device.)"""
        s = int(seed or 0)
        self.rng_sorting = np.random.default_rng(s + 2)
        self.rng_pressing = np.random.default_rng(s + 3)

    def reset(self, seed=None, options=None):
        """ref: Env_X.reset(seed) → (obs, {}) (env_super.py:365-420)."""
        if seed is not None:
            self._set_host_streams(seed)      # reset(seed=s) re-seeds every stream (env_super.py:377-378)
        self._state_cache = None
        obs, _ = self._b.reset(seed=seed)
        return obs[0].cpu().numpy().copy(), {}

    def action_masks(self):
        return self._b.action_masks()[0].cpu().numpy().copy()

    def get_obs(self):
        return self._b.get_obs()[0].cpu().numpy().copy()

    def close(self):
        self._b.close()

    def _set_step_flags(self, use_action_masking, check_overflow):
        if (use_action_masking, check_overflow) != (self._masking, self._overflow):
            self._b.set_flags(use_action_masking=use_action_masking, check_overflow=check_overflow)
            self._masking, self._overflow = use_action_masking, check_overflow

    def _do_step(self, action: int, use_action_masking=True, check_overflow=False):
        self._set_step_flags(bool(use_action_masking), bool(check_overflow))
        self._act[0] = int(action)
        self._state_cache = None
        obs, rew, term, trunc, info = self._b.step(self._act)
        out_obs = obs[0].cpu().numpy().copy()
        reward = float(rew[0].item())
        terminated = bool(term[0].item())
        d = {"action": int(info["action"][0].item())}
        if bool(info["overflow"][0].item()):
            d = {"overflow": True, "overflow_material": ALL5[int(info["overflow_material"][0].item())],
                 "action": d["action"]}
        return out_obs, reward, terminated, False, d

    # ------------------------------------------------------------------ state views (read-only)
    def _state(self):
        if self._state_cache is None:
            self._state_cache = self._b.export_state()[0]
        return self._state_cache

    @property
    def current_step(self):
        return int(self._state()["step"])

    @property
    def container_materials(self):
        s = self._state()
        d = {m: int(s["cont_true"][i]) for i, m in enumerate(MATERIALS)}
        d.update({f"{m}_False": int(s["cont_false"][i]) for i, m in enumerate(MATERIALS)})
        d["E"] = int(s["cont_e"])
        return d

    @property
    def press_state(self):
        s = self._state()
        out = {}
        for i in (1, 2):
            t = int(s["press_timer"][i - 1])
            out[f"press_{i}"] = t
            out[f"material_{i}"] = ALL5[int(s["press_mat"][i - 1])] if t > 0 else 0
            out[f"n_{i}"] = int(s["press_n"][i - 1])
            out[f"q_{i}"] = int(s["press_q"][i - 1]) / 100.0
        return out

    # ---- the reference's per-step logs, rebuilt from the device snapshots of this episode
    def _logs(self):
        if not self._b._trace.started:
            raise AttributeError("reward_data / press_actions_per_timestep / bale_count exist after reset()")  # as in the reference
        return self._b._trace.reference_logs(0)

    @property
    def reward_data(self):
        """ref: env_super.py:402-408, 928-946 — dict of per-step lists ('Reward' = (r_sort, r_press), 'Total',
        'Setting', 'Belt_Occupancy', 'Belt_Proportions', 'Accuracy', '<M>_True', '<M>_False')."""
        return self._logs()["reward_data"]

    @property
    def press_actions_per_timestep(self):
        """ref: env_super.py:631-637, 729-736 — (press, material) per step, (0, None) for the no-op, 111/222 codes."""
        return self._logs()["press_actions_per_timestep"]

    @property
    def bale_count(self):
        """ref: env_super.py:661-687 — per material the list of (size, quality) of every bale pressed this episode."""
        return self._logs()["bale_count"]

    def render(self, mode="human", save=False, show=True, log_dir="./img/log", filename="plot", title="", format="svg",
               checksum=True, steps_test=None):
        """ref: Env_Super.render (env_super.py:229-233) — hands the env to the reference's dashboard
        `utils.plotting.plot_env`.  The dashboard itself is not part of this package (plotting is out of
        scope): it is imported from the reference checkout on `sys.path`, and it receives an object carrying
        exactly the attributes it unpacks (plotting.py:32-48), rebuilt from the device snapshots."""
        try:
            from utils.plotting import plot_env
        except ImportError as e:
            raise RuntimeError("render() uses the reference's own utils/plotting.py (matplotlib, seaborn): put the "
                               "MARL-SortingEnv checkout on sys.path") from e
        return plot_env(env=self._b._trace.reference_view(0), save=save, show=show, log_dir=log_dir, filename=filename,
                        title=title, format=format, checksum=checksum, steps_test=steps_test)

    @property
    def bale_counters(self):
        """Per material: dict(count, last_size, last_quality, total_size) — the counters the device keeps
        instead of the reference's per-bale lists."""
        s = self._state()
        return {m: dict(count=int(s["bale_n"][i]), last_size=int(s["bale_last_size"][i]),
                        last_quality=int(s["bale_last_q"][i]), total_size=int(s["bale_sum"][i]))
                for i, m in enumerate(ALL5)}

    @property
    def current_material_input(self):
        return [int(x) for x in self._state()["input"]]

    @property
    def current_material_belt(self):
        return [int(x) for x in self._state()["belt"]]

    @property
    def current_material_sorting(self):
        return [int(x) for x in self._state()["sorting"]]

    @property
    def accuracy_belt(self):
        return [float(x) for x in self._state()["acc_belt"]]

    @property
    def sensor_current_setting(self):
        return int(self._state()["sensor_mode"])

    # ------------------------------------------------------------------ heuristics (host side, N=1)
    def sorting_rules(self):
        """ref: env_super.py:469-482 (float64 proportions, as the reference)."""
        belt = self.current_material_belt
        tot = sum(belt)
        p = [b / tot if tot > 0 else 0 for b in belt]
        return 0 if p[0] + p[2] > p[1] + p[3] else 1

    def check_container_level(self):
        """ref: env_super.py:689-720 — fullest container on the first free press."""
        ps, cm = self.press_state, self.container_materials
        free = 1 if ps["press_1"] == 0 else (2 if ps["press_2"] == 0 else None)
        if free is None:
            return None, None
        best_idx, best = None, 0
        for i, m in enumerate(MATERIALS):
            lvl = cm[m] + cm[f"{m}_False"]
            if lvl > best:
                best, best_idx = lvl, i
        if cm["E"] > best:
            best, best_idx = cm["E"], 4
        return (free, best_idx) if best > 0 else (None, None)

    @staticmethod
    def press_action_to_discrete(press_id, mat_id):
        return 0 if press_id == 0 else (press_id - 1) * 5 + mat_id + 1

    @staticmethod
    def press_discrete_to_action(action):
        if action == 0:
            return [0, None]
        return [1 if action <= 5 else 2, (action - 1) % 5]


class Env_1_Sorting(_SingleEnv):
    """ref: env_1_sort.py:12-154 — Box(13) / Discrete(2); the press part acts randomly under the mask."""
    kind = "sort"

    def set_agents(self, press_agent=None):
        self.press_agent = press_agent

    def step(self, action=None, use_action_masking=True, check_overflow=False):
        return self._do_step(action, use_action_masking, check_overflow)


class Env_2_Pressing(_SingleEnv):
    """ref: env_2_press.py:12-165 — Box(16) / Discrete(11); sort mode from the embedded agent or sorting_rules()."""
    kind = "press"

    def set_agents(self, sort_agent=None):
        """ref: env_2_press.py:39-40.  The agent's MLP weights are uploaded and evaluated in the kernel."""
        self.sort_agent = sort_agent
        self._b.set_sort_policy(sort_agent)

    def step(self, action, use_action_masking=True, check_overflow=False):
        return self._do_step(action, use_action_masking, check_overflow)


class Env_3_Monolith(_SingleEnv):
    """ref: env_monolith.py:12-284 — Box(29) / Discrete(22); action = 11*sort_mode + press_action."""
    kind = "mono"

    def set_agents(self, sort_agent=None, press_agent=None, mono_agent=None):
        self.sort_agent, self.press_agent, self.mono_agent = sort_agent, press_agent, mono_agent

    def step(self, action=None, mode=None, use_action_masking=True, check_overflow=False):
        """ref: env_monolith.py:109-284.  `action` given → device path.  The other action sources of the
        reference (`mono_agent`, mode='random' | 'rule_based' | 'model') are resolved on the host
        from the current obs / mask / state and then stepped on the device."""
        if action is None:
            action = self._choose_action(mode, use_action_masking)
            # Only an external action and mode='random' without masking are sanitised (env_monolith.py:132-138,
            # 246-257); a mono_agent's, rule-based, modular or masked-random choice goes straight to
            # press_action_rules (:258-262) whatever `use_action_masking` says.
            if not (mode == "random" and self.mono_agent is None):
                return self._do_step(int(action), True, check_overflow)
        return self._do_step(int(action), use_action_masking, check_overflow)

    def _obs_after_shift(self):
        """Observation as the reference's agents see it inside step(): after update_environment has
        shifted input -> belt -> sorting (env_monolith.py:114-115), before anything else changed
        (msort_observe_after_shift)."""
        return self._b.observe_after_shift()[0].cpu().numpy()

    def _choose_action(self, mode, use_action_masking):
        if self.mono_agent is not None:                                   # :144-150
            a, _ = self.mono_agent.predict(self.get_obs(), deterministic=True, action_masks=self.action_masks())
            return int(a)
        if mode == "random":                                              # :152-164
            if use_action_masking:
                return int(self._b.sample_actions(seed=(self.seed or 0) + 0x5EED, t=self.current_step)[0].item())
            return int(np.random.randint(0, self.action_space.n))   # the reference draws this one from the GLOBAL numpy stream (:163)
        if mode == "rule_based":                                          # :166-184 (device kernel)
            # the reference evaluates sorting_rules() AFTER update_environment has moved input -> belt
            return int(self._b.rule_based_actions(after_shift=True)[0].item())
        if mode == "model":                                               # :186-221
            obs = self._obs_after_shift()
            if self.sort_agent is not None:
                sm, _ = self.sort_agent.predict(obs[:13], deterministic=True)
                sort_mode = int(sm)
            else:
                sort_mode = int(self.rng_sorting.choice([0, 1]))                  # the env's own stream (:193)
            mask = self.action_masks()[:11]
            if self.press_agent is not None:
                # masks go to the agent only with masking on and a MaskablePPO-like agent (env_monolith.py:199-210)
                maskable = hasattr(self.press_agent, "policy") and "Maskable" in str(type(self.press_agent))
                if use_action_masking and maskable:
                    pa, _ = self.press_agent.predict(obs[13:], deterministic=True, action_masks=mask)
                else:
                    pa, _ = self.press_agent.predict(obs[13:], deterministic=True)
                press = int(pa)
            else:
                valid = np.flatnonzero(mask) if use_action_masking else np.arange(11)
                press = int(self.rng_pressing.choice(valid)) if valid.size else 0  # the env's own stream (:216-219)
            return sort_mode * 11 + press
        raise ValueError("Invalid action source: Provide 'action', set 'mode' to 'random', 'rule_based', "
                         "or 'model', or assign a mono_agent.")   # ref: env_monolith.py:224-225
