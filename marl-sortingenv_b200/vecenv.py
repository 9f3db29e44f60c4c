"""Duck-typed Stable-Baselines3 `VecEnv` over the batched device simulator.

The reference trains through SB3, which wraps the single env in `DummyVecEnv(Monitor(
ActionMasker(env)))` (training.py:58-69,118-143).  This adapter presents N device envs with the
same protocol so an SB3-style rollout loop can drive them without those wrappers:

* `reset()` → obs `[N, D]`;  `step_async(actions)` / `step_wait()` → `(obs, rewards, dones, infos)`
* auto-reset on `done` with `infos[i]["terminal_observation"]` (VecEnv contract),
  `infos[i]["episode"] = {"r", "l", "t"}` (Monitor contract), `"TimeLimit.truncated": False`
* `env_method("action_masks")` → list of N masks (what sb3_contrib's `get_action_masks` stacks)
* `get_attr / set_attr / env_is_wrapped / seed / close`

SB3 is not installed in the build image, so this class does not inherit from
`stable_baselines3.common.vec_env.VecEnv`; it implements the methods SB3's on-policy
collectors call.  Per-env Python dicts make it suitable for N up to ~1e4; beyond that use the
tensor API (`BatchedEnv.step`) directly.
"""
from __future__ import annotations

import time

import numpy as np

from .batched import ENV_CLASSES


class MsortVecEnv:
    def __init__(self, kind: str, num_envs: int, **env_kwargs):
        env_kwargs.setdefault("auto_reset", True)
        env_kwargs.setdefault("info_level", "episode")
        self.env = ENV_CLASSES[kind](num_envs, **env_kwargs)
        self.num_envs = int(num_envs)
        self.observation_space = self.env.observation_space
        self.action_space = self.env.action_space
        self.render_mode = None
        self.name = self.env.name
        self._actions = None
        self._t0 = time.time()
        self.reset_infos = [{} for _ in range(self.num_envs)]

    # ------------------------------------------------------------------ VecEnv protocol
    def reset(self):
        obs, _ = self.env.reset()
        return obs.cpu().numpy().copy()

    def step_async(self, actions):
        self._actions = np.asarray(actions, dtype=np.int64).reshape(self.num_envs)

    def step_wait(self):
        obs, rew, term, trunc, mask = self.env.step_host(self._actions)
        term = np.asarray(term)                      # step_host hands `terminated` over lazily (packed flag words)
        infos = [{"TimeLimit.truncated": False} for _ in range(self.num_envs)]
        done_idx = np.flatnonzero(term)
        if done_idx.size:
            b = self.env.info_buffers
            t_obs = b["terminal_observation"][done_idx.tolist()].cpu().numpy()
            ep_r = b["episode_return"][done_idx.tolist()].cpu().numpy()
            ep_l = b["episode_length"][done_idx.tolist()].cpu().numpy()
            now = round(time.time() - self._t0, 6)
            for k, i in enumerate(done_idx):
                infos[i]["terminal_observation"] = t_obs[k]
                infos[i]["episode"] = {"r": round(float(ep_r[k]), 6), "l": int(ep_l[k]), "t": now}
        return obs.copy(), rew.copy(), term.copy(), infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()

    def seed(self, seed=None):
        if seed is not None:
            self.env.reset(seed=int(seed))
        return [seed] * self.num_envs

    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices

    def env_method(self, method_name, *args, indices=None, **kwargs):
        idx = list(self._indices(indices))
        if method_name == "action_masks":
            m = self.env.action_masks().cpu().numpy()
            return [m[i].copy() for i in idx]
        if method_name == "get_obs":
            o = self.env.get_obs().cpu().numpy()
            return [o[i].copy() for i in idx]
        raise AttributeError(f"env_method({method_name!r}) is not supported by the batched env")

    def get_attr(self, attr_name, indices=None):
        return [getattr(self.env, attr_name) for _ in self._indices(indices)]

    def set_attr(self, attr_name, value, indices=None):
        setattr(self.env, attr_name, value)

    def has_attr(self, attr_name):
        return hasattr(self.env, attr_name)

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._indices(indices)]

    def get_images(self):
        return [None] * self.num_envs

    def render(self, mode=None):
        return None

    @property
    def unwrapped(self):
        return self

    def action_masks(self):
        """Stacked masks [N, A] (convenience for MaskablePPO-style loops)."""
        return self.env.action_masks().cpu().numpy().copy()
