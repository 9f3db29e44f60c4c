"""numpy-in / numpy-out adapter over the CUDA library (through the public Python host, i.e.
through the C ABI) with the same call shape as oracle.cpu_oracle.OracleEnv, so one parity
driver (parity_util.replay_and_compare) serves both."""
from __future__ import annotations

import numpy as np
import torch

from marl_sortingenv_b200 import _abi
from marl_sortingenv_b200.batched import ENV_CLASSES

_KIND = {_abi.ENV_SORT: "sort", _abi.ENV_PRESS: "press", _abi.ENV_MONO: "mono"}


class CudaBackend:
    def __init__(self, cfg: _abi.MsortConfig, device="cuda:0", config=None, policy=None, info_level="full"):
        kind = _KIND[cfg.env_kind]
        f = int(cfg.flags)
        self.env = ENV_CLASSES[kind](
            int(cfg.num_envs), device=device, max_steps=int(cfg.max_steps), seed=int(cfg.seed),
            noise_sorting=float(cfg.noise), balesize=int(cfg.bale_size), config=config,
            use_action_masking=bool(f & _abi.F_ACTION_MASKING), check_overflow=bool(f & _abi.F_CHECK_OVERFLOW),
            auto_reset=bool(f & _abi.F_AUTO_RESET),
            rng_mode="replay" if cfg.rng_mode == _abi.RNG_REPLAY else "philox",
            global_env_offset=int(cfg.global_env_offset), info_level=info_level)
        # the env is re-created from scalar arguments: make sure it digested the very same msort_config_t
        import copy
        want = copy.copy(cfg)
        want.flags = int(cfg.flags) & ~_abi.F_SORT_POLICY_MLP | (int(self.env.cfg.flags) & _abi.F_SORT_POLICY_MLP)
        assert bytes(self.env.cfg) == bytes(want), "CudaBackend: the env's config differs from the requested one"
        self.n = int(cfg.num_envs)
        self._redis = None
        if policy is not None:
            self.set_policy(policy)

    def set_policy(self, weights):
        self.env.set_sort_policy(np.asarray(weights, dtype=np.float32))

    def reset(self, which=None, first_pattern=None):
        obs, _ = self.env.reset(which=which, first_pattern=first_pattern)
        torch.cuda.synchronize()
        return obs.cpu().numpy().copy(), self.env.action_masks().cpu().numpy().copy()

    def step(self, actions, *, noise_u=None, redis_u=None, input_counts=None, press_choice=None,
             sort_mode=None, want_info=True):
        replay = None
        if self.env.rng_mode == "replay":
            if self._redis is None or self._redis[0] is not redis_u:
                t = torch.as_tensor(np.ascontiguousarray(redis_u, dtype=np.float64)).reshape(self.n, -1).cuda()
                self._redis = (redis_u, t)
            replay = dict(noise_u=np.asarray(noise_u, dtype=np.float64), redis_u=self._redis[1],
                          input_counts=input_counts, press_choice=press_choice, sort_mode=sort_mode)
        a = torch.as_tensor(np.asarray(actions, dtype=np.int64)).cuda()
        obs, rew, term, trunc, info = self.env.step(a, replay=replay)
        self.env.sync_check()
        assert not bool(trunc.any())
        out = {k: v.cpu().numpy().copy() for k, v in info.items()}
        if "terminal_observation" in out:
            out["terminal_obs"] = out.pop("terminal_observation")
        return (obs.cpu().numpy().copy(), rew.cpu().numpy().astype(np.float64), term.cpu().numpy().copy(),
                self.env.action_masks().cpu().numpy().copy(), out)

    def export_state(self):
        return self.env.export_state()

    @property
    def state(self):
        return self.env.export_state()
