"""TEST INFRASTRUCTURE — drives the UNMODIFIED reference envs with PRESCRIBED random inputs.

The production kernels draw from Philox streams; the reference draws from numpy generators.  To compare the two on
the same trajectory, the oracle's PHILOX step records what it drew in the form the reference's generators would have
had to produce for the same outcome (mso_step_out_t rec_* in msort_oracle.c).  This module feeds exactly those values to
a reference env through stand-ins for its three generator ATTRIBUTES (plain instance attributes, env_super.py:170-174
— the reference source is not touched):

  env.rng_noise.uniform(low, high, 4)   (update_accuracy, env_super.py:508)   -> low + (high - low) * u   [numpy's formula]
  env.rng.choice(4, p=leftover/total)   (sort_material, env_super.py:563)     -> numpy's own algorithm on the prescribed u:
                                                                                  cdf = p.cumsum(); cdf /= cdf[-1];
                                                                                  cdf.searchsorted(u, side='right')
  env.rng_pressing.choice(valid)        (sample_masked_press_action, :291-300)-> the prescribed press action (must be valid)

and sets the freshly built input generator's `pattern_sequence` (input_generator.py:30) after every reset.  Only usable
where a copy of the reference exists (build container).  Used by tests/golden/make_philox_golden.py.
"""
from __future__ import annotations

import numpy as np

from .ref_loader import make_reference_env


class _Noise:
    def __init__(self):
        self.u = None

    def uniform(self, low, high, size=None):
        assert size == 4 and self.u is not None
        u, self.u = np.asarray(self.u, dtype=np.float64), None
        return low + (high - low) * u                 # numpy: low + (high - low) * random()


class _Choice:
    def __init__(self):
        self.queue, self.used = [], 0

    def choice(self, a, size=None, replace=True, p=None, **_):
        assert a == 4 and size is None and p is not None
        u = self.queue[self.used]
        self.used += 1
        cdf = np.asarray(p, dtype=np.float64).cumsum()
        cdf /= cdf[-1]
        return int(cdf.searchsorted(u, side="right"))

    def __getattr__(self, name):
        raise AttributeError(f"the reference used rng.{name}, which the prescribed-stream stand-in does not model")


class _Press:
    def __init__(self):
        self.next = None

    def choice(self, valid, *a, **k):
        assert self.next is not None and int(self.next) in [int(v) for v in valid], (self.next, valid)
        v, self.next = int(self.next), None
        return v


class DrivenReferenceEnv:
    """One reference env whose every random input is handed in by the caller."""

    def __init__(self, kind: str, *, max_steps: int, noise: float, balesize: int, seed: int = 0):
        self.kind = kind
        self.env = make_reference_env(kind, max_steps=max_steps, seed=seed, noise_sorting=noise, balesize=balesize)
        self._noise, self._choice, self._press = _Noise(), _Choice(), _Press()

    def _install(self, first_pattern: int):
        e = self.env
        e.rng_noise, e.rng, e.rng_pressing = self._noise, self._choice, self._press
        e.input_generator.pattern_sequence = np.array([first_pattern, 3 - first_pattern])

    def reset(self, first_pattern: int, seed=None):
        obs, _ = self.env.reset(seed=seed)            # seed=None afterwards: like SB3's VecEnv auto-reset (env_super.py:377)
        self._install(int(first_pattern))
        return np.asarray(obs, dtype=np.float32)

    def step(self, action: int, noise_u, redis_u, press_choice=None, use_action_masking=True, check_overflow=False):
        self._noise.u = noise_u
        self._choice.queue, self._choice.used = list(np.asarray(redis_u, dtype=np.float64)), 0
        if self.kind == "sort":
            self._press.next = press_choice
        obs, reward, term, trunc, info = self.env.step(int(action), use_action_masking=use_action_masking,
                                                       check_overflow=check_overflow)
        assert self._noise.u is None, "update_accuracy did not draw"
        assert self._choice.used == len(self._choice.queue), \
            f"sort_material made {self._choice.used} choice() calls, {len(self._choice.queue)} were prescribed"
        return np.asarray(obs, dtype=np.float32), float(reward), bool(term), info
