"""Observation/action spaces of the three envs.

ref: _initialize_spaces in env_1_sort.py:43-72, env_2_press.py:45-64, env_monolith.py:49-79.
`gymnasium` is used when importable; otherwise minimal look-alikes with the attributes SB3
and the reference read (`shape`, `dtype`, `low`, `high`, `n`, `sample`, `contains`).
"""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.low = np.asarray(low, dtype=self.dtype)
        self.high = np.asarray(high, dtype=self.dtype)
        self.shape = self.low.shape
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.shape}, {self.dtype})"


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)
        self.start = 0
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return int(self._rng.integers(0, self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return f"Discrete({self.n})"


def _bounds(kind):
    sort_low = np.concatenate([np.zeros(9), np.full(4, -1.0)])   # occ, 4 props, 4 acc, 4 purity diffs
    sort_high = np.ones(13)
    press_low, press_high = np.zeros(16), np.ones(16)
    if kind == "sort":
        return sort_low, sort_high, 2
    if kind == "press":
        return press_low, press_high, 11
    return np.concatenate([sort_low, press_low]), np.concatenate([sort_high, press_high]), 22


def make_spaces(kind: str):
    low, high, n = _bounds(kind)
    try:
        from gymnasium import spaces as gsp  # pragma: no cover - not installed in the build image
        return (gsp.Box(low.astype(np.float32), high.astype(np.float32), dtype=np.float32),
                gsp.Discrete(n))
    except Exception:
        return Box(low, high, np.float32), Discrete(n)
