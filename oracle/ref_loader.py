"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference envs from /root/reference.

This file is part of the oracle (test infrastructure).  Nothing in the product
package imports it.  It only works inside the build container, where the
read-only reference checkout exists; it cannot travel to the GPU box, which is
why the trajectories it records are committed as fixtures under tests/golden/.

What it does
------------
The reference (src/envs_train/env_super.py:5,9,10 ; utils/input_generator.py:5-6 ;
utils/plotting.py:5-11) imports `gymnasium`, `matplotlib`, `seaborn` and
`mpl_toolkits` at module top.  None of them is installed here and none is on the
step()/reset() hot path: `gymnasium` supplies only the base class `gym.Env` and
the `spaces.Box/Discrete` containers, the plotting packages are used only by
`render()`.  We register minimal in-memory stand-ins for those modules and then
import the reference classes from where they lie.  No reference source is
copied or modified.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np

def _find_reference_root() -> str:
    """$MSORT_REFERENCE_ROOT, else the read-only checkout of the build container, else the copy oracle/make_ref.sh
    staged into oracle/_ref (git-ignored; it is what travels to the GPU box)."""
    env = os.environ.get("MSORT_REFERENCE_ROOT")
    if env:
        return env
    staged = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    for root in ("/root/reference", staged):
        if os.path.isfile(os.path.join(root, "src", "envs_train", "env_super.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "envs_train", "env_super.py"))


# --------------------------------------------------------------------------
# stand-in modules
# --------------------------------------------------------------------------
class _Space:
    def seed(self, seed=None):
        self._seed = seed
        return [seed]


class _Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.low = np.asarray(low, dtype=self.dtype)
        self.high = np.asarray(high, dtype=self.dtype)
        self.shape = tuple(self.low.shape) if shape is None else tuple(shape)


class _Discrete(_Space):
    def __init__(self, n, start=0):
        self.n = int(n)
        self.start = int(start)
        self.shape = ()
        self.dtype = np.dtype(np.int64)


class _Env:
    metadata = {}

    @property
    def unwrapped(self):
        return self


class _Anything(types.ModuleType):
    """A module whose every attribute is another permissive stand-in."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Anything(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):  # pragma: no cover - render() is never called
        return self


def _install_stubs() -> None:
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")
        spaces = types.ModuleType("gymnasium.spaces")
        spaces.Box, spaces.Discrete, spaces.Space = _Box, _Discrete, _Space
        gym.Env, gym.spaces = _Env, spaces
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.colors",
                 "seaborn", "mpl_toolkits", "mpl_toolkits.axes_grid1",
                 "mpl_toolkits.axes_grid1.inset_locator"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Anything(name)


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


_CLASSES = None


def load_reference():
    """Return (Env_1_Sorting, Env_2_Pressing, Env_3_Monolith, Env_Super) of the reference."""
    global _CLASSES
    if _CLASSES is not None:
        return _CLASSES
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from src.envs_train.env_1_sort import Env_1_Sorting      # noqa: E402
    from src.envs_train.env_2_press import Env_2_Pressing    # noqa: E402
    from src.envs_train.env_monolith import Env_3_Monolith   # noqa: E402
    from src.envs_train.env_super import Env_Super           # noqa: E402
    Env_Super.render = lambda self, *a, **k: None            # plotting is stubbed out
    _CLASSES = (Env_1_Sorting, Env_2_Pressing, Env_3_Monolith, Env_Super)
    return _CLASSES


def make_reference_env(kind: str, **kwargs):
    """Instantiate a reference env.  `config.yml` is resolved relative to the cwd
    (env_super.py:25), so construction happens with cwd = reference root."""
    e1, e2, e3, _ = load_reference()
    cls = {"sort": e1, "press": e2, "mono": e3}[kind]
    with _cwd(REFERENCE_ROOT):
        return cls(**kwargs)
