"""msort — B200-native batched simulator behind MARL-SortingEnv's Gymnasium surface.

Import as `marl_sortingenv_b200` (alias package at the repo root).  Everything that
computes runs in libmsort.so (hand-written sm_100a CUDA behind a C ABI, include/msort.h);
this package is the Python host: config parsing, buffer ownership (torch), the batched
env classes, the single-env reference-signature classes and the SB3 VecEnv adapter.
"""
from . import _abi, config  # noqa: F401
from .config import DEFAULT_CONFIG, load_config, make_config  # noqa: F401

__all__ = ["DEFAULT_CONFIG", "load_config", "make_config"]


def __getattr__(name):  # lazy: the env classes need torch + the CUDA library
    if name in ("BatchedSortingEnv", "BatchedPressingEnv", "BatchedMonolithEnv", "BatchedEnv"):
        from . import batched
        return getattr(batched, name)
    if name in ("Env_1_Sorting", "Env_2_Pressing", "Env_3_Monolith"):
        from . import single
        return getattr(single, name)
    if name in ("TraceRecorder",):
        from . import telemetry
        return telemetry.TraceRecorder
    if name in ("MsortVecEnv",):
        from . import vecenv
        return getattr(vecenv, name)
    raise AttributeError(name)
