"""TEST INFRASTRUCTURE — ctypes wrapper around oracle/libmsort_oracle.so (the CPU checker).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  It reuses the product's ctypes struct definitions (config, plain
env state, replay/info descriptors) so both sides are driven by identical inputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from marl_sortingenv_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmsort_oracle.so")
_lib = None


class _StepOut(C.Structure):
    """mso_step_out_t (oracle/msort_oracle.c)."""
    _fields_ = [("reward32", C.c_void_p), ("reward64", C.c_void_p), ("mlp_margin", C.c_void_p),
                ("rec_noise_u", C.c_void_p), ("rec_redis_u", C.c_void_p), ("rec_n_draws", C.c_void_p),
                ("rec_press_choice", C.c_void_p), ("rec_input_counts", C.c_void_p), ("rec_cap", C.c_int32)]


REC_CAP = 128      # uniforms recorded per env-step: one per mis-sorted unit, at most the input batch (100 by default)


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile (gcc, no FMA contraction)."""
    src = os.path.join(_HERE, "msort_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "msort.h")
    stale = (not os.path.isfile(_LIB_PATH)
             or any(os.path.isfile(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH)
                    for p in (src, hdr)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmsort_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.mso_reset.restype = C.c_int
        L.mso_reset.argtypes = [C.POINTER(_abi.MsortConfig), C.c_void_p, C.c_int64, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]
        L.mso_step.restype = C.c_int
        L.mso_step.argtypes = [C.POINTER(_abi.MsortConfig), C.c_void_p, C.c_int64, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_StepOut),
                               C.POINTER(_abi.MsortInfoOut), C.POINTER(_abi.MsortReplay),
                               C.c_void_p, C.c_int]
        L.mso_rollout.restype = C.c_int64
        L.mso_rollout.argtypes = [C.POINTER(_abi.MsortConfig), C.c_void_p, C.c_int64, C.c_int,
                                  C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]
        L.mso_sample_masked_actions.restype = None
        L.mso_sample_masked_actions.argtypes = [C.POINTER(_abi.MsortConfig), C.c_void_p, C.c_int64,
                                                C.c_uint64, C.c_uint32, C.c_void_p]
        L.mso_rule_based_actions.restype = None
        L.mso_rule_based_actions.argtypes = [C.POINTER(_abi.MsortConfig), C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.mso_philox4x32_10.restype = None
        L.mso_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mso_state_size.restype = C.c_int
        L.mso_config_size.restype = C.c_int
        assert L.mso_state_size() == C.sizeof(_abi.MsortEnvState)
        assert L.mso_config_size() == C.sizeof(_abi.MsortConfig)
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_counts(counts) -> np.ndarray:
    """[...,4] integer counts → packed uint32 A|B<<8|C<<16|D<<24 (msort_replay_t.input_counts)."""
    c = np.asarray(counts, dtype=np.uint32)
    return (c[..., 0] | (c[..., 1] << 8) | (c[..., 2] << 16) | (c[..., 3] << 24)).astype(np.uint32)


class OracleEnv:
    """Batched CPU oracle with the same call shape as the device library."""

    def __init__(self, cfg: _abi.MsortConfig, nthreads: int = 1):
        self.L = lib()
        self.cfg = cfg
        self.n = int(cfg.num_envs)
        self.D = _abi.OBS_DIM[cfg.env_kind]
        self.A = _abi.NUM_ACTIONS[cfg.env_kind]
        self.state = np.zeros(self.n, dtype=_abi.env_state_dtype())
        self.nthreads = nthreads
        self.policy = None
        self.stats = np.zeros(_abi.NUM_STATS, dtype=np.float64)

    def set_policy(self, weights):
        self.policy = np.ascontiguousarray(weights, dtype=np.float32)
        assert self.policy.size == _abi.POLICY_WEIGHTS

    def reset(self, which=None, first_pattern=None):
        obs = np.zeros((self.n, self.D), dtype=np.float32)
        mask = np.zeros((self.n, self.A), dtype=np.uint8)
        w = None if which is None else np.ascontiguousarray(which, dtype=np.uint8)
        fp = None if first_pattern is None else np.ascontiguousarray(first_pattern, dtype=np.uint8)
        rc = self.L.mso_reset(C.byref(self.cfg), _ptr(self.state), self.n, _ptr(w), _ptr(fp),
                              _ptr(obs), _ptr(mask))
        assert rc == 0
        return obs, mask.astype(bool)

    def step(self, actions, *, noise_u=None, redis_u=None, input_counts=None, press_choice=None,
             sort_mode=None, want_info=True, record=False):
        """`record=True` (PHILOX mode): the info dict also carries the step's random inputs in REPLAY form — `rec_noise_u`
        [n,4], `rec_redis_u` [n,REC_CAP] with `rec_n_draws` [n] valid entries each, `rec_press_choice` [n] (Env_1),
        `rec_input_counts` [n] packed — see mso_step_out_t in msort_oracle.c."""
        n, D, A = self.n, self.D, self.A
        actions = np.ascontiguousarray(actions, dtype=np.int64)
        obs = np.zeros((n, D), dtype=np.float32)
        term = np.zeros(n, dtype=np.uint8)
        mask = np.zeros((n, A), dtype=np.uint8)
        r64 = np.zeros(n, dtype=np.float64)
        margin = np.full(n, np.inf, dtype=np.float32)
        so = _StepOut(None, _ptr(r64), _ptr(margin))
        rec = {}
        if record:
            rec = dict(rec_noise_u=np.zeros((n, 4)), rec_redis_u=np.zeros((n, REC_CAP)), rec_n_draws=np.zeros(n, np.int32),
                       rec_press_choice=np.zeros(n, np.uint8), rec_input_counts=np.zeros(n, np.uint32))
            for k, v in rec.items():
                setattr(so, k, _ptr(v))
            so.rec_cap = REC_CAP
        info = _abi.MsortInfoOut()
        info.struct_size = C.sizeof(info)
        keep = {}
        if want_info:
            keep = dict(action=np.zeros(n, np.int64), overflow=np.zeros(n, np.uint8),
                        overflow_material=np.full(n, -1, np.int8), sort_mode=np.zeros(n, np.uint8),
                        press_action=np.zeros(n, np.uint8), invalid_action=np.zeros(n, np.uint8),
                        reward_sort=np.zeros(n, np.float32), reward_press=np.zeros(n, np.float32),
                        sorted_true=np.zeros(n, np.uint32),
                        terminal_obs=np.zeros((n, D), np.float32),
                        episode_return=np.zeros(n, np.float64),
                        episode_length=np.zeros(n, np.int32))
            for k, v in keep.items():
                setattr(info, k, _ptr(v))
            info.stats = _ptr(self.stats)
        rp = None
        hold = []
        if self.cfg.rng_mode == _abi.RNG_REPLAY:
            rp = _abi.MsortReplay()
            rp.struct_size = C.sizeof(rp)
            nz = np.ascontiguousarray(noise_u, dtype=np.float64).reshape(n, 4)
            ru = np.ascontiguousarray(redis_u, dtype=np.float64).reshape(n, -1)
            hold += [nz, ru]
            rp.noise_u, rp.redis_u, rp.redis_len = _ptr(nz), _ptr(ru), ru.shape[1]
            if input_counts is not None:
                ic = np.ascontiguousarray(input_counts, dtype=np.uint32).reshape(n)
                hold.append(ic); rp.input_counts = _ptr(ic)
            if press_choice is not None:
                pc = np.ascontiguousarray(press_choice, dtype=np.uint8).reshape(n)
                hold.append(pc); rp.press_choice = _ptr(pc)
            if sort_mode is not None:
                sm = np.ascontiguousarray(sort_mode, dtype=np.uint8).reshape(n)
                hold.append(sm); rp.sort_mode = _ptr(sm)
        rc = self.L.mso_step(C.byref(self.cfg), _ptr(self.state), n, _ptr(actions), _ptr(obs),
                             _ptr(term), _ptr(mask), C.byref(so), C.byref(info),
                             C.byref(rp) if rp is not None else None, _ptr(self.policy),
                             self.nthreads)
        if rc != 0:
            raise RuntimeError(f"mso_step returned {rc}")
        keep["mlp_margin"] = margin
        if record:
            assert int(rec["rec_n_draws"].max(initial=0)) <= REC_CAP, "more draws in one step than REC_CAP holds"
            keep.update(rec)
        return obs, r64, term.astype(bool), mask.astype(bool), keep

    def sample_masked_actions(self, seed: int, t: int):
        act = np.zeros(self.n, dtype=np.int64)
        self.L.mso_sample_masked_actions(C.byref(self.cfg), _ptr(self.state), self.n, seed, t, _ptr(act))
        return act

    def rule_based_actions(self, after_shift: bool = True):
        act = np.zeros(self.n, dtype=np.int64)
        self.L.mso_rule_based_actions(C.byref(self.cfg), _ptr(self.state), self.n, 1 if after_shift else 0, _ptr(act))
        return act

    def rollout(self, T: int, action_seed: int = 1):
        stats = np.zeros(_abi.NUM_STATS, dtype=np.float64)
        done = self.L.mso_rollout(C.byref(self.cfg), _ptr(self.state), self.n, T, action_seed,
                                  _ptr(self.policy), self.nthreads, _ptr(stats))
        return int(done), stats


def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().mso_philox4x32_10(_ptr(c), _ptr(k), _ptr(o))
    return o
