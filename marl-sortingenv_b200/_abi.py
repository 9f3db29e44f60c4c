"""ctypes mirror of include/msort.h and the loader of libmsort.so.

The library is the product: there is NO CPU fallback.  If the shared object is missing
or does not export every symbol of the header, `load_library()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 5
POLICY_ACT_WEIGHTS = 4256

ENV_SORT, ENV_PRESS, ENV_MONO = 1, 2, 3
KIND_BY_NAME = {"sort": ENV_SORT, "press": ENV_PRESS, "mono": ENV_MONO}
RNG_PHILOX, RNG_REPLAY = 0, 1
F_ACTION_MASKING, F_CHECK_OVERFLOW, F_AUTO_RESET, F_SORT_POLICY_MLP = 1, 2, 4, 8
OBS_DIM = {ENV_SORT: 13, ENV_PRESS: 16, ENV_MONO: 29}
NUM_ACTIONS = {ENV_SORT: 2, ENV_PRESS: 11, ENV_MONO: 22}
POLICY_WEIGHTS = 1570
NUM_STATS = 16

RESET_KEEP_STREAMS = 1

OK, E_INVALID, E_UNSUPPORTED, E_NO_DEVICE, E_CUDA, E_REPLAY = 0, -1, -2, -3, -4, -5


class MsortConfig(C.Structure):
    """msort_config_t (include/msort.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("env_kind", C.c_int32),
        ("num_envs", C.c_int64),
        ("global_env_offset", C.c_int64),
        ("max_steps", C.c_int32),
        ("flags", C.c_uint32),
        ("rng_mode", C.c_int32),
        ("reserved0", C.c_int32),
        ("seed", C.c_uint64),
        ("input_batch_size", C.c_int32),
        ("steps_per_pattern", C.c_int32),
        ("pattern_counts", (C.c_int32 * 4) * 2),
        ("baseline_accuracy", C.c_double * 4),
        ("boost", C.c_double),
        ("noise", C.c_double),
        ("stage_capacity", C.c_int32),
        ("press_time", C.c_int32 * 2),
        ("container_capacity", C.c_int32),
        ("bale_size", C.c_int32),
        ("reserved1", C.c_int32),
        ("bale_remainder_threshold", C.c_double),
        ("quality_threshold", C.c_double * 4),
        ("purity_theta", C.c_double),
        ("purity_scaling", C.c_double),
        ("tanh_temperature", C.c_double),
        ("overflow_penalty_catastrophic", C.c_double),
        ("overflow_penalty_severe", C.c_double),
        ("overflow_penalty_mild", C.c_double),
        ("bale_efficiency_factor", C.c_double),
        ("max_state_reward", C.c_double),
        ("overflow_termination_penalty", C.c_double),
    ]


class MsortEnvState(C.Structure):
    """msort_env_state_t (include/msort.h)."""
    _fields_ = [
        ("input", C.c_int32 * 4), ("belt", C.c_int32 * 4), ("sorting", C.c_int32 * 4),
        ("cont_true", C.c_int32 * 4), ("cont_false", C.c_int32 * 4), ("cont_e", C.c_int32),
        ("press_timer", C.c_int32 * 2), ("press_mat", C.c_int32 * 2),
        ("press_n", C.c_int32 * 2), ("press_q", C.c_int32 * 2),
        ("last_press_started", C.c_int32), ("last_press_amount", C.c_int32),
        ("gen_first", C.c_int32), ("gen_idx", C.c_int32), ("gen_counter", C.c_int32),
        ("step", C.c_int32), ("episode", C.c_int32), ("sensor_mode", C.c_int32),
        ("replay_cursor", C.c_int32),
        ("bale_n", C.c_int32 * 5), ("bale_last_size", C.c_int32 * 5),
        ("bale_last_q", C.c_int32 * 5), ("bale_sum", C.c_int32 * 5),
        ("reserved", C.c_int32),
        ("acc_belt", C.c_double * 4),
        ("ep_return", C.c_double),
    ]


# numpy structured dtype with the same layout (for export/import through torch uint8 buffers)
def env_state_dtype():
    import numpy as np
    dt = np.dtype([
        ("input", "<i4", (4,)), ("belt", "<i4", (4,)), ("sorting", "<i4", (4,)),
        ("cont_true", "<i4", (4,)), ("cont_false", "<i4", (4,)), ("cont_e", "<i4"),
        ("press_timer", "<i4", (2,)), ("press_mat", "<i4", (2,)),
        ("press_n", "<i4", (2,)), ("press_q", "<i4", (2,)),
        ("last_press_started", "<i4"), ("last_press_amount", "<i4"),
        ("gen_first", "<i4"), ("gen_idx", "<i4"), ("gen_counter", "<i4"),
        ("step", "<i4"), ("episode", "<i4"), ("sensor_mode", "<i4"),
        ("replay_cursor", "<i4"),
        ("bale_n", "<i4", (5,)), ("bale_last_size", "<i4", (5,)),
        ("bale_last_q", "<i4", (5,)), ("bale_sum", "<i4", (5,)),
        ("reserved", "<i4"),
        ("acc_belt", "<f8", (4,)),
        ("ep_return", "<f8"),
    ], align=True)
    assert dt.itemsize == C.sizeof(MsortEnvState), (dt.itemsize, C.sizeof(MsortEnvState))
    return dt


class MsortReplay(C.Structure):
    """msort_replay_t (include/msort.h) — device pointers as integers."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("reserved", C.c_uint32),
        ("noise_u", C.c_void_p), ("redis_u", C.c_void_p), ("redis_len", C.c_int64),
        ("input_counts", C.c_void_p), ("press_choice", C.c_void_p), ("sort_mode", C.c_void_p),
    ]


class MsortInfoOut(C.Structure):
    """msort_info_out_t (include/msort.h) — device pointers as integers."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("reserved", C.c_uint32),
        ("action", C.c_void_p), ("overflow", C.c_void_p), ("overflow_material", C.c_void_p),
        ("sort_mode", C.c_void_p), ("press_action", C.c_void_p), ("invalid_action", C.c_void_p),
        ("terminal_obs", C.c_void_p), ("episode_return", C.c_void_p),
        ("episode_length", C.c_void_p), ("stats", C.c_void_p),
        ("reward_sort", C.c_void_p), ("reward_press", C.c_void_p), ("sorted_true", C.c_void_p),
    ]


class MsortHostIO(C.Structure):
    """msort_host_io_t (include/msort.h)."""
    _fields_ = [("struct_size", C.c_uint32), ("chunks", C.c_uint32), ("actions_u8", C.c_void_p),
                ("actions_i64", C.c_void_p), ("obs", C.c_void_p), ("reward", C.c_void_p), ("flags", C.c_void_p),
                ("dev_obs", C.c_void_p), ("dev_reward", C.c_void_p), ("dev_terminated", C.c_void_p), ("dev_mask", C.c_void_p)]


class MsortPpoBatch(C.Structure):
    """msort_ppo_batch_t (include/msort.h)."""
    _fields_ = [("struct_size", C.c_uint32), ("obs_dim", C.c_int32), ("num_actions", C.c_int32), ("reserved", C.c_int32),
                ("num_rows", C.c_int64), ("obs", C.c_void_p), ("mask", C.c_void_p), ("actions", C.c_void_p),
                ("old_logp", C.c_void_p), ("adv", C.c_void_p), ("ret", C.c_void_p)]


class MsortPpoHparams(C.Structure):
    """msort_ppo_hparams_t (include/msort.h)."""
    _fields_ = [("struct_size", C.c_uint32), ("normalize_advantage", C.c_int32), ("clip_range", C.c_float),
                ("vf_coef", C.c_float), ("ent_coef", C.c_float), ("learning_rate", C.c_float), ("beta1", C.c_float),
                ("beta2", C.c_float), ("adam_eps", C.c_float), ("max_grad_norm", C.c_float)]


# name -> (restype, argtypes): every symbol include/msort.h declares
_P = C.c_void_p
SYMBOLS = {
    "msort_abi_version": (C.c_int, []),
    "msort_last_error": (C.c_char_p, []),
    "msort_default_config": (C.c_int, [C.c_int, C.POINTER(MsortConfig)]),
    "msort_create": (C.c_int, [C.POINTER(MsortConfig), C.c_int, C.POINTER(_P)]),
    "msort_destroy": (C.c_int, [_P]),
    "msort_state_bytes": (C.c_size_t, [_P]),
    "msort_obs_dim": (C.c_int, [_P]),
    "msort_num_actions": (C.c_int, [_P]),
    "msort_reset": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_uint32, _P]),
    "msort_set_flags": (C.c_int, [_P, C.c_uint32]),
    "msort_set_seed": (C.c_int, [_P, C.c_uint64]),
    "msort_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.POINTER(MsortInfoOut),
                             C.POINTER(MsortReplay), _P]),
    "msort_step_range": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, _P, C.POINTER(MsortInfoOut), _P]),
    "msort_set_policy": (C.c_int, [_P, _P, C.c_int, _P]),
    "msort_sample_actions": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint32, _P]),
    "msort_rule_based_actions": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "msort_observe": (C.c_int, [_P, _P, _P, _P, _P]),
    "msort_observe_after_shift": (C.c_int, [_P, _P, _P, _P, _P]),
    "msort_policy_act": (C.c_int, [_P, _P, _P, _P, C.c_uint64, C.c_uint32, C.c_int, _P, _P, _P, _P]),
    "msort_policy_act_range": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, C.c_uint64, C.c_uint32, C.c_int, _P, _P, _P, _P]),
    "msort_export_state": (C.c_int, [_P, _P, _P, _P]),
    "msort_gather_state": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P]),
    "msort_import_state": (C.c_int, [_P, _P, _P, _P]),
    "msort_reduce_stats": (C.c_int, [_P, _P, _P, _P]),
    "msort_sync_check": (C.c_int, [_P, _P]),
    "msort_launch_count": (C.c_int64, [_P]),
    "msort_step_variant": (C.c_int, [_P]),
    "msort_ppo_param_count": (C.c_int, [C.c_int, C.c_int]),
    "msort_ppo_scratch_floats": (C.c_int, [C.c_int, C.c_int]),
    "msort_ppo_forward": (C.c_int, [C.POINTER(MsortPpoBatch), _P, _P, _P, _P]),
    "msort_ppo_gae": (C.c_int, [C.c_int32, C.c_int64, _P, _P, _P, _P, C.c_float, C.c_float, _P, _P, _P]),
    "msort_ppo_gradient": (C.c_int, [C.POINTER(MsortPpoBatch), C.POINTER(MsortPpoHparams), _P, _P, _P, C.c_int64, C.c_int64, _P, _P, _P]),
    "msort_ppo_update": (C.c_int, [C.POINTER(MsortPpoBatch), C.POINTER(MsortPpoHparams), _P, _P, _P, _P, _P, _P, C.c_int32,
                                   C.c_int64, _P, _P, _P]),
    "msort_policy_eval": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P, C.c_uint64, C.c_uint32, C.c_int,
                                    _P, _P, _P, _P]),
    "msort_rollout_pack": (C.c_int, [_P, _P, _P]),
    "msort_rollout_policy": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, C.c_uint64, C.c_uint32, C.c_int, _P, _P, _P, _P]),
    "msort_rollout_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.POINTER(MsortInfoOut), _P, C.c_uint64, C.c_uint32, C.c_int,
                                     _P, _P, _P, _P]),
    "msort_generate_streams": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P]),
    "msort_host_scratch_bytes": (C.c_size_t, [_P]),
    "msort_step_host": (C.c_int, [_P, _P, _P, C.POINTER(MsortHostIO), C.POINTER(MsortInfoOut), _P]),
    "msort_set_option": (C.c_int, [_P, C.c_int, C.c_int64]),
    "msort_get_option": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int64)]),
    "msort_debug_policy_logits": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
}
ROLLOUT_WEIGHTS = 2688
OPT_TENSOR_POLICY = 1
OPT_PERSIST_CTAS = 2
OPT_DRAW_COUNTER = 3

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "csrc", "libmsort.so")
_lib = None


class MsortError(RuntimeError):
    def __init__(self, code: int, where: str, text: str):
        super().__init__(f"{where} failed with status {code}: {text}")
        self.code = code


def load_library(path: str | None = None):
    """dlopen libmsort.so and bind every symbol.  Raises if anything is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("MSORT_LIB", LIB_PATH)
    if not os.path.isfile(p):
        raise ImportError(
            f"libmsort.so not found at {p}. Build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    ver = lib.msort_abi_version()
    if ver != ABI_VERSION:
        raise ImportError(f"libmsort.so ABI version {ver} != expected {ABI_VERSION}")
    if path is None:
        _lib = lib
    return lib


def check(lib, code: int, where: str):
    if code != 0:
        msg = lib.msort_last_error()
        raise MsortError(code, where, msg.decode() if msg else "")
