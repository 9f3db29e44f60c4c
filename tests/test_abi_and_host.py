"""CPU-side checks (no GPU, no compute calls): the C-ABI library loads and exports every symbol
include/msort.h declares; struct mirrors match; config parsing follows the reference's rules;
the library refuses to run without a B200 (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import marl_sortingenv_b200 as pkg
from marl_sortingenv_b200 import _abi, config as cfgmod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "msort.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msort_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    lib = _abi.load_library()
    names = _header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"libmsort.so does not export {n}"
        assert n in _abi.SYMBOLS, f"python binding misses {n}"
    assert lib.msort_abi_version() == _abi.ABI_VERSION


def test_struct_mirrors_match_c_layout():
    from oracle import cpu_oracle
    L = cpu_oracle.lib()
    assert L.mso_config_size() == C.sizeof(_abi.MsortConfig)
    assert L.mso_state_size() == C.sizeof(_abi.MsortEnvState) == _abi.env_state_dtype().itemsize
    lib = _abi.load_library()
    c = _abi.MsortConfig()
    assert lib.msort_default_config(_abi.ENV_MONO, C.byref(c)) == 0
    assert c.struct_size == C.sizeof(_abi.MsortConfig)
    ref = cfgmod.make_config("mono", 1)
    for f, _ in _abi.MsortConfig._fields_:
        if f in ("flags", "seed", "num_envs"):
            continue
        a, b = getattr(c, f), getattr(ref, f)
        if hasattr(a, "__len__"):
            a, b = np.array(a).tolist(), np.array(b).tolist()
        assert a == b, f
    assert lib.msort_default_config(9, C.byref(c)) == _abi.E_INVALID
    assert b"unknown env kind" in lib.msort_last_error()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _abi.load_library()
    h = C.c_void_p()
    c = cfgmod.make_config("mono", 128)
    rc = lib.msort_create(C.byref(c), 0, C.byref(h))
    assert rc == _abi.E_NO_DEVICE and not h.value
    assert b"no CPU fallback" in lib.msort_last_error() or b"sm_100a" in lib.msort_last_error()
    with pytest.raises(RuntimeError):
        pkg.BatchedMonolithEnv(16)


def test_create_rejects_bad_config():
    lib = _abi.load_library()
    h = C.c_void_p()
    for mut, frag in ((dict(struct_size=12), b"struct_size"), (dict(num_envs=0), b"num_envs"),
                      (dict(input_batch_size=300), b"input_batch_size"), (dict(max_steps=0), b"max_steps"),
                      (dict(env_kind=7), b"env kind")):
        c = cfgmod.make_config("press", 64)
        for k, v in mut.items():
            setattr(c, k, v)
        rc = lib.msort_create(C.byref(c), 0, C.byref(h))
        assert rc in (_abi.E_INVALID, _abi.E_UNSUPPORTED) and frag in lib.msort_last_error(), (mut, lib.msort_last_error())


def test_config_follows_reference_rules(tmp_path):
    c = cfgmod.make_config("sort", 4, noise_sorting=None, balesize=None)
    assert c.noise == 0.05 and c.bale_size == 200            # config.yml values when ctor passes None
    c = cfgmod.make_config("sort", 4, noise_sorting=0.0, balesize=150)
    assert c.noise == 0.0 and c.bale_size == 150             # ctor overrides (env_super.py:71,87)
    assert c.steps_per_pattern == 20                         # env_super.py:375: default after reset
    assert [c.pattern_counts[0][m] for m in range(4)] == [40, 15, 35, 10]
    assert [c.pattern_counts[1][m] for m in range(4)] == [15, 40, 10, 35]
    assert c.purity_scaling == 2.0                           # hard-coded at env_super.py:971
    y = tmp_path / "config.yml"
    y.write_text("pressing_station:\n  container_capacity: 900\n  press_times: {1: 8, 2: 9}\n"
                 "simulation:\n  steps_per_pattern: 5\nrewards:\n  sorting:\n    tanh_temperature: 0.25\n")
    c = cfgmod.make_config("mono", 4, config_path=str(y))
    assert c.container_capacity == 900 and (c.press_time[0], c.press_time[1]) == (8, 9)
    assert c.tanh_temperature == 0.25 and c.steps_per_pattern == 20   # dead key after reset
    assert c.boost == 0.5                                    # untouched keys keep the shipped values


def test_spaces_match_reference_shapes():
    from marl_sortingenv_b200.spaces import make_spaces
    for kind, (D, A) in {"sort": (13, 2), "press": (16, 11), "mono": (29, 22)}.items():
        ob, ac = make_spaces(kind)
        assert ob.shape == (D,) and ob.dtype == np.float32 and ac.n == A
    ob, _ = make_spaces("mono")
    assert ob.low[9:13].tolist() == [-1.0] * 4 and ob.low[:9].tolist() == [0.0] * 9 and ob.high.max() == 1.0


def test_policy_flattening_sb3_layout():
    import torch
    from marl_sortingenv_b200.policy import SB3_KEYS, SHAPES, flatten_sort_policy, sb3_style_init
    sd = {k: torch.randn(*s) for k, s in zip(SB3_KEYS, SHAPES)}
    w = flatten_sort_policy(sd)
    assert w.numel() == 1570 and torch.equal(w[:416].reshape(32, 13), sd[SB3_KEYS[0]])
    assert torch.equal(w[-2:], sd[SB3_KEYS[5]])
    assert flatten_sort_policy(sb3_style_init(1)).numel() == 1570
    with pytest.raises(ValueError):
        flatten_sort_policy(np.zeros(10, np.float32))


def test_load_sb3_zip_layout(tmp_path):
    import io
    import zipfile
    import torch
    from marl_sortingenv_b200.policy import SB3_KEYS, SHAPES, load_sb3_zip
    sd = {k: torch.randn(*s) for k, s in zip(SB3_KEYS, SHAPES)}
    sd["mlp_extractor.value_net.0.weight"] = torch.randn(32, 13)      # extra keys of a real archive are ignored
    sd["log_std"] = torch.zeros(1)
    buf = io.BytesIO(); torch.save(sd, buf)
    path = tmp_path / "sort_100000.zip"
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("policy.pth", buf.getvalue())
        z.writestr("data", "{}")
    w = load_sb3_zip(str(path))
    assert w.numel() == 1570 and torch.equal(w[416:448], sd[SB3_KEYS[1]]) and torch.equal(w[1504:1568].reshape(2, 32), sd[SB3_KEYS[4]])


def test_handle_free_entry_points_reject_bad_arguments_without_a_gpu():
    """msort_ppo_* and msort_rollout_pack take no handle: their argument checks run before any CUDA call, so they are testable
    here — wrong struct sizes, unsupported (obs_dim, num_actions), NULL buffers and misaligned pointers come back as error
    codes with a message, never as a crash."""
    import ctypes as C
    from marl_sortingenv_b200 import _abi
    lib = _abi.load_library()
    assert lib.msort_ppo_param_count(29, 22) == 2 * (32 * 29 + 32 + 32 * 32 + 32) + (22 * 32 + 22) + (32 + 1) == 4791
    assert lib.msort_ppo_param_count(13, 2) == 2 * (32 * 13 + 32 + 32 * 32 + 32) + (2 * 32 + 2) + 33
    assert lib.msort_ppo_scratch_floats(29, 22) > 2 * 4791 // 2 and lib.msort_ppo_scratch_floats(29, 11) == 0
    buf = (C.c_float * 64)()
    p = C.c_void_p((C.addressof(buf) + 15) // 16 * 16)                  # a 16-byte aligned address inside the buffer
    good = _abi.MsortPpoBatch(C.sizeof(_abi.MsortPpoBatch), 29, 22, 0, 16, p, p, p, p, p, p)
    hp = _abi.MsortPpoHparams(C.sizeof(_abi.MsortPpoHparams), 1, 0.2, 0.5, 0.05, 3e-4, 0.9, 0.999, 1e-5, 0.5)
    assert lib.msort_ppo_forward(None, p, p, p, None) == _abi.E_INVALID
    assert lib.msort_ppo_forward(C.byref(good), None, p, p, None) == _abi.E_INVALID
    bad = _abi.MsortPpoBatch(C.sizeof(_abi.MsortPpoBatch) - 4, 29, 22, 0, 16, p, p, p, p, p, p)
    assert lib.msort_ppo_forward(C.byref(bad), p, p, p, None) == _abi.E_INVALID
    dims = _abi.MsortPpoBatch(C.sizeof(_abi.MsortPpoBatch), 29, 11, 0, 16, p, p, p, p, p, p)
    assert lib.msort_ppo_forward(C.byref(dims), p, p, p, None) == _abi.E_UNSUPPORTED
    assert b"obs_dim" in lib.msort_last_error()
    noadv = _abi.MsortPpoBatch(C.sizeof(_abi.MsortPpoBatch), 29, 22, 0, 16, p, p, p, p, None, p)
    assert lib.msort_ppo_gradient(C.byref(noadv), C.byref(hp), p, p, None, 0, 16, p, None, None) == _abi.E_INVALID
    assert lib.msort_ppo_gradient(C.byref(good), C.byref(hp), p, p, None, 8, 16, p, None, None) == _abi.E_INVALID    # first + count > rows
    badhp = _abi.MsortPpoHparams(4, 1, 0.2, 0.5, 0.05, 3e-4, 0.9, 0.999, 1e-5, 0.5)
    assert lib.msort_ppo_gradient(C.byref(good), C.byref(badhp), p, p, None, 0, 16, p, None, None) == _abi.E_INVALID
    assert lib.msort_ppo_update(C.byref(good), C.byref(hp), p, p, p, p, p, None, 1, 8, p, None, None) == _abi.E_INVALID   # no permutations
    assert lib.msort_ppo_gae(0, 16, p, p, p, p, 0.99, 0.95, p, p, None) == _abi.E_INVALID
    assert lib.msort_rollout_pack(None, p, None) == _abi.E_INVALID
    assert lib.msort_rollout_pack(p, C.c_void_p(p.value + 4), None) == _abi.E_INVALID                                # 16-byte alignment
    for fn in ("msort_rollout_step", "msort_rollout_policy", "msort_policy_eval"):                                   # NULL handle
        args = [0 if t in (C.c_int, C.c_int64, C.c_uint32, C.c_uint64) else None for t in _abi.SYMBOLS[fn][1]]
        assert getattr(lib, fn)(*args) == _abi.E_INVALID, fn
