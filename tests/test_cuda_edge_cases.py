"""GPU edge cases of the C-ABI path: tiny and ragged batch sizes, partial and unseeded resets,
state export/import round trips, the state-statistics kernel, observe(), argument errors,
out-of-range actions, replay-stream under-runs."""
import ctypes as C

import numpy as np
import pytest

from parity_util import assert_float_close, config_for, state_rows

pytestmark = pytest.mark.gpu

META = dict(kind="mono", max_steps=30, noise=0.05, balesize=200, use_action_masking=True,
            check_overflow=False, auto_reset=True)


def _pair(n, kind="mono", seed=5, **over):
    from cuda_backend import CudaBackend
    from oracle.cpu_oracle import OracleEnv
    meta = dict(META, kind=kind, **over)
    return OracleEnv(config_for(meta, n, rng_mode="philox", seed=seed)), CudaBackend(config_for(meta, n, rng_mode="philox", seed=seed))


@pytest.mark.parametrize("n", [1, 2, 31, 127, 128, 129, 1000])
def test_ragged_batch_sizes(n):
    ora, gpu = _pair(n)
    ora.reset(); gpu.reset()
    for t in range(35):
        a = ora.sample_masked_actions(1, t)
        oo, orw, ot, om, _ = ora.step(a)
        go, grw, gt, gm, _ = gpu.step(a)
        assert np.array_equal(state_rows(ora.state), state_rows(gpu.export_state()))
        assert np.array_equal(om, gm) and np.array_equal(ot, gt)
        assert_float_close(go, oo, "obs"); assert_float_close(grw, orw, "reward")


def test_partial_and_unseeded_reset_semantics():
    import torch
    from marl_sortingenv_b200 import BatchedMonolithEnv
    n = 300
    env = BatchedMonolithEnv(n, max_steps=1000, seed=9, auto_reset=False)
    env.reset()
    for t in range(20):
        env.step(env.sample_actions(3, t))
    before = env.export_state()
    which = np.zeros(n, dtype=np.uint8); which[::3] = 1
    obs_before = env.obs.clone()
    env.reset(which=which)                       # unseeded (streams run on), selected envs only
    after = env.export_state()
    sel = which.astype(bool)
    assert np.array_equal(state_rows(after)[~sel], state_rows(before)[~sel])
    assert np.all(after["step"][sel] == 0) and np.all(after["episode"][sel] == before["episode"][sel] + 1)
    assert np.all(after["cont_true"][sel] == 0) and np.all(after["bale_n"][sel] == 0)
    assert np.all(after["acc_belt"][sel] == 0.75)
    assert torch.equal(env.obs[~torch.as_tensor(sel)], obs_before[~torch.as_tensor(sel)])
    assert torch.all(env.obs[torch.as_tensor(sel)][:, 5:9] == 0.75)
    assert bool(env.action_masks()[torch.as_tensor(sel)][:, 1:11].any()) is False
    env.reset(seed=9)                            # seeded: episode numbering restarts
    st = env.export_state()
    assert np.all(st["episode"] == 0) and np.all(st["step"] == 0)
    first = st["gen_first"].copy()
    env.reset(seed=9)
    assert np.array_equal(env.export_state()["gen_first"], first)     # deterministic in the seed
    assert 0.3 < (first == 1).mean() < 0.7                            # permutation([1,2]) is a fair coin


def test_export_import_round_trip_and_observe():
    import torch
    from marl_sortingenv_b200 import BatchedPressingEnv
    n = 517
    a_env = BatchedPressingEnv(n, max_steps=40, seed=21)
    a_env.reset()
    for t in range(25):
        a_env.step(a_env.sample_actions(2, t))
    snap = a_env.export_state()
    b_env = BatchedPressingEnv(n, max_steps=40, seed=21)
    b_env.import_state(snap)
    assert np.array_equal(state_rows(b_env.export_state()), state_rows(snap))
    assert torch.equal(b_env.get_obs(), a_env.obs) and torch.equal(b_env.action_masks(), a_env.action_masks())
    for t in range(25, 45):                      # both continue identically (PHILOX counters live in the state)
        act = a_env.sample_actions(2, t)
        a_env.step(act); b_env.step(act.clone())
    assert torch.equal(a_env.state[: 13 * 16 * 0 + a_env.state.numel()], a_env.state)
    assert np.array_equal(state_rows(a_env.export_state()), state_rows(b_env.export_state()))
    assert torch.equal(a_env.obs, b_env.obs)


def test_state_statistics_kernel():
    from marl_sortingenv_b200 import BatchedMonolithEnv
    n = 2222
    env = BatchedMonolithEnv(n, max_steps=60, seed=2)
    env.reset()
    for t in range(41):
        env.step(env.sample_actions(7, t))
    st = env.export_state()
    s = env.state_stats().cpu().numpy()
    assert s[0] == n
    assert s[1] == (st["cont_true"].sum() + st["cont_false"].sum() + st["cont_e"].sum())
    assert np.array_equal(s[2:7], st["bale_n"].sum(0)) and np.array_equal(s[7:12], st["bale_sum"].sum(0))
    assert s[13] == (st["press_timer"] > 0).sum() and s[15] == st["step"].sum()
    assert abs(s[14] - st["ep_return"].sum()) < 1e-6
    tot = st["cont_true"] + st["cont_false"]
    pur = np.where(tot > 0, np.rint(st["cont_true"] / np.maximum(tot, 1) * 100) / 100, 0.9)
    assert abs(s[12] - pur.mean(1).sum()) < 1e-6


def test_out_of_range_actions_are_clamped_and_counted():
    import torch
    from marl_sortingenv_b200 import BatchedSortingEnv
    env = BatchedSortingEnv(64, max_steps=10, seed=1)
    env.reset()
    a = torch.full((64,), 7, dtype=torch.int64, device="cuda"); a[:5] = -3
    _, _, _, _, info = env.step(a)
    assert int(env.stats[8].item()) == 64
    assert torch.all(info["action"][:5] == 0) and torch.all(info["action"][5:] == 1)


def test_argument_errors_are_reported_not_crashed():
    import torch
    from marl_sortingenv_b200 import BatchedMonolithEnv, _abi
    env = BatchedMonolithEnv(256, seed=1)
    env.reset()
    lib, h = env.lib, env._h
    a = torch.zeros(256, dtype=torch.int64, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.msort_step(h, p(env.state), None, p(env.obs), p(env.reward), p(env.terminated), p(env.mask), None, None, None)
    assert rc == _abi.E_INVALID and b"NULL" in lib.msort_last_error()
    rc = lib.msort_step(h, C.c_void_p(env.state.data_ptr() + 4), p(a), p(env.obs), p(env.reward), p(env.terminated),
                        p(env.mask), None, None, None)
    assert rc == _abi.E_INVALID and b"aligned" in lib.msort_last_error()
    rc = lib.msort_step(h, p(env.state), p(a), C.c_void_p(env.obs.data_ptr() + 4), p(env.reward), p(env.terminated),
                        p(env.mask), None, None, None)
    assert rc == _abi.E_INVALID and b"16-byte" in lib.msort_last_error()       # obs tiles leave by TMA bulk copies
    rc = lib.msort_step(h, p(env.state), p(a), p(env.obs), p(env.reward), p(env.terminated),
                        C.c_void_p(env.mask.data_ptr() + 4), None, None, None)
    assert rc == _abi.E_INVALID and b"16-byte" in lib.msort_last_error()
    rp = _abi.MsortReplay(); rp.struct_size = C.sizeof(rp)
    rc = lib.msort_step(h, p(env.state), p(a), p(env.obs), p(env.reward), p(env.terminated), p(env.mask), None,
                        C.byref(rp), None)
    assert rc == _abi.E_INVALID and b"PHILOX" in lib.msort_last_error()
    with pytest.raises(ValueError):
        env.step(torch.zeros(3, dtype=torch.int64, device="cuda"))
    env.sync_check()                              # nothing sticky was left behind


def test_replay_mode_requires_streams_and_reports_underrun():
    import torch
    from marl_sortingenv_b200 import BatchedSortingEnv, _abi
    env = BatchedSortingEnv(128, max_steps=50, seed=0, rng_mode="replay", noise_sorting=0.05)
    env.reset(first_pattern=np.ones(128, np.uint8))
    a = torch.zeros(128, dtype=torch.int64, device="cuda")
    with pytest.raises(_abi.MsortError) as e:
        env.step(a)
    assert e.value.code == _abi.E_REPLAY
    rng = np.random.default_rng(0)
    replay = dict(noise_u=rng.random((128, 4)), redis_u=rng.random((128, 3)), press_choice=np.zeros(128, np.uint8))
    for t in range(6):                            # 3 uniforms per env run out once real sorting starts (step 3)
        replay["noise_u"] = rng.random((128, 4))
        env.step(a, replay=replay)
    env.sync_check()
    assert env.stats[9].item() > 0                # under-runs are counted, the run stays defined
    st = env.export_state()
    total = (st["cont_true"].sum(1) + st["cont_false"].sum(1) + st["cont_e"] + st["press_n"].sum(1)
             + st["bale_sum"].sum(1) + st["input"].sum(1) + st["belt"].sum(1))
    assert np.all(total == 600)


def test_step_writes_into_caller_buffers():
    """step(out_obs=, out_mask=) writes the new observation / mask in place into caller tensors (zero-copy
    rollout buffers) and leaves everything else identical to the default call."""
    import torch
    import marl_sortingenv_b200 as ms
    n = 1000
    e1 = ms.BatchedMonolithEnv(n, max_steps=20, seed=5)
    e2 = ms.BatchedMonolithEnv(n, max_steps=20, seed=5)
    e1.reset(); e2.reset()
    buf_o = torch.zeros((8, n, e2.D), device="cuda")
    buf_m = torch.zeros((8, n, e2.A), dtype=torch.bool, device="cuda")
    for t in range(30):
        a = e1.sample_actions(3, t)
        o1, r1, d1, _, _ = e1.step(a)
        o2, r2, d2, _, _ = e2.step(a, out_obs=buf_o[t % 8], out_mask=buf_m[t % 8])
        assert o2.data_ptr() == buf_o[t % 8].data_ptr() and e2.action_masks().data_ptr() == buf_m[t % 8].data_ptr()
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
        assert torch.equal(e1.action_masks(), e2.action_masks())
    assert torch.equal(e1.state, e2.state)
    with pytest.raises(ValueError):
        e2.step(a, out_obs=torch.zeros((n, e2.D + 1), device="cuda"))


def test_persistent_env2_kernel_with_unaligned_actions_and_kernel_variants():
    """Env_2's persistent HOT kernel fetches a tile's actions by TMA only when the action tensor is 16-byte
    aligned; an 8-byte-aligned view takes the plain-load path.  Both, and the FAST kernel that runs when a
    per-step info array is requested, must leave the same state behind.  More tiles than resident CTAs, ragged."""
    import torch
    from marl_sortingenv_b200 import BatchedPressingEnv
    from marl_sortingenv_b200.policy import sb3_style_init
    n = 128 * 148 * 6 + 77
    from marl_sortingenv_b200 import _abi
    envs = [BatchedPressingEnv(n, max_steps=20, seed=21, info_level=lvl) for lvl in ("episode", "episode", "full", "episode")]
    for e in envs:
        e.set_sort_policy(sb3_style_init(4, action_gain=1.0))
        e.reset()
    envs[3].set_option(_abi.OPT_TENSOR_POLICY, 0)      # the per-thread FFMA2 policy instead of the tensor-core one
    buf = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    for t in range(45):
        a = envs[0].sample_actions(seed=2, t=t)
        buf[1:].copy_(a)
        assert buf[1:].data_ptr() % 16 == 8
        envs[0].step(a); envs[1].step(buf[1:]); envs[2].step(a); envs[3].step(a)
    tv = "hot_tensor_split" if envs[0].get_option(_abi.OPT_TENSOR_POLICY) == 2 else "hot_tensor"
    assert [e.step_variant for e in envs] == [tv, tv, "fast", "hot_persistent"]
    assert torch.equal(envs[0].state, envs[1].state) and torch.equal(envs[3].state, envs[2].state)
    assert torch.equal(envs[0].obs, envs[1].obs) and torch.equal(envs[3].obs, envs[2].obs)
    assert torch.equal(envs[3].mask, envs[2].mask) and torch.equal(envs[0].reward, envs[1].reward)
    # tensor-core vs fp32 policy: the argmax may differ only on a numerical tie of the two logits (tests/test_tc_mlp_gpu.py)
    diverged = (envs[0].state.view(torch.int32).reshape(13, -1, 4) != envs[3].state.view(torch.int32).reshape(13, -1, 4)).any(2).any(0)
    assert int(diverged.sum()) <= 3, f"{int(diverged.sum())} envs diverged between the tensor-core and the FFMA2 policy"


def test_imported_state_with_unusual_stages_is_stepped_exactly():
    """The HOT kernel relies on host-proved facts about what a stage can hold (a pattern batch or nothing: at most
    twelve mis-sorted units per station).  An imported state may hold anything — here a sorting stage of
    60/0/40/0 and 0/55/0/45 units — so the handle falls back to the FAST kernel (general draw loop) until the
    next full reset; the step must still match the oracle exactly."""
    from cuda_backend import CudaBackend
    from oracle.cpu_oracle import OracleEnv
    n = 600
    meta = dict(META, kind="mono", max_steps=30)
    ora = OracleEnv(config_for(meta, n, rng_mode="philox", seed=8))
    gpu = CudaBackend(config_for(meta, n, rng_mode="philox", seed=8), info_level="episode")
    ora.reset(); gpu.reset()
    for t in range(7):
        a = ora.sample_masked_actions(4, t)
        ora.step(a); gpu.step(a)
    assert gpu.env.step_variant == "hot"
    st = ora.state.copy()
    st["sorting"][0::2] = (60, 0, 40, 0)
    st["sorting"][1::2] = (0, 55, 0, 45)
    ora.state[:] = st
    gpu.env.import_state(st)
    for t in range(7, 20):
        a = ora.sample_masked_actions(4, t)
        oo, orw, ot, om, _ = ora.step(a)
        go, grw, gt, gm, _ = gpu.step(a)
        assert gpu.env.step_variant == "fast"
        assert np.array_equal(state_rows(ora.state), state_rows(gpu.export_state())), t
        assert np.array_equal(om, gm) and np.array_equal(ot, gt)
        assert_float_close(grw, orw, f"step {t} reward"); assert_float_close(go, oo, f"step {t} obs")
    gpu.reset()
    gpu.step(np.zeros(n, dtype=np.int64))
    assert gpu.env.step_variant == "hot"           # a full reset restores the proved facts
