"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI,
against (i) the committed reference trajectories and (ii) the CPU oracle on seeded inputs.

Bar (BASELINE.json north_star): integer state, masks, done flags bit-exact; obs / rewards
within 1e-5 relative (absolute floor 1e-7 where the reference value is 0)."""
import numpy as np
import pytest

from parity_util import (FLOAT_ATOL, FLOAT_RTOL, assert_float_close, config_dict_for, config_for, golden_group,
                         golden_group_names, replay_and_compare, state_rows)

pytestmark = pytest.mark.gpu

NAMES = golden_group_names()


@pytest.fixture(scope="module")
def cuda_backend():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from cuda_backend import CudaBackend
    return CudaBackend


@pytest.mark.parametrize("name", NAMES)
def test_cuda_replays_reference(cuda_backend, name):
    """REPLAY mode on the recorded numpy streams of the unmodified reference."""
    meta, batch = golden_group(name)
    n = batch["action"].shape[1]
    env = cuda_backend(config_for(meta, n), config=config_dict_for(meta))
    if meta.get("mlp"):
        env.set_policy(batch["mlp_weights"])
    use_counts = not name.startswith("gen_")
    steps = replay_and_compare(env, batch, meta, use_input_counts=use_counts)
    assert steps == meta["steps"] * n


def test_cuda_embedded_mlp_matches_reference_argmax(cuda_backend):
    meta, batch = golden_group("mlp_press")
    env = cuda_backend(config_for(meta, batch["action"].shape[1]))
    env.set_policy(batch["mlp_weights"])
    replay_and_compare(env, batch, meta, check_mlp=True)


def _philox_cross_check(cuda_backend, kind, n, T, *, masking=True, overflow=False, noise=0.05,
                        max_steps=50, balesize=200, seed=42, offset=0, mlp=None, auto_reset=True, config=None):
    """PHILOX mode: CUDA and the oracle share the counter scheme, so everything must agree."""
    from oracle.cpu_oracle import OracleEnv
    meta = dict(kind=kind, max_steps=max_steps, noise=noise, balesize=balesize,
                use_action_masking=masking, check_overflow=overflow, auto_reset=auto_reset, mlp=mlp is not None)
    cfg_o = config_for(meta, n, rng_mode="philox", seed=seed, global_env_offset=offset, config=config)
    cfg_c = config_for(meta, n, rng_mode="philox", seed=seed, global_env_offset=offset, config=config)
    ora, gpu = OracleEnv(cfg_o, nthreads=8), cuda_backend(cfg_c, config=config)
    if mlp is not None:
        ora.set_policy(mlp); gpu.set_policy(mlp)
    o0, m0 = ora.reset()
    g0, gm0 = gpu.reset()
    assert np.array_equal(o0, g0) and np.array_equal(m0, gm0)
    A = m0.shape[1]
    rng = np.random.default_rng(1234)
    n_done = 0
    for t in range(T):
        if masking:
            a = ora.sample_masked_actions(99, t)
        else:
            a = rng.integers(0, A, size=n)
        oo, orw, ot, om, oi = ora.step(a)
        go, grw, gt, gm, gi = gpu.step(a)
        so, sg = state_rows(ora.state), state_rows(gpu.export_state())
        if not np.array_equal(so, sg):
            bad = np.argwhere(so != sg)[0]
            raise AssertionError(f"{kind} step {t}: state mismatch env {bad[0]} col {bad[1]}: oracle {so[bad[0]]} cuda {sg[bad[0]]}")
        for f in ("gen_first", "gen_idx", "gen_counter", "episode", "sensor_mode"):
            assert np.array_equal(ora.state[f], gpu.export_state()[f]), f
        assert np.array_equal(ot, gt), f"step {t}: terminated"
        assert np.array_equal(om, gm), f"step {t}: mask"
        for k in ("overflow", "overflow_material", "sort_mode", "press_action", "invalid_action", "action", "sorted_true"):
            assert np.array_equal(np.asarray(oi[k]).astype(np.int64), np.asarray(gi[k]).astype(np.int64)), f"step {t}: {k}"
        assert_float_close(grw, orw, f"step {t}: reward")
        assert_float_close(go, oo, f"step {t}: obs")
        assert_float_close(gi["reward_sort"], oi["reward_sort"], f"step {t}: reward_sort")
        assert_float_close(gi["reward_press"], oi["reward_press"], f"step {t}: reward_press")
        if ot.any():
            assert_float_close(gi["terminal_obs"][ot], oi["terminal_obs"][ot], "terminal obs")
            assert_float_close(gi["episode_return"][ot], oi["episode_return"][ot], "episode return")
            assert np.array_equal(gi["episode_length"][ot], oi["episode_length"][ot])
            n_done += int(ot.sum())
    assert_float_close(gpu.env.stats.cpu().numpy()[:10], ora.stats[:10], "stats accumulators", rtol=1e-9)
    return n_done


@pytest.mark.parametrize("kind", ["sort", "press", "mono"])
@pytest.mark.parametrize("masking", [True, False])
def test_cuda_matches_oracle_philox(cuda_backend, kind, masking):
    done = _philox_cross_check(cuda_backend, kind, 4096 + 37, 110, masking=masking)
    assert done >= 2 * 4096


def test_cuda_matches_oracle_philox_overflow_and_offsets(cuda_backend):
    _philox_cross_check(cuda_backend, "mono", 1000, 120, masking=False, overflow=True, offset=123456789012)
    _philox_cross_check(cuda_backend, "press", 777, 80, masking=True, overflow=True, noise=0.0, max_steps=33, balesize=150)


def _variant_config(changes):
    """config.yml with some keys replaced: {(section, key): value}."""
    import copy
    from marl_sortingenv_b200.config import load_config
    cfg = copy.deepcopy(load_config(None))
    for (sec, key), val in changes.items():
        cfg[sec][key] = val
    return cfg


@pytest.mark.parametrize("kind", ["sort", "press", "mono"])
def test_cuda_matches_oracle_philox_generic_instantiation(cuda_backend, kind):
    """The step kernel has a FAST instantiation for configs where the host can prove that boosted
    accuracies clip to 1.0 (default config.yml) and a generic one.  These configs force the generic
    one: a boost that does not saturate (boosted stations mis-sort too, accuracies really clip),
    and a batch larger than the packed-byte class selection allows."""
    weak = _variant_config({("sorting_station", "boost"): 0.22})
    done = _philox_cross_check(cuda_backend, kind, 2048 + 5, 70, config=weak, noise=0.08)
    assert done >= 2048
    big = _variant_config({("simulation", "input_batch_size"): 200})
    _philox_cross_check(cuda_backend, kind, 1024 + 3, 60, config=big, masking=False)


@pytest.mark.parametrize("config_kind", ["default", "weak_boost"])
def test_cuda_matches_oracle_philox_wide_layout(cuda_backend, config_kind):
    """Without auto-reset the state keeps 32-bit container counts (LAYOUT_WIDE): FAST and generic."""
    cfg = None if config_kind == "default" else _variant_config({("sorting_station", "boost"): 0.22})
    _philox_cross_check(cuda_backend, "mono", 1500, 90, max_steps=400, auto_reset=False, config=cfg)
    _philox_cross_check(cuda_backend, "sort", 700, 60, max_steps=400, auto_reset=False, config=cfg, masking=False)


TIE_BAND = 1e-5      # parity contract for the embedded policy: identical argmax except where |logit0 - logit1| < 1e-5


def _hot_cross_check(cuda_backend, kind, n, T, *, mlp=None, seed=11, max_steps=25, variant=None):
    """The HOT instantiation of the step kernel — the one bench.py times: FAST config, compact layout, action
    masking + auto-reset, mask output, NO per-step info arrays (`info_level='episode'`), and for Env_2 the
    persistent TMA-staged form — against the oracle: exported state bit-exact, masks / terminated equal,
    obs / reward / terminal observation / episode return within the float tolerance, statistics equal.
    With an embedded policy (Env_2) the tensor-core kernel may pick the other sort mode only inside the
    contract's tie band (the oracle reports its fp32 margin); such an env is re-synchronised and counted."""
    from oracle.cpu_oracle import OracleEnv
    meta = dict(kind=kind, max_steps=max_steps, noise=0.05, balesize=200, use_action_masking=True,
                check_overflow=False, auto_reset=True, mlp=mlp is not None)
    ora = OracleEnv(config_for(meta, n, rng_mode="philox", seed=seed), nthreads=8)
    gpu = cuda_backend(config_for(meta, n, rng_mode="philox", seed=seed), info_level="episode")
    if mlp is not None:
        ora.set_policy(mlp); gpu.set_policy(mlp)
    if variant is None:
        variant = (("hot_tensor", "hot_tensor_split") if mlp is not None else "hot_persistent") if kind == "press" else "hot"
    o0, m0 = ora.reset()
    g0, gm0 = gpu.reset()
    assert np.array_equal(o0, g0) and np.array_equal(m0, gm0)
    n_done = n_flips = 0
    for t in range(T):
        a = ora.sample_masked_actions(5, t)
        oo, orw, ot, om, oi = ora.step(a)
        go, grw, gt, gm, gi = gpu.step(a)
        assert set(gi) == {"terminal_obs", "episode_return", "episode_length"}      # no per-step info arrays ...
        assert gpu.env.step_variant == variant or gpu.env.step_variant in variant, gpu.env.step_variant   # ... so the HOT kernel ran
        gstate = gpu.export_state()
        ok = np.ones(n, dtype=bool)
        if mlp is not None:
            flipped = ora.state["sensor_mode"] != gstate["sensor_mode"]
            if flipped.any():
                worst = float(np.max(oi["mlp_margin"][flipped]))
                assert worst < TIE_BAND, f"step {t}: sort-policy argmax differs at an fp32 margin of {worst:g} (band {TIE_BAND:g})"
                ora.state[flipped] = gstate[flipped]          # both sides continue from the same plant
                n_flips += int(flipped.sum()); ok = ~flipped
        so, sg = state_rows(ora.state), state_rows(gstate)
        if not np.array_equal(so, sg):
            bad = np.argwhere(so != sg)[0]
            raise AssertionError(f"{kind} step {t}: state mismatch env {bad[0]} col {bad[1]}: oracle {so[bad[0]]} cuda {sg[bad[0]]}")
        assert np.array_equal(ot, gt), f"step {t}: terminated"
        assert np.array_equal(om, gm), f"step {t}: mask"
        assert_float_close(grw[ok], orw[ok], f"step {t}: reward")
        assert_float_close(go[ok], oo[ok], f"step {t}: obs")
        if ot.any():
            assert_float_close(gi["terminal_obs"][ot], oi["terminal_obs"][ot], "terminal obs")
            assert_float_close(gi["episode_return"][ot], oi["episode_return"][ot], "episode return")
            assert np.array_equal(gi["episode_length"][ot], oi["episode_length"][ot])
            n_done += int(ot.sum())
    assert_float_close(gpu.env.stats.cpu().numpy()[:10], ora.stats[:10], "stats accumulators", rtol=1e-9)
    if mlp is not None:
        print(f"embedded policy [{variant}]: {n_flips} tie-band flips in {n * T} env-steps")
        assert n_flips <= max(3, n * T // 20000), f"{n_flips} tie-band flips in {n * T} env-steps"
    return n_done


@pytest.mark.parametrize("kind", ["sort", "press", "mono"])
def test_cuda_hot_kernel_matches_oracle(cuda_backend, kind):
    """Ragged size (partial last tile), two episodes."""
    assert _hot_cross_check(cuda_backend, kind, 4096 + 37, 60) >= 2 * 4096


@pytest.mark.parametrize("kind", ["sort", "press", "mono"])
def test_cuda_hot_kernel_matches_oracle_long_episodes(cuda_backend, kind):
    """batch * max_steps > 8192: the HOT instantiation without the small-level shortcut (container levels up to
    12 000 here; the reference's benchmark episodes are 200 steps long)."""
    assert _hot_cross_check(cuda_backend, kind, 1024 + 9, 125, max_steps=120) >= 1024


def test_cuda_hot_kernel_matches_oracle_more_than_one_wave(cuda_backend):
    """More tiles than resident CTAs, so Env_2's persistent kernel really loops (its second and later tiles come
    from the TMA-staged buffer) and Env_3's L2 prefetch reaches real tiles; ragged last tile; Env_2 with the
    embedded policy evaluated on the tensor cores (whole tiles) and by the FFMA2 kernel (the ragged tail)."""
    from marl_sortingenv_b200.policy import sb3_style_init
    n = 128 * 148 * 8 + 128 + 5
    _hot_cross_check(cuda_backend, "press", n, 30, mlp=sb3_style_init(3, action_gain=1.0).numpy())
    _hot_cross_check(cuda_backend, "mono", n, 30)


def test_cuda_matches_oracle_philox_mlp(cuda_backend):
    from marl_sortingenv_b200.policy import sb3_style_init
    w = sb3_style_init(3, action_gain=1.0).numpy()
    _philox_cross_check(cuda_backend, "press", 2048, 60, mlp=w)


@pytest.mark.parametrize("info_level", ["episode", "full"])
@pytest.mark.parametrize("name", ["philox_sort", "philox_press", "philox_mono", "philox_mono_unmasked"])
def test_cuda_philox_matches_reference(cuda_backend, name, info_level):
    """The PRODUCTION instantiations of the step kernel — HOT (`info_level='episode'`: what bench.py times) and FAST
    (`'full'`) with the Philox generator — against the UNMODIFIED reference stepped on the very random inputs that
    generator produces (tests/golden/reference_philox.npz, made by tests/golden/make_philox_golden.py through
    oracle/ref_drive.py): integer state, masks, flags bit-exact; obs / reward within 1e-5."""
    from parity_util import compare_with_philox_reference, philox_golden
    meta, batch = philox_golden()[name]
    cfg = config_for(meta, meta["envs"], rng_mode="philox", seed=meta["seed"], global_env_offset=meta["offset"])
    gpu = cuda_backend(cfg, info_level=info_level)
    assert compare_with_philox_reference(gpu, meta, batch) == meta["envs"] * meta["steps"]
    hot = info_level == "episode" and meta["use_action_masking"] and not meta["check_overflow"]
    want = ("hot_persistent" if meta["kind"] == "press" else "hot") if hot else "fast"
    assert gpu.env.step_variant == want, gpu.env.step_variant


def test_config2_env1_65536_replay_bit_exact(cuda_backend):
    """BASELINE config 2: Env_1_Sorting, 65 536 envs, env i seeded 1+i, random actions, replayed
    numpy streams (generated here with numpy exactly as the reference's generators produce
    them) — CUDA vs oracle, all T steps, integer state bit-exact."""
    from oracle.cpu_oracle import OracleEnv
    n, T, L = 65536, 50, 23 * 50
    meta = dict(kind="sort", max_steps=50, noise=0.05, balesize=200, use_action_masking=True,
                check_overflow=False, auto_reset=True)
    noise = np.empty((T, n, 4)); redis = np.empty((n, L)); first = np.empty(n, dtype=np.uint8)
    for i in range(n):
        s = 1 + i
        noise[:, i, :] = np.random.default_rng(s + 4).random((T, 4))
        redis[i] = np.random.default_rng(s + 99).random(L)
        first[i] = np.random.default_rng(s).permutation([1, 2])[0]
    import torch
    actions = torch.randint(0, 2, (T, n), generator=torch.Generator().manual_seed(1234)).numpy()
    ora, gpu = OracleEnv(config_for(meta, n), nthreads=8), cuda_backend(config_for(meta, n))
    ora.reset(first_pattern=first); gpu.reset(first_pattern=first)
    for t in range(T):
        # Env_1's internal press action is a recorded stream in REPLAY mode; here the oracle's
        # Philox sampler stands in for the recording and is fed to both sides.
        pc = _masked_press_choice(ora, t)
        oo, orw, ot, om, oi = ora.step(actions[t], noise_u=noise[t], redis_u=redis, press_choice=pc)
        go, grw, gt, gm, gi = gpu.step(actions[t], noise_u=noise[t], redis_u=redis, press_choice=pc)
        assert np.array_equal(state_rows(ora.state), state_rows(gpu.export_state())), f"step {t}"
        assert np.array_equal(ot, gt) and np.array_equal(om, gm)
        assert_float_close(grw, orw, f"step {t}: reward")
        assert_float_close(go, oo, f"step {t}: obs")
    assert ot.all()


def _masked_press_choice(ora, t):
    """A valid press action per env from the oracle's current state (post-sort levels are not
    known before the step, so choose among actions valid under ANY level: the no-op or a press
    that is idle and whose container is already >= one bale)."""
    from marl_sortingenv_b200 import _abi
    import ctypes as C
    saved = ora.cfg.env_kind
    ora.cfg.env_kind = _abi.ENV_PRESS          # reuse the 11-action masked sampler
    try:
        a = ora.sample_masked_actions(555, t)
    finally:
        ora.cfg.env_kind = saved
    return a.astype(np.uint8)


@pytest.mark.parametrize("n", [1 << 20, 1 << 23])
def test_full_size_properties_config4(cuda_backend, n):
    """BASELINE config 4 size (1 048 576 Monolith envs, masked) and config 5's total size on one GPU (8 388 608
    envs, 1.7 GB of state): size-independent properties — material conservation, mask consistency, step
    counters, reward bounds, determinism."""
    import torch
    from marl_sortingenv_b200 import BatchedMonolithEnv
    env = BatchedMonolithEnv(n, max_steps=50, seed=7)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    rewards = []
    for t in range(60):
        m = env.action_masks().clone()      # step() overwrites the mask buffer in place
        a = torch.multinomial(m.float(), 1, generator=g).squeeze(1)
        obs, r, term, trunc, info = env.step(a)
        assert bool(m.gather(1, a[:, None]).all())
        assert torch.isfinite(obs).all() and float(obs.min()) >= -1.0 and float(obs.max()) <= 1.0
        assert float(r.min()) >= -2.0 - 1e-6 and float(r.max()) <= 2.0 + 1e-6
        assert int(term.sum()) == (n if (t + 1) % 50 == 0 else 0)
        rewards.append(r.double().sum().item())
    st = env.export_state()
    total = (st["cont_true"].sum(1) + st["cont_false"].sum(1) + st["cont_e"] + st["press_n"].sum(1)
             + st["bale_sum"].sum(1) + st["input"].sum(1) + st["belt"].sum(1))
    assert np.all(total == 100 * st["step"]) and np.all(st["step"] == 10) and np.all(st["episode"] == 1)
    s = env.stats.cpu().numpy()
    assert s[0] == n and s[3] == 60 * n and abs(s[4] - sum(rewards)) < 1e-6 * max(1.0, abs(s[4]))
    # determinism: same seed, same actions -> identical state
    env2 = BatchedMonolithEnv(n, max_steps=50, seed=7)
    env2.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    for t in range(60):
        a = torch.multinomial(env2.action_masks().float(), 1, generator=g).squeeze(1)
        env2.step(a)
    assert torch.equal(env.state, env2.state)
