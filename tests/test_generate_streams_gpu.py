"""K3, the generator stream export (`msort_generate_streams`; ref: SeasonalInputGenerator.generate_input
input_generator.py:37-64, rng_noise.uniform env_super.py:508): what it writes is what the PHILOX step kernel consumes,
and a PHILOX trajectory replays bit-exactly through the REPLAY instantiation when fed those streams (plus the
state-dependent redistribution uniforms the oracle records)."""
import numpy as np
import pytest

from parity_util import config_for, state_rows

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["sort", "mono"])
def test_philox_trajectory_replays_through_the_replay_kernel(kind):
    import torch
    from cuda_backend import CudaBackend
    from oracle.cpu_oracle import REC_CAP, OracleEnv
    n, T = 128 * 5 + 33, 45
    meta = dict(kind=kind, max_steps=T, noise=0.05, balesize=200, use_action_masking=True, check_overflow=False, auto_reset=False)
    seed, off = 31, 777
    ora = OracleEnv(config_for(meta, n, rng_mode="philox", seed=seed, global_env_offset=off), nthreads=4)
    phi = CudaBackend(config_for(meta, n, rng_mode="philox", seed=seed, global_env_offset=off))
    rep = CudaBackend(config_for(meta, n, rng_mode="replay", seed=seed, global_env_offset=off))
    ora.reset(); phi.reset()
    st = phi.env.generate_streams(episode=0, first_step=0, num_steps=T, draw_words=True)
    torch.cuda.synchronize()
    fp = st["first_pattern"].cpu().numpy()
    assert np.array_equal(fp, ora.state["gen_first"].astype(np.uint8))
    rep.reset(first_pattern=fp)
    counts, noise = st["input_counts"].cpu().numpy().view(np.uint32), st["noise_u"].cpu().numpy()
    redis, cur = np.zeros((n, REC_CAP * T)), np.zeros(n, dtype=np.int64)
    for t in range(T):
        a = ora.sample_masked_actions(9, t)
        _, _, _, _, oi = ora.step(a, record=True)
        # K3 wrote exactly what the step consumed
        assert np.array_equal(counts[t], oi["rec_input_counts"]), f"step {t}: input batch"
        assert np.array_equal(noise[t], oi["rec_noise_u"]), f"step {t}: noise uniforms"
        for i in range(n):
            k = int(oi["rec_n_draws"][i])
            redis[i, cur[i]:cur[i] + k] = oi["rec_redis_u"][i, :k]
            cur[i] += k
        po, prw, pt, pm, _ = phi.step(a)
        ro, rrw, rt, rm, _ = rep.step(a, noise_u=noise[t], redis_u=redis.copy(), input_counts=counts[t],   # (a fresh array: the backend caches the upload by identity)
                                      press_choice=oi["rec_press_choice"])
        sp, sr, so = state_rows(phi.export_state()), state_rows(rep.export_state()), state_rows(ora.state)
        assert np.array_equal(sp, so), f"step {t}: PHILOX kernel vs oracle"
        assert np.array_equal(sr, sp), f"step {t}: REPLAY kernel on the exported streams vs PHILOX kernel"
        assert np.array_equal(pm, rm) and np.array_equal(pt, rt)
        assert np.allclose(po, ro, rtol=1e-5, atol=1e-7) and np.allclose(prw, rrw, rtol=1e-5, atol=1e-7)
    assert phi.env.step_variant == "fast" and rep.env.step_variant == "replay"


def test_generate_streams_later_episode_and_offsets():
    """Episode 2, steps 7..: the batches follow the 20-step pattern rule from the episode's own pattern order."""
    import torch
    import marl_sortingenv_b200 as ms
    env = ms.BatchedMonolithEnv(300, max_steps=50, seed=5, global_env_offset=1 << 33)
    env.reset()
    st = env.generate_streams(episode=2, first_step=7, num_steps=30)
    torch.cuda.synchronize()
    fp = st["first_pattern"].cpu().numpy().astype(int)
    counts = st["input_counts"].cpu().numpy().view(np.uint32)
    assert set(np.unique(fp)) <= {1, 2} and 0.3 < (fp == 1).mean() < 0.7
    pat = {1: 40 | 15 << 8 | 35 << 16 | 10 << 24, 2: 15 | 40 << 8 | 10 << 16 | 35 << 24}
    for k in range(30):
        t = 7 + k
        want = np.where(((t // 20) % 2 == 0) == (fp == 1), pat[1], pat[2])
        assert np.array_equal(counts[k], want.astype(np.uint32)), k
    u = st["noise_u"].cpu().numpy()
    assert u.shape == (30, 300, 4) and (u >= 0).all() and (u < 1).all() and abs(u.mean() - 0.5) < 0.01
    env.close()
