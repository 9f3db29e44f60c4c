"""PHILOX (generator) mode has different random bits than numpy, so it is validated
statistically against anchors measured on the unmodified reference (SURVEY.md Appendix C:
2 000 episodes per env, max_steps=50, noise 0.05, actions uniform over the valid mask)."""
import numpy as np
import pytest

from oracle.cpu_oracle import OracleEnv
from parity_util import config_for

# (mean, standard error) measured on the reference; sd of the episode return in comment
ANCHORS = {
    "sort": dict(ret=(14.282, 0.035), bales=(8.415, 0.035), level=(2013.2, 8.9), e=(193.3, 3.1)),   # sd 1.59
    "press": dict(ret=(-12.715, 0.122), bales=(9.047, 0.035), level=(1831.4, 8.8), e=(192.5, 1.4)),  # sd 5.44
    "mono": dict(ret=(1.390, 0.133), bales=(8.636, 0.036), level=(1968.2, 9.1), e=(196.0, 3.1)),     # sd 5.96
}


def run_episodes(make_env, kind, n=6000, seed=2024):
    meta = dict(kind=kind, max_steps=50, noise=0.05, balesize=200, use_action_masking=True,
                check_overflow=False, auto_reset=False)
    env = make_env(config_for(meta, n, rng_mode="philox", seed=seed))
    env.reset()
    ret = np.zeros(n)
    for t in range(50):
        a = sample_uniform_valid(env, kind, t)
        _, r, term, _, _ = env.step(a)
        ret += r
    st = env.export_state() if hasattr(env, "export_state") else env.state
    assert term.all()
    level = st["cont_true"].sum(1) + st["cont_false"].sum(1) + st["cont_e"]
    return dict(ret=ret, bales=st["bale_n"].sum(1), level=level, e=st["cont_e"])


_rng = np.random.default_rng(77)


def sample_uniform_valid(env, kind, t):
    st = env.export_state() if hasattr(env, "export_state") else env.state
    n = len(st)
    if kind == "sort":
        return _rng.integers(0, 2, size=n)
    lvl = np.concatenate([st["cont_true"] + st["cont_false"], st["cont_e"][:, None]], 1) >= 200
    m = np.zeros((n, 11), bool); m[:, 0] = True
    m[:, 1:6] = lvl & (st["press_timer"][:, :1] == 0)
    m[:, 6:11] = lvl & (st["press_timer"][:, 1:2] == 0)
    if kind == "mono":
        m = np.concatenate([m, m], 1)
    u = _rng.random(n)
    cnt = m.sum(1)
    pick = np.minimum((u * cnt).astype(np.int64), cnt - 1)
    order = np.cumsum(m, 1) - 1
    return np.argmax(m & (order == pick[:, None]), 1)


def check(kind, out, n):
    for key, (mu, se_ref) in ANCHORS[kind].items():
        x = np.asarray(out[key], dtype=np.float64)
        se = np.hypot(x.std(ddof=1) / np.sqrt(n), se_ref)
        assert abs(x.mean() - mu) < 4.5 * se, f"{kind} {key}: {x.mean():.3f} vs reference {mu} (se {se:.3f})"


@pytest.mark.parametrize("kind", ["sort", "press", "mono"])
def test_oracle_philox_matches_reference_statistics(kind):
    n = 6000
    check(kind, run_episodes(lambda cfg: OracleEnv(cfg, nthreads=8), kind, n), n)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sort", "press", "mono"])
def test_cuda_philox_matches_reference_statistics(kind):
    from cuda_backend import CudaBackend
    n = 20000
    check(kind, run_episodes(lambda cfg: CudaBackend(cfg), kind, n, seed=99), n)
