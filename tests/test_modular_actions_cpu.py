"""Host logic of the Env_3 mode='model' action source (ppo.modular_actions; ref env_monolith.py:186-221) against
the fixture recorded from the unmodified reference (tests/golden/make_model_mode_golden.py).  The env is a
stand-in fed from the fixture, so this runs without a GPU; tests/test_api_surfaces.py repeats it on the device,
where the agents' observation comes from msort_observe_after_shift."""
import numpy as np
import torch

from marl_sortingenv_b200.ppo import modular_actions
from parity_util import TorchPressStub, TorchSortStub, model_mode_golden


class _FixtureEnv:
    def __init__(self, batch):
        self.b, self.t = batch, 0
        self.num_envs, self.device = batch["action"].shape[1], torch.device("cpu")

    def observe_after_shift(self):
        return torch.as_tensor(self.b["agent_obs"][self.t])

    def action_masks(self):                       # the mask before step t is the one step t-1 left (reset: no-op only)
        if self.t == 0:
            m = np.zeros((self.num_envs, 22), dtype=bool)
            m[:, [0, 11]] = True
            return torch.as_tensor(m)
        return torch.as_tensor(self.b["mask"][self.t - 1])


def test_fixture_observations_are_the_shifted_plant():
    """What the reference showed its agents = the previous state with input -> belt -> sorting applied."""
    from parity_util import STATE_FIELDS
    meta, b = model_mode_golden()
    col, o = {}, 0
    for name, w in STATE_FIELDS:
        col[name] = slice(o, o + w); o += w
    for t in range(1, meta["steps"]):
        prev = b["state"][t - 1]
        inp, belt = prev[:, col["input"]].astype(np.float64), prev[:, col["belt"]].astype(np.float64)
        seen = b["agent_obs"][t]
        assert np.allclose(seen[:, 0], np.minimum(inp.sum(1) / 100.0, 1.0), atol=1e-6)
        assert np.allclose(seen[:, 1:5], inp / np.maximum(inp.sum(1, keepdims=True), 1), atol=1e-6)
        assert np.allclose(seen[:, 23:27], belt / 100.0, atol=1e-6)


def test_modular_actions_compose_the_reference_action():
    meta, b = model_mode_golden()
    env = _FixtureEnv(b)
    for t in range(meta["steps"]):
        env.t = t
        a = modular_actions(env, TorchSortStub(), TorchPressStub(), use_action_masking=True)
        assert np.array_equal(a.numpy(), b["action"][t]), f"step {t}"
    assert len(np.unique(b["action"])) > 10


def test_modular_actions_fallbacks_stay_valid():
    """A missing agent falls back to a uniform draw (env_monolith.py:194,213-219); with masking the press draw is valid."""
    meta, b = model_mode_golden()
    env = _FixtureEnv(b)
    for t in (0, 57, 150):
        env.t = t
        a = modular_actions(env, None, None, use_action_masking=True, seed=3, t=t)
        assert bool(env.action_masks().gather(1, a[:, None]).all())
        a = modular_actions(env, TorchSortStub(), None, use_action_masking=False, seed=3, t=t)
        assert int(a.min()) >= 0 and int(a.max()) < 22
