"""GPU-resident MaskablePPO-style loop over the batched env — a short training
run must beat the reference's published Rule-Based return (44.03) on the reference's protocol."""
import pytest

pytestmark = pytest.mark.gpu


def test_ppo_learns_on_device():
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200.ppo import MaskablePPO, evaluate_policy
    env = ms.BatchedMonolithEnv(1024, max_steps=200, seed=42, noise_sorting=0.0, info_level="none", track_stats=False)
    model = MaskablePPO(env, n_steps=64, batch_size=16384, n_epochs=10, seed=0)
    before, _ = evaluate_policy(model, ms.BatchedMonolithEnv, n_envs=512, noise_sorting=0.0)
    model.learn(5_000_000)
    after, std = evaluate_policy(model, ms.BatchedMonolithEnv, n_envs=512, noise_sorting=0.0)
    assert after > 44.03 and after > before + 50, (before, after, std)
    a, _ = model.predict(env.obs[0].cpu().numpy(), action_masks=env.action_masks()[0].cpu().numpy())
    assert 0 <= a < 22


def test_two_stream_rollout_fills_the_same_buffers():
    """collect_rollout on two streams over two env ranges (the default from 131 072 envs) leaves exactly the buffers,
    advantages and env state the single-stream rollout leaves: ranges are independent and every draw is keyed by the
    global env id."""
    import torch
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200.ppo import MaskablePPO
    n = 131072 + 128 * 3 + 17
    runs = []
    for streams in (1, 2):
        env = ms.BatchedMonolithEnv(n, max_steps=20, seed=4, info_level="none", track_stats=False)
        model = MaskablePPO(env, n_steps=24, seed=7, rollout_streams=streams)
        assert len(model._ranges) == streams
        adv, ret = model.collect_rollout()
        adv2, _ = model.collect_rollout()                  # a second rollout continues from the first one's last obs
        torch.cuda.synchronize()
        runs.append((model.buf, adv, ret, adv2, env.state.clone()))
    (b1, a1, r1, a1b, s1), (b2, a2, r2, a2b, s2) = runs
    for k in b1:
        assert torch.equal(b1[k], b2[k]), k
    assert torch.equal(a1, a2) and torch.equal(r1, r2) and torch.equal(a1b, a2b) and torch.equal(s1, s2)
