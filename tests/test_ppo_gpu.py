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
