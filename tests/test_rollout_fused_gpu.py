"""SURVEY.md §8f item 1, the rollout half as ONE kernel per env-step: `msort_rollout_step` = the fused step() plus, on the
observation / mask tile still in shared memory, the actor-critic forward + masked categorical draw of the NEXT step
(tcgen05).  ref: the MaskablePPO rollout loop, training.py:118-143 (sb3 collect_rollouts) over Env_3_Monolith.step
(env_monolith.py:109-284).  The step half must be bit-identical to `msort_step`; the policy half is compared with a
PyTorch fp32 evaluation of the same towers (tolerance 5e-3 on log-prob / value: fp16 operands, MUFU tanh) and with the
stand-alone policy kernel's draw."""
import pytest

pytestmark = pytest.mark.gpu


def _pair(n, seed=5, max_steps=12, scale=1.0):
    import torch
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200.ppo import MaskableActorCritic, flatten_parameters, pack_actor_critic
    envs = [ms.BatchedMonolithEnv(n, max_steps=max_steps, seed=seed, info_level="episode") for _ in range(2)]
    torch.manual_seed(seed)
    pol = MaskableActorCritic(29, 22).cuda()
    with torch.no_grad():
        for p in pol.parameters():                     # away from SB3's tiny-gain init: logits that decide the draw
            p.mul_(scale).add_(0.15 * scale * torch.randn_like(p))
    flat = flatten_parameters(pol)
    return envs, pol, flat, pack_actor_critic(pol)


@pytest.mark.parametrize("n", [128 * 6, 128 * 5 + 37, 100])
def test_fused_step_is_step_plus_policy(n):
    import torch
    (fa, pl), pol, flat, packed = _pair(n)
    pf = fa.rollout_pack(flat)
    fa.reset(); pl.reset()
    a, _, _ = pl.policy_act(packed, seed=9, t=0)
    agree = total = 0
    for t in range(40):                                # max_steps 12: every env auto-resets three times
        nxt = (torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"))
        fo, fr, ft, _, _ = fa.rollout_step(a, pf, 9, t + 1, nxt)
        po, pr, pt, _, _ = pl.step(a)
        assert fa.step_variant == "hot_fused" and pl.step_variant == "hot"
        assert torch.equal(fa.state, pl.state), f"step {t}: env state"
        assert torch.equal(fo, po) and torch.equal(fr, pr) and torch.equal(ft, pt) and torch.equal(fa.mask, pl.mask), f"step {t}"
        na, nlp, nv = nxt
        assert bool(pl.mask.gather(1, na[:, None]).all()), "drawn action must be valid under the mask"
        with torch.no_grad():
            ref = torch.log_softmax(pol.masked_logits(po, pl.mask), dim=-1)
            assert torch.allclose(nlp, ref.gather(1, na[:, None]).squeeze(1), atol=5e-3), (nlp - ref.gather(1, na[:, None]).squeeze(1)).abs().max()
            assert torch.allclose(nv, pol.vf(po).squeeze(1), atol=5e-3, rtol=5e-3)
        sa, _, _ = pl.policy_act(packed, seed=9, t=t + 1)      # same Philox uniform, same inverse-CDF rule: equal away from cdf edges
        agree += int((sa == na).sum()); total += n
        a = na
    assert agree >= 0.99 * total, (agree, total)


def test_fused_deterministic_is_the_argmax():
    import torch
    n = 128 * 4
    (fa, _), pol, flat, packed = _pair(n, seed=8, scale=2.0)
    pf = fa.rollout_pack(flat)
    fa.reset()
    a, _, _ = fa.policy_act(packed, seed=1, t=0, deterministic=True)
    nxt = (torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"))
    for t in range(10):
        fa.rollout_step(a, pf, 1, t + 1, nxt, deterministic=True)
        with torch.no_grad():
            lg = pol.masked_logits(fa.obs, fa.mask)
        top2 = lg.topk(2, dim=-1).values
        clear = (top2[:, 0] - top2[:, 1]) > 2e-2
        assert torch.equal(nxt[0][clear], lg.argmax(-1)[clear])
        a = nxt[0].clone()


def test_configurations_outside_hot_are_refused_not_miscomputed():
    import torch
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200 import _abi
    (fa, _), pol, flat, packed = _pair(256)
    pf = fa.rollout_pack(flat)
    full = ms.BatchedMonolithEnv(256, max_steps=12, seed=5, info_level="full")      # per-step info arrays: not the HOT kernel
    full.reset()
    a, _, _ = full.policy_act(packed, seed=1, t=0)
    nxt = (torch.empty(256, dtype=torch.int64, device="cuda"), torch.empty(256, device="cuda"), torch.empty(256, device="cuda"))
    before = full.state.clone()
    with pytest.raises(_abi.MsortError) as e:
        full.rollout_step(a, pf, 1, 1, nxt)
    assert e.value.code == _abi.E_UNSUPPORTED and torch.equal(full.state, before)
    press = ms.BatchedPressingEnv(256, max_steps=12, seed=5, info_level="episode")
    press.reset()
    with pytest.raises(ValueError):
        press.rollout_step(torch.zeros(256, dtype=torch.int64, device="cuda"), pf, 1, 1, nxt)


@pytest.mark.parametrize("graph", [False, True])
def test_ppo_fused_rollout_buffers_replay_through_the_plain_step_kernel(graph):
    """collect_rollout with one fused kernel per env-step: feeding the recorded actions to a fresh env through the plain step
    kernel reproduces the recorded observations, masks, rewards and dones bit for bit (three rollouts: eager, captured, replayed)."""
    import torch
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200.ppo import MaskablePPO
    n, T = 128 * 9, 12
    env = ms.BatchedMonolithEnv(n, max_steps=20, seed=21, info_level="none", track_stats=False)
    model = MaskablePPO(env, n_steps=T, seed=3, graph_rollout=graph)
    assert model.fused_rollout
    ref = ms.BatchedMonolithEnv(n, max_steps=20, seed=21, info_level="none", track_stats=False)
    obs, _ = ref.reset()
    mask = ref.action_masks()
    for r in range(3):
        adv, ret = model.collect_rollout()
        torch.cuda.synchronize()
        assert model.fused_rollout and env.step_variant == "hot_fused"
        b = model.buf
        for t in range(T):
            assert torch.equal(b["obs"][t], obs) and torch.equal(b["mask"][t], mask), (r, t)
            assert bool(mask.gather(1, b["act"][t][:, None]).all())
            obs, rew, term, _, _ = ref.step(b["act"][t])
            mask = ref.action_masks()
            assert torch.equal(b["rew"][t], rew) and torch.equal(b["done"][t], term), (r, t)
        assert torch.isfinite(adv).all() and torch.isfinite(ret).all()
    assert torch.equal(env.state, ref.state)
    # a configuration the fused kernel refuses falls back to policy_act + step without losing a step
    env2 = ms.BatchedMonolithEnv(256, max_steps=20, seed=2, info_level="full")
    m2 = MaskablePPO(env2, n_steps=4, seed=3, graph_rollout=graph)
    m2.collect_rollout(); m2.collect_rollout()
    assert not m2.fused_rollout and m2.num_timesteps == 2 * 4 * 256


@pytest.mark.parametrize("n", [128 * 7, 128 * 3 + 50])
def test_standalone_policy_kernel_equals_the_fused_policy_half(n):
    """msort_rollout_policy runs the very code of the fused kernel's policy half on tiles loaded from HBM: fed the observation /
    mask the fused step wrote, it returns bit-identical actions, log-probs and values; ranges compose; unaligned views work."""
    import torch
    (fa, _), pol, flat, packed = _pair(n, seed=13)
    pf = fa.rollout_pack(flat)
    fa.reset()
    a, lp, v = fa.rollout_policy(pf, seed=4, t=0)
    with torch.no_grad():
        ref = torch.log_softmax(pol.masked_logits(fa.obs, fa.mask), dim=-1)
        assert torch.allclose(lp, ref.gather(1, a[:, None]).squeeze(1), atol=5e-3)
        assert torch.allclose(v, pol.vf(fa.obs).squeeze(1), atol=5e-3, rtol=5e-3)
    assert bool(fa.mask.gather(1, a[:, None]).all())
    nxt = (torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"))
    for t in range(6):
        fa.rollout_step(a, pf, 4, t + 1, nxt)
        sa, slp, sv = fa.rollout_policy(pf, seed=4, t=t + 1)
        assert torch.equal(sa, nxt[0]) and torch.equal(slp, nxt[1]) and torch.equal(sv, nxt[2]), t
        a = nxt[0].clone()
    # two ranges == the whole batch
    out = (torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"))
    fa.rollout_policy(pf, seed=4, t=6, out=out, env_range=(0, 256)); fa.rollout_policy(pf, seed=4, t=6, out=out, env_range=(256, n))
    assert torch.equal(out[0], sa) and torch.equal(out[1], slp)
    # tensors that are not 16-byte aligned take the plain-load path
    ob = torch.empty(n * 29 + 1, device="cuda")[1:].view(n, 29); ob.copy_(fa.obs)
    mk = torch.empty(n * 22 + 1, dtype=torch.bool, device="cuda")[1:].view(n, 22); mk.copy_(fa.mask)
    ua, ulp, uv = fa.rollout_policy(pf, seed=4, t=6, obs=ob, mask=mk)
    assert torch.equal(ua, sa) and torch.equal(ulp, slp) and torch.equal(uv, sv)


def test_ppo_split_rollout_equals_fused_rollout():
    """fused_rollout="split" (msort_rollout_policy + msort_step per env-step, same arithmetic) fills the buffers exactly as the
    one-kernel rollout does."""
    import torch
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200.ppo import MaskablePPO
    runs = []
    for mode in (True, "split"):
        env = ms.BatchedMonolithEnv(128 * 10, max_steps=15, seed=6, info_level="none", track_stats=False)
        model = MaskablePPO(env, n_steps=10, seed=2, fused_rollout=mode)
        outs = [model.collect_rollout() for _ in range(3)]
        torch.cuda.synchronize()
        assert model.fused_rollout == (mode is True) and model.split_rollout == (mode == "split")
        runs.append((model.buf, outs, env.state.clone()))
    (b1, o1, s1), (b2, o2, s2) = runs
    for k in b1:
        assert torch.equal(b1[k], b2[k]), k
    assert torch.equal(s1, s2) and all(torch.equal(x[0], y[0]) for x, y in zip(o1, o2))


def test_fused_step_accumulates_the_same_episode_statistics():
    """The fused kernel keeps the step kernel's statistics path (per-warp REDUX sums -> one atomic per CTA and slot): after the
    same 60 steps the 16-slot stats vector, episode returns and lengths equal those of the plain kernel."""
    import torch
    n = 128 * 9 + 5
    (fa, pl), pol, flat, packed = _pair(n, seed=21, max_steps=10)
    pf = fa.rollout_pack(flat)
    fa.reset(); pl.reset()
    a, _, _ = pl.rollout_policy(pf, seed=2, t=0)
    nxt = (torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"))
    for t in range(60):
        fa.rollout_step(a, pf, 2, t + 1, nxt); pl.step(a)
        a = nxt[0].clone()
    torch.cuda.synchronize()
    assert fa.stats is not None and torch.allclose(fa.stats, pl.stats, rtol=1e-12, atol=1e-9)    # (float64 atomics: order-dependent last bits)
    assert float(fa.stats[0]) == n * 6 and float(fa.stats[3]) == n * 60          # episodes (10 steps each), env-steps
    for k in ("episode_return", "episode_length", "terminal_observation"):
        if k in fa.info_buffers and fa.info_buffers[k] is not None:
            assert torch.equal(fa.info_buffers[k], pl.info_buffers[k]), k
