"""msort_policy_act (tcgen05 actor-critic inference + masked categorical draw) against a plain PyTorch
fp32 reference of the same op: masked log-softmax of the policy tower, value tower.  The kernel feeds
fp16 operands to the tensor cores (fp32 accumulation) and uses the hardware tanh, so log-probs and
values agree to ~1e-3 absolute (measured 1.1e-3 / 7e-4), not bitwise; the tolerance below is 5e-3."""
import pytest

pytestmark = pytest.mark.gpu


def _setup(kind, n, seed=0, gain=1.0):
    import torch
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200.batched import ENV_CLASSES
    from marl_sortingenv_b200.ppo import MaskableActorCritic, pack_actor_critic
    env = ENV_CLASSES[kind](n, max_steps=50, seed=3, info_level="none", track_stats=False)
    env.reset()
    for t in range(37):                                  # a state with varied levels / masks
        env.step(env.sample_actions(5, t))
    torch.manual_seed(seed)
    pol = MaskableActorCritic(env.D, env.A).cuda()
    with torch.no_grad():                                # non-trivial logits and biases
        pol.pi[4].weight.mul_(gain / 0.01)
        for net in (pol.pi, pol.vf):
            for i in (0, 2, 4):
                net[i].bias.uniform_(-0.3, 0.3)
    return env, pol, pack_actor_critic(pol)


@pytest.mark.parametrize("kind,n", [("mono", 128 * 37 + 5), ("press", 1000), ("sort", 300), ("mono", 77)])
def test_policy_act_matches_torch_reference(kind, n):
    import torch
    env, pol, packed = _setup(kind, n)
    obs, mask = env.obs.clone(), env.mask.clone()
    a, lp, v = env.policy_act(packed, seed=11, t=4)
    torch.cuda.synchronize()
    with torch.no_grad():
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        logp_all = torch.log_softmax(pol.masked_logits(obs, mask), dim=-1)
        v_ref = pol.vf(obs).squeeze(1)
        torch.backends.cuda.matmul.allow_tf32 = prev
    assert a.dtype == torch.int64 and int(a.min()) >= 0 and int(a.max()) < env.A
    assert bool(mask.gather(1, a[:, None]).all()), "sampled an invalid action"
    lp_ref = logp_all.gather(1, a[:, None]).squeeze(1)
    print("max |dlogp|", float((lp - lp_ref).abs().max()), "max |dv|", float((v - v_ref).abs().max()))
    assert torch.allclose(lp, lp_ref, atol=5e-3, rtol=5e-3), float((lp - lp_ref).abs().max())
    assert torch.allclose(v, v_ref, atol=5e-3, rtol=5e-3), float((v - v_ref).abs().max())
    # unaligned views of a larger buffer (rows of a [T, n, D] rollout buffer with odd n): plain-load path
    big_o = torch.zeros(n * env.D + 1, device="cuda"); big_m = torch.zeros(n * env.A + 1, dtype=torch.bool, device="cuda")
    o2 = big_o[1:].view(n, env.D); o2.copy_(obs); m2 = big_m[1:].view(n, env.A); m2.copy_(mask)
    a_u, lp_u, v_u = env.policy_act(packed, seed=11, t=4, obs=o2, mask=m2)
    assert torch.equal(a_u, a) and torch.equal(lp_u, lp) and torch.equal(v_u, v)
    # same key -> same draw; another step index -> (mostly) different draws
    a2, _, _ = env.policy_act(packed, seed=11, t=4)
    assert torch.equal(a, a2)
    if n >= 1000 and kind != "sort":
        a3, _, _ = env.policy_act(packed, seed=11, t=5)
        assert float((a3 != a).float().mean()) > 0.05


def test_policy_act_deterministic_is_argmax_and_draws_follow_the_distribution():
    import torch
    env, pol, packed = _setup("mono", 4096, seed=2)
    obs, mask = env.obs.clone(), env.mask.clone()
    with torch.no_grad():
        logits = pol.masked_logits(obs, mask)
        p = torch.softmax(logits, dim=-1)
    a, _, _ = env.policy_act(packed, deterministic=True)
    top2 = logits.topk(2, dim=-1).values
    clear = (top2[:, 0] - top2[:, 1]) > 2e-2               # away from numerical ties
    assert torch.equal(a[clear], logits.argmax(-1)[clear])
    # empirical action frequencies over many keys vs the mean policy distribution
    counts = torch.zeros(env.A, device="cuda")
    T = 64
    for t in range(T):
        at, _, _ = env.policy_act(packed, seed=1, t=t)
        counts += torch.bincount(at, minlength=env.A).float()
    freq = counts / counts.sum()
    want = p.mean(0)
    assert float((freq - want).abs().max()) < 4e-3, (freq, want)


def test_fused_rollout_matches_torch_rollout_statistics():
    """collect_rollout with the fused kernel vs the torch path: same policy, same env seed -> the mean
    step reward of a rollout agrees statistically, and log-probs stored by the kernel agree with
    evaluate() (what PPO's ratio compares them to)."""
    import torch
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200.ppo import MaskablePPO
    stats = {}
    for fused in (True, False):
        env = ms.BatchedMonolithEnv(2048, max_steps=200, seed=42, noise_sorting=0.0, info_level="none", track_stats=False)
        model = MaskablePPO(env, n_steps=32, seed=0, fused_act=fused)
        model.collect_rollout()
        b = model.buf
        with torch.no_grad():
            lp_new, _, v_new = model.policy.evaluate(b["obs"].reshape(-1, env.D), b["mask"].reshape(-1, env.A), b["act"].reshape(-1))
        stats[fused] = (float(b["rew"].mean()), float((lp_new - b["logp"].reshape(-1)).abs().max()),
                        float((v_new - b["val"].reshape(-1)).abs().max()))
    assert abs(stats[True][0] - stats[False][0]) < 0.02, stats
    assert stats[True][1] < 5e-3 and stats[True][2] < 5e-3, stats


@pytest.mark.parametrize("kind", ["mono", "press", "sort"])
def test_env_ranges_on_two_streams_compose_to_the_whole_batch(kind):
    """msort_step_range / msort_policy_act_range: two disjoint env ranges (ragged second one), each on its own CUDA
    stream, leave exactly what the whole-batch calls leave — state, obs, mask, reward, done, actions, log-probs,
    values, statistics.  Range starts must be multiples of 128."""
    import torch
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200.ppo import MaskableActorCritic, pack_actor_critic
    from marl_sortingenv_b200.batched import ENV_CLASSES
    from marl_sortingenv_b200.policy import sb3_style_init
    # Env_2: more tiles per range than resident CTAs, so its persistent kernel loops inside each range
    n, cut = (128 * 1500 + 51, 128 * 800) if kind == "press" else (128 * 37 + 51, 128 * 20)
    whole = ENV_CLASSES[kind](n, max_steps=20, seed=5, info_level="episode")
    parts = ENV_CLASSES[kind](n, max_steps=20, seed=5, info_level="episode")
    if kind == "press":
        whole.set_sort_policy(sb3_style_init(2, action_gain=1.0)); parts.set_sort_policy(sb3_style_init(2, action_gain=1.0))
    whole.reset(); parts.reset()
    torch.manual_seed(1)
    packed = pack_actor_critic(MaskableActorCritic(whole.D, whole.A).cuda())
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    out_p = (torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"))
    for t in range(45):
        a, lp, v = whole.policy_act(packed, seed=9, t=t)
        whole.step(a)
        cur = torch.cuda.current_stream()
        for s, r in zip(streams, ((0, cut), (cut, n))):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                parts.policy_act(packed, seed=9, t=t, out=out_p, env_range=r)
                parts.step(out_p[0], env_range=r)
        for s in streams:
            cur.wait_stream(s)
        assert torch.equal(out_p[0], a) and torch.equal(out_p[1], lp) and torch.equal(out_p[2], v), t
        assert torch.equal(parts.state, whole.state) and torch.equal(parts.obs, whole.obs), t
        assert torch.equal(parts.mask, whole.mask) and torch.equal(parts.reward, whole.reward), t
        assert torch.equal(parts.terminated, whole.terminated)
    for k in ("episode_return", "episode_length", "terminal_observation"):
        assert torch.equal(parts.info_buffers[k], whole.info_buffers[k]), k
    assert torch.allclose(parts.stats, whole.stats, rtol=1e-12, atol=0)      # atomics: order differs, sums agree
    with pytest.raises(Exception):
        parts.step(out_p[0], env_range=(64, 256))                             # not on a tile boundary


def test_policy_eval_on_column_slices_and_modular_agents():
    """`msort_policy_eval`: the actor-critic kernels on strided rows (the 13- and 16-wide parts of Env_3's 29-wide observation),
    with and without a mask — the in-batch evaluation of the two agents of Env_3.step(mode='model') (env_monolith.py:186-221).
    Deterministic actions equal the PyTorch fp32 argmax wherever the top-2 logit gap is clear; log-probs / values within 5e-3;
    `ppo.modular_actions` with two agents trained here takes this path and composes 11 * mode + press."""
    import torch
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200.ppo import MaskableActorCritic, MaskablePPO, modular_actions, pack_actor_critic
    n = 128 * 21 + 19
    env = ms.BatchedMonolithEnv(n, max_steps=40, seed=12, info_level="episode")
    env.reset()
    for t in range(25):
        env.step(env.sample_actions(3, t))
    torch.manual_seed(4)
    obs = env.observe_after_shift()
    for (D, A, lo) in ((13, 2, 0), (16, 11, 13)):
        pol = MaskableActorCritic(D, A).cuda()
        with torch.no_grad():
            for p in pol.parameters():
                p.add_(0.3 * torch.randn_like(p))
        o = obs[:, lo:lo + D]
        m = env.action_masks()[:, :11] if A == 11 else None
        assert o.stride(0) == 29 and not o.is_contiguous()
        a, lp, v = env.policy_eval(pack_actor_critic(pol), o, m, num_actions=A, deterministic=True)
        with torch.no_grad():
            mm = m if m is not None else torch.ones((n, A), dtype=torch.bool, device="cuda")
            lg = pol.masked_logits(o, mm)
            ref_lp = torch.log_softmax(lg, -1)
            top2 = lg.topk(2, dim=-1).values
            clear = (top2[:, 0] - top2[:, 1]) > 2e-2
            assert torch.equal(a[clear], lg.argmax(-1)[clear]) and int(clear.sum()) > n // 2
            assert bool(mm.gather(1, a[:, None]).all())
            assert torch.allclose(lp, ref_lp.gather(1, a[:, None]).squeeze(1), atol=5e-3)
            assert torch.allclose(v, pol.vf(o).squeeze(1), atol=5e-3, rtol=5e-3)
        # sampling mode draws valid actions too
        a2, _, _ = env.policy_eval(pack_actor_critic(pol), o, m, num_actions=A, deterministic=False, seed=5, t=1)
        assert bool(mm.gather(1, a2[:, None]).all())
    # modular_actions with agents of this module = the kernel path; with SB3-style stand-ins = predict()
    sort_env = ms.BatchedSortingEnv(256, max_steps=20, seed=1, info_level="none", track_stats=False)
    press_env = ms.BatchedPressingEnv(256, max_steps=20, seed=1, info_level="none", track_stats=False)
    sa, pa = MaskablePPO(sort_env, n_steps=4, seed=1), MaskablePPO(press_env, n_steps=4, seed=2)
    with torch.no_grad():
        for p in list(sa.policy.parameters()) + list(pa.policy.parameters()):
            p.add_(0.3 * torch.randn_like(p))
    before = env.launch_count
    act = modular_actions(env, sa, pa, use_action_masking=True)
    assert env.launch_count - before >= 3            # observe_after_shift + two policy_eval launches

    class Eager:                                     # the same towers through predict(): the fallback path
        def __init__(self, m): self.m = m
        def predict(self, o, deterministic=True, action_masks=None): return self.m.predict(o, action_masks=action_masks, deterministic=deterministic)
    ref = modular_actions(env, Eager(sa), Eager(pa), use_action_masking=True)
    agree = float((act == ref).float().mean())
    assert agree > 0.97, agree                       # equal except near-ties of the fp16-operand logits
    assert bool(((act >= 0) & (act < 22)).all()) and bool(env.action_masks().gather(1, act[:, None]).all())
