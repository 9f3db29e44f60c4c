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


# ---------------------------------------------------------------------------------------------------------------------
# the hand-written update kernels (csrc/msort_ppo.cu, C ABI msort_ppo_*) against PyTorch fp32 autograd / torch.optim.Adam
def _ppo_case(D, A, rows, seed):
    import ctypes as C
    import torch
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200 import _abi
    from marl_sortingenv_b200.ppo import MaskableActorCritic, flatten_parameters
    lib = _abi.load_library()
    g = torch.Generator(device="cuda").manual_seed(seed)
    torch.manual_seed(seed)
    pol = MaskableActorCritic(D, A).cuda()
    with torch.no_grad():                                   # away from the tiny-gain init: every term of the gradient matters
        for p in pol.parameters():
            p.add_(0.3 * torch.randn(p.shape, device="cuda", generator=g))
    flat = flatten_parameters(pol)
    obs = torch.rand((rows, D), device="cuda", generator=g) * 2 - 0.5
    mask = torch.rand((rows, A), device="cuda", generator=g) < 0.6
    mask[:, 0] |= ~mask.any(dim=1)
    act = torch.multinomial(mask.float(), 1, generator=g).squeeze(1)
    old_logp = torch.log_softmax(pol.masked_logits(obs, mask), -1).gather(1, act[:, None]).squeeze(1).detach()
    old_logp = old_logp + 0.3 * torch.randn(rows, device="cuda", generator=g)      # ratios on both sides of the clip range
    adv = torch.randn(rows, device="cuda", generator=g) * 2 + 0.3
    ret = torch.randn(rows, device="cuda", generator=g)
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())                   # noqa: E731
    batch = _abi.MsortPpoBatch(C.sizeof(_abi.MsortPpoBatch), D, A, 0, rows, p(obs), p(mask), p(act), p(old_logp), p(adv), p(ret))
    return dict(lib=lib, pol=pol, flat=flat, obs=obs, mask=mask, act=act, old_logp=old_logp, adv=adv, ret=ret, batch=batch, p=p,
                keep=(obs, mask, act, old_logp, adv, ret), ms=ms)


def _torch_loss(pol, c, idx, clip=0.2, vf_coef=0.5, ent_coef=0.05, normalize=True):
    import torch
    logp, ent, v = pol.evaluate(c["obs"][idx], c["mask"][idx], c["act"][idx])
    a = c["adv"][idx]
    if normalize:
        a = (a - a.mean()) / (a.std() + 1e-8)
    ratio = (logp - c["old_logp"][idx]).exp()
    pg = -torch.min(a * ratio, a * ratio.clamp(1 - clip, 1 + clip)).mean()
    vl = torch.nn.functional.mse_loss(v, c["ret"][idx])
    return pg + vf_coef * vl - ent_coef * ent.mean(), pg, vl, ent.mean()


@pytest.mark.parametrize("D,A,rows", [(29, 22, 5000), (16, 11, 1283), (13, 2, 640), (29, 22, 77)])
def test_native_forward_and_gradient_match_torch_autograd(D, A, rows):
    import ctypes as C
    import torch
    from marl_sortingenv_b200 import _abi
    c = _ppo_case(D, A, rows, seed=D * 100 + A)
    lib, pol, flat, p = c["lib"], c["pol"], c["flat"], c["p"]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    # forward: log-prob of the taken action and the value
    logp, val = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    _abi.check(lib, lib.msort_ppo_forward(C.byref(c["batch"]), p(flat), p(logp), p(val), st), "forward")
    with torch.no_grad():
        rl, _, rv = pol.evaluate(c["obs"], c["mask"], c["act"])
    assert torch.allclose(logp, rl, rtol=1e-4, atol=2e-5) and torch.allclose(val, rv, rtol=1e-4, atol=2e-5)
    # gradient of a shuffled minibatch (idx) of ~2/3 of the rows
    perm = torch.randperm(rows, device="cuda")
    first, count = rows // 7, (2 * rows) // 3
    hp = _abi.MsortPpoHparams(C.sizeof(_abi.MsortPpoHparams), 1, 0.2, 0.5, 0.05, 3e-4, 0.9, 0.999, 1e-5, 0.5)
    P = lib.msort_ppo_param_count(D, A)
    assert P == flat.numel()
    grads, scratch, stats = torch.zeros(P, device="cuda"), torch.zeros(lib.msort_ppo_scratch_floats(D, A), device="cuda"), torch.zeros(5, device="cuda")
    _abi.check(lib, lib.msort_ppo_gradient(C.byref(c["batch"]), C.byref(hp), p(flat), p(grads), p(perm), first, count, p(scratch),
                                           p(stats), st), "gradient")
    idx = perm[first:first + count]
    loss, pg, vl, ent = _torch_loss(pol, c, idx)
    pol.zero_grad()
    loss.backward()
    ps = [pol.pi[0].weight, pol.pi[0].bias, pol.pi[2].weight, pol.pi[2].bias, pol.pi[4].weight, pol.pi[4].bias,
          pol.vf[0].weight, pol.vf[0].bias, pol.vf[2].weight, pol.vf[2].bias, pol.vf[4].weight, pol.vf[4].bias]
    ref = torch.cat([q.grad.reshape(-1) for q in ps])
    err = (grads - ref).abs().max().item()
    assert err <= 1e-3 * ref.abs().max().item() + 1e-7, (err, ref.abs().max().item())       # the bar VERDICT names: rel 1e-3
    assert torch.allclose(grads, ref, rtol=2e-3, atol=2e-4 * ref.abs().max().item())
    s = stats.tolist()
    assert s[4] == count
    assert abs(s[0] / count - pg.item()) < 1e-4 + 1e-4 * abs(pg.item())
    assert abs(s[1] / count - vl.item()) < 1e-4 * max(1.0, vl.item())
    assert abs(s[2] / count - ent.item()) < 1e-4


def test_native_update_matches_torch_adam_steps():
    """msort_ppo_update (gradient -> global-norm clip -> Adam per minibatch) against the same minibatches through autograd,
    clip_grad_norm_ and torch.optim.Adam(eps=1e-5): parameters after 6 optimizer steps."""
    import ctypes as C
    import torch
    from marl_sortingenv_b200 import _abi
    from marl_sortingenv_b200.ppo import MaskableActorCritic, flatten_parameters
    D, A, rows, bs = 29, 22, 3000, 1024
    c = _ppo_case(D, A, rows, seed=5)
    lib, pol, flat, p = c["lib"], c["pol"], c["flat"], c["p"]
    ref = MaskableActorCritic(D, A).cuda()
    ref.load_state_dict(pol.state_dict())
    opt = torch.optim.Adam(ref.parameters(), lr=3e-3, eps=1e-5)
    perms = torch.stack([torch.randperm(rows, device="cuda") for _ in range(2)])
    hp = _abi.MsortPpoHparams(C.sizeof(_abi.MsortPpoHparams), 1, 0.2, 0.5, 0.05, 3e-3, 0.9, 0.999, 1e-5, 0.5)
    P = flat.numel()
    z = lambda k, dt=torch.float32: torch.zeros(k, device="cuda", dtype=dt)      # noqa: E731
    grads, m, v, step, scratch, stats = z(P), z(P), z(P), z(1, torch.int32), z(lib.msort_ppo_scratch_floats(D, A)), z(5)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _abi.check(lib, lib.msort_ppo_update(C.byref(c["batch"]), C.byref(hp), p(flat), p(grads), p(m), p(v), p(step), p(perms), 2, bs,
                                         p(scratch), p(stats), st), "update")
    for e in range(2):
        for s in range(0, rows, bs):
            loss, *_ = _torch_loss(ref, c, perms[e, s:s + bs])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
            opt.step()
    assert int(step.item()) == 6 and float(grads.abs().max()) == 0.0
    for (k, a), b in zip(pol.state_dict().items(), ref.state_dict().values()):
        assert torch.allclose(a, b, rtol=1e-3, atol=3e-5), (k, (a - b).abs().max().item())


def test_native_gae_matches_the_torch_scan():
    import ctypes as C
    import torch
    from marl_sortingenv_b200 import _abi
    lib = _abi.load_library()
    T, n = 37, 1000
    g = torch.Generator(device="cuda").manual_seed(1)
    rew, val = torch.randn((T, n), device="cuda", generator=g), torch.randn((T, n), device="cuda", generator=g)
    done = torch.rand((T, n), device="cuda", generator=g) < 0.1
    last_v = torch.randn(n, device="cuda", generator=g)
    adv, ret = torch.empty_like(rew), torch.empty_like(rew)
    p = lambda t: C.c_void_p(t.data_ptr())      # noqa: E731
    _abi.check(lib, lib.msort_ppo_gae(T, n, p(rew), p(val), p(done), p(last_v), 0.99, 0.95, p(adv), p(ret),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "gae")
    ref, gae = torch.zeros_like(rew), torch.zeros(n, device="cuda")
    for t in reversed(range(T)):
        nonterm = (~done[t]).float()
        nv = last_v if t == T - 1 else val[t + 1]
        gae = rew[t] + 0.99 * nv * nonterm - val[t] + 0.99 * 0.95 * nonterm * gae
        ref[t] = gae
    assert torch.allclose(adv, ref, rtol=1e-5, atol=1e-5) and torch.allclose(ret, ref + val, rtol=1e-5, atol=1e-5)


def test_graph_rollout_equals_the_eager_rollout():
    """The CUDA-graph replay of the n_steps loop (draw index = t + the device-side counter) fills the buffers exactly as the
    eager loop does, rollout after rollout; and the importance ratio of the first epoch starts at exactly 1."""
    import torch
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200.ppo import MaskablePPO
    runs = []
    for graph in (False, True):
        env = ms.BatchedMonolithEnv(4096, max_steps=30, seed=11, info_level="none", track_stats=False)
        model = MaskablePPO(env, n_steps=16, seed=3, graph_rollout=graph)
        outs = []
        for _ in range(3):                       # rollout 1 is the eager warm-up, 2 captures + replays, 3 replays
            adv, ret = model.collect_rollout()
            torch.cuda.synchronize()
            outs.append((adv.clone(), ret.clone(), {k: v.clone() for k, v in model.buf.items()}))
        assert (model._graph is not None) == graph
        runs.append((outs, env.state.clone()))
    for (a1, r1, b1), (a2, r2, b2) in zip(runs[0][0], runs[1][0]):
        assert torch.equal(a1, a2) and torch.equal(r1, r2)
        for k in b1:
            assert torch.equal(b1[k], b2[k]), k
    assert torch.equal(runs[0][1], runs[1][1])
    # old log-probs come from the update's own fp32 forward: ratio == 1 before the first optimizer step
    b = model.buf
    with torch.no_grad():
        lp, _, _ = model.policy.evaluate(b["obs"].reshape(-1, model.D), b["mask"].reshape(-1, model.A), b["act"].reshape(-1))
    assert torch.allclose(lp, b["logp"].reshape(-1), rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("kind", ["sort", "press"])
def test_ppo_iterations_on_the_other_env_kinds(kind):
    """The loop is not Env_3-only: Env_1 (13 obs, 2 actions) and Env_2 (16, 11) run the r01 policy kernel in the graph rollout and
    the native update kernels' other two instantiations; three iterations change the weights and keep everything finite."""
    import torch
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200.ppo import MaskablePPO
    cls = {"sort": ms.BatchedSortingEnv, "press": ms.BatchedPressingEnv}[kind]
    env = cls(640, max_steps=30, seed=3, info_level="none", track_stats=False)
    model = MaskablePPO(env, n_steps=16, batch_size=2048, n_epochs=3, seed=1)
    assert not model.fused_rollout and not model.split_rollout
    before = model.flat.clone()
    model.learn(3 * 16 * 640, log_every=1)
    torch.cuda.synchronize()
    assert model._graph is not None and model.num_timesteps == 3 * 16 * 640
    assert torch.isfinite(model.flat).all() and not torch.equal(before, model.flat)
    assert bool(model.buf["mask"].gather(2, model.buf["act"][..., None]).all())
    st = model.log[-1]
    assert all(k in st for k in ("pg", "vf", "ent", "clip_fraction")) and st["ent"] > 0


def test_sort_agent_archive_round_trip_into_env2(tmp_path):
    """The reference's pipeline (training.py:271-287 `model.save()` -> main.py loads the sort agent -> Env_2_Pressing.set_agents):
    a sort agent trained here on Env_1 is saved in SB3's archive layout (policy.pth with SB3's key names), read back by
    `load_sb3_zip` / `set_agents(sort_agent=<path>)`, and Env_2 then walks exactly the trajectory it walks with the live model."""
    import torch
    import marl_sortingenv_b200 as ms
    from marl_sortingenv_b200.policy import SB3_KEYS, flatten_sort_policy, load_sb3_zip
    from marl_sortingenv_b200.ppo import MaskablePPO
    env = ms.BatchedSortingEnv(512, max_steps=50, seed=1, info_level="none", track_stats=False)
    model = MaskablePPO(env, n_steps=16, batch_size=4096, n_epochs=2, seed=5)
    model.learn(4 * 16 * 512)
    path = model.save(str(tmp_path / "sort_32768"))
    assert path.endswith(".zip")
    import io, zipfile
    with zipfile.ZipFile(path) as z:
        assert {"policy.pth", "data", "_stable_baselines3_version"} <= set(z.namelist())
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
    assert set(SB3_KEYS) <= set(sd) and "value_net.weight" in sd and "mlp_extractor.value_net.2.bias" in sd
    assert torch.equal(load_sb3_zip(path), flatten_sort_policy(model))
    # a second model restores the towers from the archive (parameters are views of its flat kernel buffer)
    m2 = MaskablePPO(ms.BatchedSortingEnv(512, max_steps=50, seed=1, info_level="none", track_stats=False), n_steps=16, seed=99).load(path)
    assert torch.equal(m2.flat, model.flat)
    a, b = (ms.BatchedPressingEnv(128 * 9 + 3, max_steps=30, seed=8, info_level="episode") for _ in range(2))
    a.set_agents(sort_agent=path); b.set_agents(sort_agent=model)
    a.reset(); b.reset()
    for t in range(40):
        act = a.sample_actions(4, t)
        a.step(act); b.step(act)
    assert torch.equal(a.state, b.state) and torch.equal(a.obs, b.obs)
