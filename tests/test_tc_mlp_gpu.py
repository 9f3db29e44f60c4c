"""Env_2's embedded sort policy on the tensor cores (step kernel instantiation HOT_TENSOR; tc_mlp_mode in
msort_kernels.cu) — ref: sort_agent.predict(sort_obs, deterministic=True), env_2_press.py:106-109, an SB3
MlpPolicy 13 -> 32 -> 32 -> 2 with tanh (training.py:115).

Parity contract: identical argmax except where |logit0 - logit1| < 1e-5.  The tests
bound the logit error of the fp16-split tcgen05 evaluation against a float64 evaluation of the same network
(and show the fp32 reference's own error beside it), check which kernel runs, and count how often the tensor
form and the per-thread fp32 FFMA2 form disagree over a large batch."""
from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-6          # |tensor-core logit - float64 logit|; the contract's tie band is 1e-5


def _net_f64(w, x):
    w = np.asarray(w, dtype=np.float64)
    W1, b1 = w[:416].reshape(32, 13), w[416:448]
    W2, b2 = w[448:1472].reshape(32, 32), w[1472:1504]
    W3, b3 = w[1504:1568].reshape(2, 32), w[1568:1570]
    h = np.tanh(x.astype(np.float64) @ W1.T + b1)
    h = np.tanh(h @ W2.T + b2)
    return h @ W3.T + b3


def _net_f32(w, x):
    import torch
    w = torch.as_tensor(np.asarray(w, dtype=np.float32))
    x = torch.as_tensor(x.astype(np.float32))
    W1, b1 = w[:416].reshape(32, 13), w[416:448]
    W2, b2 = w[448:1472].reshape(32, 32), w[1472:1504]
    W3, b3 = w[1504:1568].reshape(2, 32), w[1568:1570]
    h = torch.tanh(x @ W1.T + b1)
    h = torch.tanh(h @ W2.T + b2)
    return (h @ W3.T + b3).numpy()


def _press_env(n, weights, **kw):
    import torch
    from marl_sortingenv_b200.batched import BatchedPressingEnv
    env = BatchedPressingEnv(n, device="cuda:0", max_steps=25, seed=3, info_level="episode", **kw)
    env.set_sort_policy(torch.as_tensor(np.asarray(weights, dtype=np.float32)))
    return env


@pytest.mark.parametrize("case", ["sb3_init_gain1", "sb3_init", "wide", "tiny_last_layer"])
def test_tensor_policy_logits_match_float64(case):
    import torch
    from marl_sortingenv_b200.policy import sb3_style_init
    rng = np.random.default_rng(5)
    if case == "sb3_init_gain1":
        w = sb3_style_init(3, action_gain=1.0).numpy()
    elif case == "sb3_init":
        w = sb3_style_init(0).numpy()                       # SB3's own init: last layer gain 0.01
    elif case == "wide":
        w = rng.normal(0.0, 1.2, size=1570).astype(np.float32)   # saturating pre-activations, |weights| up to ~5
    else:
        w = sb3_style_init(7, action_gain=1.0).numpy()
        w[1504:] *= 1e-4                                    # logits ~1e-4: the host's power-of-two scale must carry them
    n = 128 * 300 + 77                                      # ragged: the diagnostics kernel pads its last tile
    x = rng.uniform(-1.0, 1.0, size=(n, 13)).astype(np.float32)
    x[:, 0] = rng.uniform(0.0, 1.0, size=n)                 # like the real observation: occupancy / proportions / accuracies in [0, 1]
    x[::7, 5:9] = 1.0                                       # boosted accuracies are exactly 1.0
    x[::11] = 0.0
    env = _press_env(256, w)
    got = env.policy_logits_tensor(torch.as_tensor(x).cuda()).cpu().numpy().astype(np.float64)
    want = _net_f64(w, x)
    scale = max(1.0, float(np.abs(want).max()))
    err = np.abs(got - want).max()
    err32 = np.abs(_net_f32(w, x).astype(np.float64) - want).max()
    # what decides the argmax is the DIFFERENCE of the two logits
    derr = np.abs((got[:, 1] - got[:, 0]) - (want[:, 1] - want[:, 0])).max()
    print(f"{case}: max |tensor - f64| = {err:.3e} (logit difference {derr:.3e}); torch fp32 vs f64 = {err32:.3e}; max |logit| = {scale:.3g}")
    assert err <= LOGIT_TOL * scale, f"{case}: tensor-core logits off by {err:g}"
    flips = (got[:, 1] > got[:, 0]) != (want[:, 1] > want[:, 0])
    margin = np.abs(want[:, 1] - want[:, 0])
    assert not flips.any() or margin[flips].max() < 1e-5 * scale, "argmax differs away from a tie"
    env.close()


def test_tensor_policy_is_what_the_hot_kernel_runs_and_can_be_switched_off():
    import torch
    from marl_sortingenv_b200 import _abi
    from marl_sortingenv_b200.policy import sb3_style_init
    w = sb3_style_init(3, action_gain=1.0).numpy()
    env = _press_env(128 * 64, w)
    env.reset(seed=3)
    a = env.sample_actions(1, 0)
    env.step(a)
    assert env.step_variant == ("hot_tensor_split" if env.get_option(_abi.OPT_TENSOR_POLICY) == 2 else "hot_tensor")
    env.set_option(_abi.OPT_TENSOR_POLICY, 0)
    env.step(env.sample_actions(1, 1))
    assert env.step_variant == "hot_persistent"
    env.set_option(_abi.OPT_TENSOR_POLICY, 1)
    env.step(env.sample_actions(1, 3))
    assert env.step_variant == "hot_tensor"                 # 1 = inside the step kernel, 2 = policy kernel + step kernel
    env.set_option(_abi.OPT_TENSOR_POLICY, 2)
    env.step(env.sample_actions(1, 4))
    assert env.step_variant == "hot_tensor_split"
    # a policy outside fp16's range cannot be split: the FFMA2 kernel keeps evaluating it
    big = w.copy(); big[5] = 1.0e5
    env.set_sort_policy(torch.as_tensor(big))
    env.step(env.sample_actions(1, 2))
    assert env.step_variant == "hot_persistent"
    torch.cuda.synchronize()
    env.close()


def test_tensor_and_ffma2_forms_walk_the_same_trajectories():
    """1 048 576 envs x 50 steps (BASELINE config 3's kernel at config 4's size), the same actions into a handle with the
    tensor-core policy and one with the fp32 FFMA2 policy: envs whose sort mode ever differs must be rare (ties of the
    two logits inside ~1e-6), every other env must end in the bit-identical state."""
    import torch
    from marl_sortingenv_b200 import _abi
    from marl_sortingenv_b200.policy import sb3_style_init
    n, T = 1 << 20, 50
    w = sb3_style_init(3, action_gain=1.0).numpy()
    a_env, b_env = _press_env(n, w), _press_env(n, w)
    b_env.set_option(_abi.OPT_TENSOR_POLICY, 0)
    a_env.reset(seed=9); b_env.reset(seed=9)
    act = torch.empty(n, dtype=torch.int64, device="cuda:0")
    for t in range(T):
        a_env.sample_actions(2, t, out=act)
        # masks agree on every env that has not diverged; a diverged env may hold an action the other side's mask
        # forbids — in masked mode the kernel then simply starts that press, which is fine for a divergence count
        a_env.step(act); b_env.step(act)
    assert a_env.step_variant in ("hot_tensor", "hot_tensor_split") and b_env.step_variant == "hot_persistent"
    sa = a_env.state.view(torch.int32).reshape(-1, a_env.state.numel() // 4 // 13)    # [13 planes, n_pad * 4 words]
    sb = b_env.state.view(torch.int32).reshape(-1, b_env.state.numel() // 4 // 13)
    diff = (sa != sb).reshape(13, -1, 4).any(dim=2).any(dim=0)[:n]
    n_div = int(diff.sum())
    print(f"tensor vs FFMA2 policy: {n_div} of {n} envs diverged within {T} steps ({n_div / (n * T):.2e} per env-step)")
    assert n_div <= n * T * 2e-6, f"{n_div} envs diverged"
    a_env.close(); b_env.close()


def test_split_form_equals_the_fused_tensor_kernel_bit_for_bit():
    """MSORT_OPT_TENSOR_POLICY 2 (press_policy_kernel writes one mode byte per env, the policy-free EXTMODE step kernel reads it)
    and 1 (policy inside the step kernel) call the same tc_mlp_mode on the same 13-wide observation: identical state, obs,
    reward, mask and statistics after every one of 60 steps, ragged batch and env ranges included."""
    import torch
    from marl_sortingenv_b200 import _abi
    from marl_sortingenv_b200.policy import sb3_style_init
    n = 128 * 37 + 55
    w = sb3_style_init(5, action_gain=1.0).numpy()
    a_env, b_env = _press_env(n, w), _press_env(n, w)
    a_env.set_option(_abi.OPT_TENSOR_POLICY, 1); b_env.set_option(_abi.OPT_TENSOR_POLICY, 2)
    a_env.reset(seed=4); b_env.reset(seed=4)
    act = torch.empty(n, dtype=torch.int64, device="cuda:0")
    for t in range(60):
        a_env.sample_actions(6, t, out=act)
        a_env.step(act)
        if t % 2:
            b_env.step(act)
        else:                                               # two env ranges on the current stream
            b_env.step(act, env_range=(0, 128 * 20)); b_env.step(act, env_range=(128 * 20, n))
        assert a_env.step_variant == "hot_tensor" and b_env.step_variant == "hot_tensor_split"
        assert torch.equal(a_env.state, b_env.state), t
        assert torch.equal(a_env.obs, b_env.obs) and torch.equal(a_env.reward, b_env.reward) and torch.equal(a_env.mask, b_env.mask), t
    if a_env.stats is not None:
        assert torch.allclose(a_env.stats, b_env.stats, rtol=1e-12, atol=1e-9)
    a_env.close(); b_env.close()
