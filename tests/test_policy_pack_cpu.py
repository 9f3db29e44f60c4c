"""Host-side check of ppo.pack_actor_critic: the packed buffer, read back in the operand order the tcgen05
kernel uses ([k/8][n][k%8] fp16 words per layer, then the fp32 biases), must reproduce the PyTorch
actor-critic (both towers) up to fp16 weight rounding."""
import numpy as np
import pytest
import torch

from marl_sortingenv_b200 import _abi
from marl_sortingenv_b200.ppo import MaskableActorCritic, pack_actor_critic


def _unpack(packed: torch.Tensor):
    w = packed.detach().cpu().contiguous()
    assert w.dtype == torch.float32 and w.numel() == _abi.POLICY_ACT_WEIGHTS
    halves = w[:4096].view(torch.float16).float().numpy()          # 8192 fp16 weights
    bias = w[4096:].numpy()
    mats, o = [], 0
    for (K, N) in ((32, 64), (64, 64), (64, 32)):
        blk = halves[o:o + K * N].reshape(K // 8, N, 8)             # [k/8][n][k%8]
        mats.append(blk.transpose(1, 0, 2).reshape(N, K)); o += K * N
    return mats, (bias[:64], bias[64:128], bias[128:160])


@pytest.mark.parametrize("D,A", [(29, 22), (16, 11), (13, 2)])
def test_packed_layout_reproduces_the_towers(D, A):
    torch.manual_seed(D)
    pol = MaskableActorCritic(D, A)
    with torch.no_grad():
        pol.pi[4].weight.mul_(50.0)
        for net in (pol.pi, pol.vf):
            for i in (0, 2, 4):
                net[i].bias.uniform_(-0.5, 0.5)
    (W1, W2, W3), (b1, b2, b3) = _unpack(pack_actor_critic(pol))
    x = torch.rand(257, D)
    xp = np.zeros((257, 32), dtype=np.float32); xp[:, :D] = x.numpy()
    h1 = np.tanh(xp @ W1.T + b1)
    h2 = np.tanh(h1 @ W2.T + b2)
    out = h2 @ W3.T + b3
    with torch.no_grad():
        logits, value = pol.pi(x).numpy(), pol.vf(x).squeeze(1).numpy()
    assert np.allclose(out[:, :A], logits, atol=3e-3, rtol=3e-3)
    assert np.allclose(out[:, A], value, atol=3e-3, rtol=3e-3)
    # the towers do not leak into each other: block structure of the packed matrices
    assert not W2[:32, 32:].any() and not W2[32:, :32].any()
    assert not W3[:A, 32:].any() and not W3[A, :32].any() and not W3[A + 1:].any()
    assert not W1[:, D:].any()


def test_flatten_parameters_follows_the_layout_of_the_update_kernels():
    """ppo.flatten_parameters: ONE flat fp32 vector in the order include/msort.h documents for msort_ppo_* / msort_rollout_pack
    (pi W1 b1 W2 b2 W3 b3 | vf W1 b1 W2 b2 W3 b3, torch Linear layout), every nn.Parameter a view of it."""
    import torch
    from marl_sortingenv_b200.ppo import MaskableActorCritic, flatten_parameters
    torch.manual_seed(0)
    for D, A in ((29, 22), (16, 11), (13, 2)):
        pol = MaskableActorCritic(D, A)
        want = torch.cat([p.detach().reshape(-1).clone() for p in
                          (pol.pi[0].weight, pol.pi[0].bias, pol.pi[2].weight, pol.pi[2].bias, pol.pi[4].weight, pol.pi[4].bias,
                           pol.vf[0].weight, pol.vf[0].bias, pol.vf[2].weight, pol.vf[2].bias, pol.vf[4].weight, pol.vf[4].bias)])
        flat = flatten_parameters(pol)
        assert flat.numel() == 2 * (32 * D + 32 + 32 * 32 + 32) + (A * 32 + A) + (32 + 1) and torch.equal(flat, want)
        flat[0] = 123.0                                       # the modules see what the kernels write
        assert float(pol.pi[0].weight[0, 0]) == 123.0
        flat[-1] = -7.0
        assert float(pol.vf[4].bias[0]) == -7.0
        x = torch.rand(5, D)
        assert torch.isfinite(pol.pi(x)).all()
