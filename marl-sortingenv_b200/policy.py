"""Weights of Env_2's embedded sorting policy (13 -> 32 -> 32 -> 2, tanh).

ref: `sort_agent.predict(sort_obs, deterministic=True)` (env_2_press.py:106-109); the agent is
an SB3 `MlpPolicy` with `net_arch=dict(pi=[32,32], vf=[32,32])` (training.py:115), i.e.
Flatten -> mlp_extractor.policy_net (Linear,Tanh,Linear,Tanh) -> action_net -> argmax.
The kernel wants one flat fp32 vector [W1(32x13) b1(32) W2(32x32) b2(32) W3(2x32) b3(2)].
"""
from __future__ import annotations

import numpy as np
import torch

from ._abi import POLICY_WEIGHTS

SB3_KEYS = ["mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias",
            "mlp_extractor.policy_net.2.weight", "mlp_extractor.policy_net.2.bias",
            "action_net.weight", "action_net.bias"]
SHAPES = [(32, 13), (32,), (32, 32), (32,), (2, 32), (2,)]


def flatten_sort_policy(agent) -> torch.Tensor:
    """Accepts: flat array/tensor of 1570 floats; the path of a model archive (`.zip` with `policy.pth`); a state_dict with SB3's key names; an object
    with `.policy.state_dict()` (an SB3 PPO model) or `.state_dict()` (an SB3 policy); this package's MaskablePPO / MaskableActorCritic
    trained on BatchedSortingEnv (its policy tower)."""
    if isinstance(agent, str):                                      # a model archive (SB3's `model.save()` layout / MaskablePPO.save)
        return load_sb3_zip(agent if agent.endswith(".zip") else agent + ".zip")
    if isinstance(agent, (np.ndarray, list, tuple)):
        agent = torch.as_tensor(np.asarray(agent, dtype=np.float32))
    if isinstance(agent, torch.Tensor):
        w = agent.detach().to(torch.float32).reshape(-1).cpu()
        if w.numel() != POLICY_WEIGHTS:
            raise ValueError(f"sort policy needs {POLICY_WEIGHTS} weights, got {w.numel()}")
        return w
    sd = agent
    pol = getattr(agent, "policy", agent)
    if hasattr(pol, "pi") and hasattr(pol, "vf") and not isinstance(agent, dict):   # this package's MaskablePPO / MaskableActorCritic
        from .ppo import sort_policy_weights
        return flatten_sort_policy(sort_policy_weights(pol))
    if not isinstance(sd, dict):
        if hasattr(agent, "policy") and hasattr(agent.policy, "state_dict"):
            sd = agent.policy.state_dict()
        elif hasattr(agent, "state_dict"):
            sd = agent.state_dict()
        else:
            raise TypeError("unsupported sort agent: pass 1570 weights, a state_dict or an SB3 model")
    parts = []
    for key, shape in zip(SB3_KEYS, SHAPES):
        t = torch.as_tensor(sd[key]).detach().to(torch.float32).cpu()
        if tuple(t.shape) != shape:
            raise ValueError(f"{key}: expected shape {shape}, got {tuple(t.shape)}")
        parts.append(t.reshape(-1))
    return torch.cat(parts)


def load_sb3_zip(path: str) -> torch.Tensor:
    """Weights of the sorting policy from an SB3 model archive written by `model.save()`
    (ref: training.py:271-287 saves `./models/{prefix}_{timesteps}.zip`).  The archive holds
    `policy.pth`, a torch state-dict with the keys listed in SB3_KEYS."""
    import io
    import zipfile
    with zipfile.ZipFile(path) as z:
        with z.open("policy.pth") as f:
            sd = torch.load(io.BytesIO(f.read()), map_location="cpu", weights_only=True)
    return flatten_sort_policy(sd)


def sb3_style_init(seed: int = 0, action_gain: float = 0.01) -> torch.Tensor:
    """Random-init weights the way SB3 initialises an MlpPolicy: orthogonal with gains
    (sqrt 2, sqrt 2, action_gain), zero biases."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for shape, gain in (((32, 13), 2 ** 0.5), ((32, 32), 2 ** 0.5), ((2, 32), action_gain)):
        w = torch.empty(shape)
        torch.nn.init.orthogonal_(w, gain=gain, generator=g)
        out += [w.reshape(-1), torch.zeros(shape[0])]
    return torch.cat(out)
