"""GPU-resident MaskablePPO-style rollout + update loop (SURVEY.md §8f item 1 — a CALLER of the hot path).

The reference trains with sb3_contrib.MaskablePPO on one CPU env (training.py:118-143:
`net_arch=dict(pi=[32,32], vf=[32,32])`, `ent_coef=0.05`, SB3 defaults n_epochs=10, clip 0.2,
gamma 0.99, lambda 0.95, lr 3e-4, vf_coef 0.5, max_grad_norm 0.5).  SB3's collector loops over
envs in Python, so it cannot drive 1e4..1e6 device envs; this module is the same algorithm with
every tensor (observations, masks, actions, advantages, minibatches) resident on the GPU:

* rollout  : per env-step the tensor-core policy kernel (`msort_policy_act`) + the fused step kernel, the whole
             n_steps loop replayed as ONE CUDA graph (`graph_rollout`);
* update   : hand-written kernels (`csrc/msort_ppo.cu`, C ABI `msort_ppo_*`): GAE scan, fused actor-critic forward +
             masked log-softmax / entropy / clipped-surrogate loss + backward per minibatch, fused global-norm clip + Adam.
             The old log-probs and values are recomputed once with the update's own fp32 forward, so the importance
             ratio starts at exactly 1 whatever precision the rollout kernel used.
`native_update=False` keeps the PyTorch autograd + torch.optim.Adam path (the reference the gradient tests compare with).
"""
from __future__ import annotations

import ctypes as C
import math
import time

import torch
import torch.nn as nn

from . import _abi


def _mlp(inp: int, out: int, gain: float) -> nn.Sequential:
    net = nn.Sequential(nn.Linear(inp, 32), nn.Tanh(), nn.Linear(32, 32), nn.Tanh(), nn.Linear(32, out))
    for i, g in ((0, math.sqrt(2)), (2, math.sqrt(2)), (4, gain)):       # SB3's orthogonal init
        nn.init.orthogonal_(net[i].weight, gain=g)
        nn.init.zeros_(net[i].bias)
    return net


class MaskableActorCritic(nn.Module):
    """Separate 32-32 tanh towers for policy and value (MaskableActorCriticPolicy, training.py:115)."""

    def __init__(self, obs_dim: int, n_actions: int):
        super().__init__()
        self.pi = _mlp(obs_dim, n_actions, 0.01)
        self.vf = _mlp(obs_dim, 1, 1.0)

    def masked_logits(self, obs, mask):
        return self.pi(obs).masked_fill(~mask, -1e8)                     # sb3_contrib masks logits with -1e8

    def act(self, obs, mask, deterministic=False):
        logits = self.masked_logits(obs, mask)
        logp_all = torch.log_softmax(logits, dim=-1)
        if deterministic:
            a = logits.argmax(dim=-1)
        else:
            a = torch.multinomial(logp_all.exp(), 1).squeeze(1)
        return a, logp_all.gather(1, a[:, None]).squeeze(1), self.vf(obs).squeeze(1)

    def evaluate(self, obs, mask, actions):
        logp_all = torch.log_softmax(self.masked_logits(obs, mask), dim=-1)
        p = logp_all.exp()
        entropy = -(p * logp_all.masked_fill(~mask, 0.0)).sum(-1)
        return logp_all.gather(1, actions[:, None]).squeeze(1), entropy, self.vf(obs).squeeze(1)


def pack_actor_critic(policy: "MaskableActorCritic", out: torch.Tensor | None = None) -> torch.Tensor:
    """The two towers as the single 32 -> 64 -> 64 -> 32 network `msort_policy_act` evaluates on the
    tensor cores (include/msort.h): layer 1 concatenates the towers' first layers, layer 2 is block-
    diagonal, layer 3 puts the A logits in rows 0..A-1 and the value in row A.  Each weight matrix
    W[n][k] is stored as fp16 in the kernel's shared-memory operand order [k/8][n][k%8] (no-swizzle
    K-major core matrices; two values per 32-bit word), followed by the three fp32 bias vectors.
    Returns MSORT_POLICY_ACT_WEIGHTS 32-bit words as a float32 tensor."""
    pi, vf = policy.pi, policy.vf
    dev = pi[0].weight.device
    D, A = pi[0].weight.shape[1], pi[4].weight.shape[0]
    assert D <= 32 and A <= 31
    W1 = torch.zeros((64, 32), device=dev); W1[:32, :D] = pi[0].weight; W1[32:, :D] = vf[0].weight
    W2 = torch.zeros((64, 64), device=dev); W2[:32, :32] = pi[2].weight; W2[32:, 32:] = vf[2].weight
    W3 = torch.zeros((32, 64), device=dev); W3[:A, :32] = pi[4].weight; W3[A, 32:] = vf[4].weight[0]
    b3 = torch.zeros(32, device=dev); b3[:A] = pi[4].bias; b3[A] = vf[4].bias[0]

    def canon(W):                                   # [N, K] fp32 -> [K/8][N][8] fp16, viewed as 32-bit words
        N, K = W.shape
        return W.reshape(N, K // 8, 8).permute(1, 0, 2).reshape(-1).half().view(torch.float32)
    packed = torch.cat([canon(W1), canon(W2), canon(W3),
                        torch.cat([pi[0].bias, vf[0].bias, pi[2].bias, vf[2].bias, b3]).float()])
    if out is not None:
        out.copy_(packed)
        return out
    return packed.contiguous()


def flatten_parameters(policy: "MaskableActorCritic") -> torch.Tensor:
    """Move the two towers' parameters into ONE flat fp32 buffer in the order `msort_ppo_*` expects (include/msort.h: pi W1 b1
    W2 b2 W3 b3 | vf W1 b1 W2 b2 W3 b3, torch Linear layout) and make every nn.Parameter a view of it, so PyTorch code and
    the native kernels see the same weights."""
    ps = [policy.pi[0].weight, policy.pi[0].bias, policy.pi[2].weight, policy.pi[2].bias, policy.pi[4].weight, policy.pi[4].bias,
          policy.vf[0].weight, policy.vf[0].bias, policy.vf[2].weight, policy.vf[2].bias, policy.vf[4].weight, policy.vf[4].bias]
    flat = torch.cat([p.detach().reshape(-1) for p in ps]).contiguous()
    o = 0
    for p in ps:
        p.data = flat[o:o + p.numel()].view_as(p)
        o += p.numel()
    return flat


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class MaskablePPO:
    MIN_ENVS_PER_STREAM = 65536      # below this a second stream only adds launches to the rollout loop

    def __init__(self, env, n_steps: int = 64, batch_size: int = 8192, n_epochs: int = 10, gamma: float = 0.99,
                 gae_lambda: float = 0.95, clip_range: float = 0.2, ent_coef: float = 0.05, vf_coef: float = 0.5,
                 learning_rate: float = 3e-4, max_grad_norm: float = 0.5, seed: int = 42, fused_act: bool = True,
                 rollout_streams: int = 2, native_update: bool = True, graph_rollout: bool = True, fused_rollout: bool = True):
        self.env = env
        # rollout inference: one fused tensor-core kernel (msort_policy_act) instead of ~25 torch kernels/step
        self.fused_act = fused_act and hasattr(env, "policy_act")
        self.seed = seed
        self.n, self.D, self.A = env.num_envs, env.D, env.A
        self.dev = env.device
        torch.manual_seed(seed)
        self.policy = MaskableActorCritic(self.D, self.A).to(self.dev)
        self.lib = env.lib if hasattr(env, "lib") else None
        self.native_update = native_update and self.lib is not None
        self.graph_rollout = graph_rollout and fused_act and hasattr(env, "policy_act")
        # Env_3 in its training configuration: step + the next step's policy as one kernel per env-step (msort_rollout_step)
        # (the library refuses configurations outside its HOT instantiation: the first rollout then falls back, see _rollout_body)
        self.fused_rollout = fused_rollout and self.fused_act and getattr(env, "kind", "") == "mono" and hasattr(env, "rollout_step")
        # ... or as two kernels per env-step with the same arithmetic (msort_rollout_policy + msort_step): `fused_rollout="split"`
        self.split_rollout = fused_rollout == "split" and self.fused_rollout
        self._rollout_auto = fused_rollout is True
        self.flat = flatten_parameters(self.policy)
        self.opt = torch.optim.Adam(self.policy.parameters(), lr=learning_rate, eps=1e-5)
        self.lr = learning_rate
        if self.native_update:
            P = int(self.lib.msort_ppo_param_count(self.D, self.A))
            assert P == self.flat.numel(), (P, self.flat.numel())
            z = lambda *sh, dt=torch.float32: torch.zeros(*sh, dtype=dt, device=self.dev)   # noqa: E731
            self._grads, self._m, self._v, self._step = z(P), z(P), z(P), z(1, dt=torch.int32)
            self._scratch, self._stats = z(int(self.lib.msort_ppo_scratch_floats(self.D, self.A))), z(5)
            self._hp = _abi.MsortPpoHparams(C.sizeof(_abi.MsortPpoHparams), 1, clip_range, vf_coef, ent_coef, learning_rate,
                                            0.9, 0.999, 1e-5, max_grad_norm)
        self._graph = None
        self._packed = None
        self._tail = None
        self._warm = False
        self._draw_counter = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.n_steps, self.batch_size, self.n_epochs = n_steps, batch_size, n_epochs
        self.gamma, self.lam, self.clip = gamma, gae_lambda, clip_range
        self.ent_coef, self.vf_coef, self.max_grad_norm = ent_coef, vf_coef, max_grad_norm
        T, n, dev = n_steps, self.n, self.dev
        self.buf = dict(obs=torch.zeros((T, n, self.D), device=dev), mask=torch.zeros((T, n, self.A), dtype=torch.bool, device=dev),
                        act=torch.zeros((T, n), dtype=torch.int64, device=dev), logp=torch.zeros((T, n), device=dev),
                        val=torch.zeros((T, n), device=dev), rew=torch.zeros((T, n), device=dev),
                        done=torch.zeros((T, n), dtype=torch.bool, device=dev))
        self._adv, self._ret = torch.zeros((T, n), device=dev), torch.zeros((T, n), device=dev)
        self._last_v = torch.zeros(n, device=dev)
        self.num_timesteps = 0
        self._obs = None
        self.log = []
        # The rollout runs on `rollout_streams` CUDA streams over disjoint env ranges (msort_*_range): while the
        # tensor-core policy kernel of one range waits on its MMA round trips, the step kernel of another range uses
        # the SMs (and the tail of every kernel overlaps the head of the next): 134 -> 124 us per env-step at 1 M envs.
        # (only when a range is big enough for its kernels to outlast their launch overhead: the loop is eager)
        if self.n < rollout_streams * self.MIN_ENVS_PER_STREAM:
            rollout_streams = 1
        per = -(-self.n // max(1, rollout_streams)) if rollout_streams > 1 else self.n
        per = -(-per // 128) * 128                                       # ranges start on whole 128-env tiles
        self._ranges = [(lo, min(self.n, lo + per)) for lo in range(0, self.n, per)] if self.fused_act else [(0, self.n)]
        self._streams = [torch.cuda.Stream(device=self.dev) for _ in self._ranges] if len(self._ranges) > 1 else []
        # measured on one B200 (bench.py `rollout`, 1 048 576 envs): split on two streams 104 us per env-step, one fused kernel
        # 117 us (its 2 100 straight-line instructions per warp stall on instruction fetch), r01's two kernels 127 us; small
        # batches are launch-bound and take the one-kernel form
        if self.fused_rollout and self._rollout_auto and len(self._ranges) > 1:
            self.split_rollout = True
        if self.split_rollout:
            self.fused_rollout = False

    # ------------------------------------------------------------------ rollout
    def _batch(self, rows, obs, mask, act, logp=None, adv=None, ret=None):
        return _abi.MsortPpoBatch(C.sizeof(_abi.MsortPpoBatch), self.D, self.A, 0, int(rows), _p(obs), _p(mask), _p(act),
                                  _p(logp), _p(adv), _p(ret))

    def _cstream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def _rollout_body(self, packed, t0: int):
        """The n_steps loop on the rollout buffers (fused_act): slot 0 <- the last observation / mask, then per step
        policy_act -> step, the step writing its observation / mask straight into slot t+1 (the last one into the
        `_tail` pair), every env range on its own stream.  Only stream-ordered device work: capturable as a CUDA graph."""
        env, b = self.env, self.buf
        b["obs"][0].copy_(self._tail[0]); b["mask"][0].copy_(self._tail[1])
        if self.fused_rollout:
            # ONE kernel per env-step (msort_rollout_step): step(a_t) and, on the observation tile still on chip, the policy
            # forward + masked draw for step t+1; only the first action of the rollout needs the stand-alone policy kernel.
            # Reward / done go straight into their buffer slots (the env's output pointers are aimed at them).
            T = self.n_steps
            env.rollout_pack(self.flat, out=self._packed_fused)
            env.rollout_policy(self._packed_fused, seed=self.seed, t=t0, obs=b["obs"][0], mask=b["mask"][0],
                               out=(b["act"][0], b["logp"][0], b["val"][0]))
            keep = (env.reward, env.terminated)
            refused = False
            try:
                for t in range(T):
                    last = t + 1 == T
                    oo, om = self._tail if last else (b["obs"][t + 1], b["mask"][t + 1])
                    nxt = self._spare if last else (b["act"][t + 1], b["logp"][t + 1], b["val"][t + 1])
                    env.reward, env.terminated = b["rew"][t], b["done"][t]
                    try:
                        env.rollout_step(b["act"][t], self._packed_fused, self.seed, t0 + t + 1, nxt, out_obs=oo, out_mask=om)
                    except _abi.MsortError as e:
                        if t == 0 and e.code == _abi.E_UNSUPPORTED:   # refused before anything was launched: two kernels per step
                            refused = True
                            break
                        raise
            finally:
                env.reward, env.terminated = keep                     # whatever happens, the env's own output buffers come back
            if refused:
                self.fused_rollout = False
                return self._rollout_body(packed, t0)
            return
        cur = torch.cuda.current_stream(self.dev)
        streams = self._streams or [cur]
        split = self.split_rollout                       # Env_3: the 128-thread policy kernel (msort_rollout_policy) + step
        if split:
            env.rollout_pack(self.flat, out=self._packed_fused)
        for s in self._streams:
            s.wait_stream(cur)
        for t in range(self.n_steps):
            oo, om = self._tail if t + 1 == self.n_steps else (b["obs"][t + 1], b["mask"][t + 1])
            for s, (lo, hi) in zip(streams, self._ranges):
                with torch.cuda.stream(s):
                    rng = None if len(self._ranges) == 1 else (lo, hi)
                    if split:
                        env.rollout_policy(self._packed_fused, seed=self.seed, t=t0 + t, obs=b["obs"][t], mask=b["mask"][t],
                                           out=(b["act"][t], b["logp"][t], b["val"][t]), env_range=rng)
                    else:
                        env.policy_act(packed, seed=self.seed, t=t0 + t, obs=b["obs"][t], mask=b["mask"][t],
                                       out=(b["act"][t], b["logp"][t], b["val"][t]), env_range=rng)
                    env.step(b["act"][t], out_obs=oo, out_mask=om, env_range=rng)   # fused CUDA step (auto-reset inside)
                    b["rew"][t][lo:hi].copy_(env.reward[lo:hi]); b["done"][t][lo:hi].copy_(env.terminated[lo:hi])
        for s in self._streams:
            cur.wait_stream(s)

    @torch.no_grad()
    def collect_rollout(self):
        env, b = self.env, self.buf
        if self._obs is None:
            self._obs, _ = env.reset()
        t0 = self.num_timesteps // self.n
        # zero-copy rollout: step t writes its observation / mask straight into slot t+1 of the buffers
        # (the last step into a spare pair), so nothing but the first slot is ever copied
        direct = self.fused_act and (self.n * self.D * 4) % 16 == 0 and (self.n * self.A) % 16 == 0
        if direct:
            if self._tail is None:
                self._tail = (torch.zeros((self.n, self.D), device=self.dev), torch.zeros((self.n, self.A), dtype=torch.bool, device=self.dev))
                self._tail[0].copy_(self._obs); self._tail[1].copy_(env.action_masks())
                self._packed = torch.zeros(_abi.POLICY_ACT_WEIGHTS, device=self.dev)
                self._packed_fused = torch.zeros(_abi.ROLLOUT_WEIGHTS, dtype=torch.int32, device=self.dev)
                self._spare = (torch.zeros(self.n, dtype=torch.int64, device=self.dev), torch.zeros(self.n, device=self.dev),
                               torch.zeros(self.n, device=self.dev))
                if self.graph_rollout:
                    env.set_option(_abi.OPT_DRAW_COUNTER, self._draw_counter.data_ptr())
            packed = pack_actor_critic(self.policy, out=self._packed)
            if self.graph_rollout and self._warm:
                # ONE CUDA graph of the whole n_steps loop, replayed every rollout: the draw index of step t is
                # t + *draw_counter (MSORT_OPT_DRAW_COUNTER), the weights live in the fixed `_packed` buffer.
                # (the first rollout ran eagerly: it is the warm-up every lazy initialisation needs before a capture)
                self._draw_counter.fill_(t0)
                if self._graph is None:
                    self._graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self._graph):
                        self._rollout_body(packed, 0)
                self._graph.replay()
            else:
                self._draw_counter.zero_()
                self._rollout_body(packed, t0)
                self._warm = True
            return self._finish_rollout(*self._tail)
        packed = pack_actor_critic(self.policy) if self.fused_act else None
        b["obs"][0].copy_(self._obs); b["mask"][0].copy_(env.action_masks())
        for t in range(self.n_steps):
            if self.fused_act:                                           # writes straight into the rollout buffers
                a, _, _ = env.policy_act(packed, seed=self.seed, t=t0 + t, obs=b["obs"][t], mask=b["mask"][t],
                                         out=(b["act"][t], b["logp"][t], b["val"][t]))
            else:
                a, logp, v = self.policy.act(b["obs"][t], b["mask"][t])
                b["act"][t], b["logp"][t], b["val"][t] = a, logp, v
            obs, rew, term, _, _ = env.step(a)
            if t + 1 < self.n_steps:
                b["obs"][t + 1].copy_(obs); b["mask"][t + 1].copy_(env.action_masks())
            b["rew"][t].copy_(rew); b["done"][t].copy_(term)
            self._obs = obs
        return self._finish_rollout(self._obs, env.action_masks())

    def _finish_rollout(self, last_obs, last_mask):
        """Old log-probs / values recomputed in fp32, the bootstrap value of the last observation, GAE(lambda).
        Returns (advantages, returns), both [n_steps, n]."""
        b, T, n = self.buf, self.n_steps, self.n
        self._obs = last_obs
        self.num_timesteps += T * n
        if self.native_update:
            # the update's own fp32 forward over the whole buffer: the importance ratio of epoch 0 is exactly 1 whatever
            # precision the rollout kernel computed its log-probs in (fp16 operands on the tensor cores)
            with torch.cuda.device(self.dev):
                whole = self._batch(T * n, b["obs"], b["mask"], b["act"])
                _abi.check(self.lib, self.lib.msort_ppo_forward(C.byref(whole), _p(self.flat), _p(b["logp"]), _p(b["val"]), self._cstream()),
                           "msort_ppo_forward")
                last = self._batch(n, last_obs, last_mask, b["act"])
                _abi.check(self.lib, self.lib.msort_ppo_forward(C.byref(last), _p(self.flat), None, _p(self._last_v), self._cstream()),
                           "msort_ppo_forward")
                _abi.check(self.lib, self.lib.msort_ppo_gae(T, n, _p(b["rew"]), _p(b["val"]), _p(b["done"]), _p(self._last_v),
                                                            self.gamma, self.lam, _p(self._adv), _p(self._ret), self._cstream()),
                           "msort_ppo_gae")
            return self._adv, self._ret
        with torch.no_grad():
            if self.fused_act:
                for t in range(T):                                       # same purpose, PyTorch fp32
                    b["logp"][t], _, b["val"][t] = self.policy.evaluate(b["obs"][t], b["mask"][t], b["act"][t])
            last_v = self.policy.vf(last_obs).squeeze(1)
            adv = torch.zeros_like(b["rew"])
            gae = torch.zeros(n, device=self.dev)
            for t in reversed(range(T)):                                 # GAE(lambda); `terminated` ends the episode
                nonterm = (~b["done"][t]).float()
                next_v = last_v if t == T - 1 else b["val"][t + 1]
                delta = b["rew"][t] + self.gamma * next_v * nonterm - b["val"][t]
                gae = delta + self.gamma * self.lam * nonterm * gae
                adv[t] = gae
        return adv, adv + b["val"]

    # ------------------------------------------------------------------ update
    def update(self, adv, ret):
        """n_epochs passes over the rollout buffer in shuffled minibatches (SB3 PPO.train with MaskablePPO's masked
        distribution, training.py:118-143).  Native: per epoch ONE `msort_ppo_update` call = per minibatch the advantage
        statistics kernel, the fused forward + loss + backward kernel and the clip + Adam kernel; nothing returns to the
        host until the loss statistics are read."""
        b = self.buf
        N = self.n_steps * self.n
        if self.native_update:
            batch = self._batch(N, b["obs"], b["mask"], b["act"], b["logp"], adv, ret)
            nb = -(-N // self.batch_size)
            with torch.cuda.device(self.dev):
                for _ in range(self.n_epochs):
                    perm = torch.randperm(N, device=self.dev)
                    self._stats.zero_()
                    _abi.check(self.lib, self.lib.msort_ppo_update(C.byref(batch), C.byref(self._hp), _p(self.flat), _p(self._grads),
                                                                  _p(self._m), _p(self._v), _p(self._step), _p(perm), 1,
                                                                  self.batch_size, _p(self._scratch), _p(self._stats), self._cstream()),
                               "msort_ppo_update")
            self._last_stats = (self._stats, nb)
            return {}
        flat = {k: v.reshape(N, *v.shape[2:]) for k, v in b.items()}
        adv, ret = adv.reshape(N), ret.reshape(N)
        stats = {}
        for _ in range(self.n_epochs):
            perm = torch.randperm(N, device=self.dev)
            for s in range(0, N, self.batch_size):
                idx = perm[s:s + self.batch_size]
                logp, ent, v = self.policy.evaluate(flat["obs"][idx], flat["mask"][idx], flat["act"][idx])
                a = adv[idx]
                a = (a - a.mean()) / (a.std() + 1e-8)
                ratio = (logp - flat["logp"][idx]).exp()
                pg = -torch.min(a * ratio, a * ratio.clamp(1 - self.clip, 1 + self.clip)).mean()
                vl = torch.nn.functional.mse_loss(v, ret[idx])
                loss = pg + self.vf_coef * vl - self.ent_coef * ent.mean()
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
                self.opt.step()
            stats = dict(pg=float(pg.detach()), vf=float(vl.detach()), ent=float(ent.detach().mean()))
        return stats

    def last_update_stats(self) -> dict:
        """Loss terms of the last epoch of the last native update (means over that epoch's rows; one device read)."""
        if not self.native_update or getattr(self, "_last_stats", None) is None:
            return {}
        s, nb = self._last_stats
        s = s.tolist()
        # the surrogate sum adds per-minibatch-normalised terms: its mean over minibatches is sum / rows as well
        return dict(pg=s[0] / max(s[4], 1.0), vf=s[1] / max(s[4], 1.0), ent=s[2] / max(s[4], 1.0), clip_fraction=s[3] / max(s[4], 1.0))

    def learn(self, total_timesteps: int, log_every: int = 10):
        it, t0 = 0, time.time()
        while self.num_timesteps < total_timesteps:
            adv, ret = self.collect_rollout()
            st = self.update(adv, ret)
            it += 1
            if it % log_every == 0:
                st = st or self.last_update_stats()
                st.update(timesteps=self.num_timesteps, mean_step_reward=float(self.buf["rew"].mean()),
                          sps=self.num_timesteps / (time.time() - t0))
                self.log.append(st)
        return self

    # ------------------------------------------------------------------ SB3-style archive
    SB3_STATE_KEYS = (("mlp_extractor.policy_net.0", "pi", 0), ("mlp_extractor.policy_net.2", "pi", 2), ("action_net", "pi", 4),
                      ("mlp_extractor.value_net.0", "vf", 0), ("mlp_extractor.value_net.2", "vf", 2), ("value_net", "vf", 4))

    def sb3_state_dict(self) -> dict:
        """The towers under the key names of an SB3 `ActorCriticPolicy.state_dict()` with `net_arch=dict(pi=[32,32], vf=[32,32])`
        (training.py:115): mlp_extractor.policy_net.{0,2}, action_net, mlp_extractor.value_net.{0,2}, value_net."""
        sd = {}
        for name, tower, k in self.SB3_STATE_KEYS:
            lin = getattr(self.policy, tower)[k]
            sd[name + ".weight"] = lin.weight.detach().cpu().clone()
            sd[name + ".bias"] = lin.bias.detach().cpu().clone()
        return sd

    def save(self, path: str) -> str:
        """Write `path`(.zip) in the layout `model.save()` of SB3 produces (ref: training.py:271-287 saves
        `./models/{prefix}_{timesteps}.zip`): a zip holding `policy.pth` (torch state-dict, SB3's key names), `data` (JSON of the
        hyper-parameters) and `_stable_baselines3_version`.  `policy.load_sb3_zip` / `BatchedPressingEnv.set_agents(sort_agent=path)`
        read it; `load()` restores it here.  (SB3 itself is not installed in this image, so loading the archive INTO SB3 is untested.)"""
        import io
        import json
        import zipfile
        if not path.endswith(".zip"):
            path += ".zip"
        buf = io.BytesIO()
        torch.save(self.sb3_state_dict(), buf)
        data = dict(policy_class="MaskableActorCriticPolicy", net_arch=dict(pi=[32, 32], vf=[32, 32]), activation_fn="tanh",
                    observation_dim=self.D, num_actions=self.A, n_steps=self.n_steps, batch_size=self.batch_size, n_epochs=self.n_epochs,
                    gamma=self.gamma, gae_lambda=self.lam, clip_range=self.clip, ent_coef=self.ent_coef, vf_coef=self.vf_coef,
                    learning_rate=self.lr, max_grad_norm=self.max_grad_norm, num_timesteps=self.num_timesteps, seed=self.seed)
        with zipfile.ZipFile(path, "w") as z:
            z.writestr("policy.pth", buf.getvalue())
            z.writestr("data", json.dumps(data))
            z.writestr("_stable_baselines3_version", "msort (SB3 policy.pth key layout)")
        return path

    def load(self, path: str) -> "MaskablePPO":
        """Restore the towers from an archive written by `save()` or by SB3's `model.save()` (its `policy.pth`)."""
        import io
        import zipfile
        with zipfile.ZipFile(path if path.endswith(".zip") else path + ".zip") as z:
            with z.open("policy.pth") as f:
                sd = torch.load(io.BytesIO(f.read()), map_location="cpu", weights_only=True)
        with torch.no_grad():
            for name, tower, k in self.SB3_STATE_KEYS:
                lin = getattr(self.policy, tower)[k]
                lin.weight.copy_(sd[name + ".weight"]); lin.bias.copy_(sd[name + ".bias"])   # (views of self.flat: the kernels see them)
        return self

    # ------------------------------------------------------------------ SB3-style inference
    @torch.no_grad()
    def predict(self, obs, action_masks=None, deterministic=True):
        o = torch.as_tensor(obs, device=self.dev, dtype=torch.float32)
        single = o.dim() == 1
        o = o.reshape(-1, self.D)
        m = (torch.ones((o.shape[0], self.A), dtype=torch.bool, device=self.dev) if action_masks is None
             else torch.as_tensor(action_masks, device=self.dev, dtype=torch.bool).reshape(-1, self.A))
        a, _, _ = self.policy.act(o, m, deterministic=deterministic)
        return (int(a[0]) if single else a), None


@torch.no_grad()
def modular_actions(env, sort_agent=None, press_agent=None, use_action_masking: bool = True, seed: int = 0, t: int = 0):
    """Batched action source of `Env_3_Monolith.step(mode='model')` (env_monolith.py:186-221): the sort agent
    picks the sensor mode from the sorting part of the observation, the press agent the press action from the
    pressing part (under the press mask when masking is on), both shown the plant AFTER update_environment's
    shift (env_monolith.py:114-115) and both deterministic; a missing agent falls back to a uniform draw — over
    the valid press actions when masking is on (:213-219).  action = 11 * sort_mode + press_action.
    `sort_agent` / `press_agent`: objects with SB3's `predict(obs, deterministic=True[, action_masks=])` that
    accept batched CUDA tensors, or None.  Agents trained here (this module's MaskablePPO, 13 -> 2 and 16 -> 11) are
    evaluated IN-BATCH BY THE TENSOR-CORE POLICY KERNEL (`msort_policy_eval`) straight on the 13- and 16-wide column
    slices of the 29-wide observation and the first 11 mask columns — no copy, no eager PyTorch forward."""
    n, dev = env.num_envs, env.device
    obs = env.observe_after_shift()
    g = torch.Generator(device=dev).manual_seed(int(seed) * 1_000_003 + int(t))

    def act(agent, o, m, A):
        pol = getattr(agent, "policy", None)
        if isinstance(pol, MaskableActorCritic) and hasattr(env, "policy_eval") and pol.pi[0].weight.shape[1] == o.shape[1] \
                and pol.pi[4].weight.shape[0] == A and pol.pi[0].weight.is_cuda:
            return env.policy_eval(pack_actor_critic(pol), o, m, num_actions=A, deterministic=True)[0]
        return agent.predict(o, deterministic=True, action_masks=m)[0] if m is not None else agent.predict(o, deterministic=True)[0]

    if sort_agent is not None:
        mode = act(sort_agent, obs[:, :13], None, 2)
    else:
        mode = torch.randint(0, 2, (n,), device=dev, generator=g)
    pmask = env.action_masks()[:, :11]
    if press_agent is not None:
        press = act(press_agent, obs[:, 13:], pmask if use_action_masking else None, 11)
    elif use_action_masking:
        press = torch.multinomial(pmask.float(), 1, generator=g).squeeze(1)
    else:
        press = torch.randint(0, 11, (n,), device=dev, generator=g)
    return mode.to(torch.int64) * 11 + press.to(torch.int64)


def sort_policy_weights(model) -> torch.Tensor:
    """The policy tower of a sort agent trained here (MaskablePPO on BatchedSortingEnv) as the flat 1570-float
    vector Env_2's embedded policy takes (`BatchedPressingEnv.set_agents(sort_agent=...)`, env_2_press.py:39-40)."""
    pi = model.policy.pi if hasattr(model, "policy") else model.pi
    return torch.cat([pi[0].weight.reshape(-1), pi[0].bias, pi[2].weight.reshape(-1), pi[2].bias,
                      pi[4].weight.reshape(-1), pi[4].bias]).detach().float().cpu()


@torch.no_grad()
def evaluate_policy(model, env_cls, n_envs: int = 1024, steps: int = 200, seed: int = 1, **env_kwargs):
    """Mean / std of the cumulative reward of the deterministic masked policy over `n_envs` fresh episodes
    (the reference's protocol: 200 steps, noise 0 — main.py:42-52, benchmark_models.py:126-171)."""
    env = env_cls(n_envs, max_steps=steps, seed=seed, auto_reset=False, device=model.dev, **env_kwargs)
    obs, _ = env.reset()
    total = torch.zeros(n_envs, dtype=torch.float64, device=model.dev)
    for _ in range(steps):
        a, _ = model.predict(obs, action_masks=env.action_masks(), deterministic=True)
        obs, r, term, _, _ = env.step(a)
        total += r.double()
    env.close()
    return total.mean().item(), total.std().item()
