"""Batched device environments: the Python host over libmsort.so.

`BatchedSortingEnv / BatchedPressingEnv / BatchedMonolithEnv` keep the reference's Gymnasium
surface (ref: env_1_sort.py:19-154, env_2_press.py:19-165, env_monolith.py:22-284) for N
environments at once: same constructor arguments, observation/action spaces, `reset`,
`step`, `action_masks`, info keys and reward terms.  All buffers are torch CUDA tensors owned
here and handed to the C ABI as raw pointers; the computation is the fused CUDA step kernel.
There is no CPU path: constructing an env without a B200 raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi
from .config import make_config
from .spaces import make_spaces

_KIND_NAME = {_abi.ENV_SORT: "sort", _abi.ENV_PRESS: "press", _abi.ENV_MONO: "mono"}


def expand_mask_bits(flags: np.ndarray, num_actions: int) -> np.ndarray:
    """uint16 flag words of `msort_step_host` -> bool [N, A] action masks: bit k = press action k valid
    (press_action_masks, env_super.py:869-885); Env_3's 22-wide mask is that mask twice (monolith_action_masks,
    :887-898); Env_1's two actions are always valid (env_1_sort.py:74-76)."""
    f = np.asarray(flags, dtype=np.uint16)
    m = ((f[:, None] >> np.arange(min(num_actions, 11), dtype=np.uint16)) & 1).astype(bool)
    return np.concatenate([m, m], axis=1) if num_actions == 22 else m


class LazyHostArray:
    """A host result that is produced from the packed flag words only when somebody looks at it."""

    def __init__(self, make, shape):
        self._make, self._value, self.shape = make, None, shape

    @property
    def value(self) -> np.ndarray:
        if self._value is None:
            self._value = self._make()
        return self._value

    def __array__(self, dtype=None, copy=None):
        v = self.value
        return v if dtype is None else v.astype(dtype)

    def __getitem__(self, i):
        return self.value[i]

    def __len__(self):
        return self.shape[0]

    def any(self, *a, **k):
        return self.value.any(*a, **k)

    def all(self, *a, **k):
        return self.value.all(*a, **k)

    def sum(self, *a, **k):
        return self.value.sum(*a, **k)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedEnv:
    """N independent plants stepped by one kernel launch per `step()`."""

    kind = "mono"

    def __init__(self, num_envs: int, device="cuda:0", max_steps: int = 50, seed: int | None = None,
                 noise_sorting: float | None = 0.05, balesize: int | None = 200, simulation=False,
                 config_path: str | None = None, config: dict | None = None,
                 use_action_masking: bool = True, check_overflow: bool = False,
                 auto_reset: bool = True, rng_mode: str = "philox", sort_policy=None,
                 global_env_offset: int = 0, info_level: str = "full", track_stats: bool = True):
        self.lib = _abi.load_library()              # raises when the CUDA extension is missing
        if not torch.cuda.is_available():
            raise RuntimeError("marl_sortingenv_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("marl_sortingenv_b200 runs on CUDA devices only")
        self.num_envs = int(num_envs)
        self.name = self.kind                        # ref: training.py:63 keys on env.name
        self.max_steps = int(max_steps)
        self.seed = seed
        self.cfg = make_config(self.kind, self.num_envs, max_steps=max_steps, seed=seed,
                               noise_sorting=noise_sorting, balesize=balesize, config=config,
                               config_path=config_path, use_action_masking=use_action_masking,
                               check_overflow=check_overflow, auto_reset=auto_reset,
                               rng_mode=rng_mode, sort_policy_mlp=sort_policy is not None,
                               global_env_offset=global_env_offset)
        self.rng_mode = rng_mode
        self.D = _abi.OBS_DIM[self.cfg.env_kind]
        self.A = _abi.NUM_ACTIONS[self.cfg.env_kind]
        self.observation_space, self.action_space = make_spaces(self.kind)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _abi.check(self.lib, self.lib.msort_create(C.byref(self.cfg), dev_index, C.byref(handle)),
                       "msort_create")
        self._h = handle
        n, D, A, dev = self.num_envs, self.D, self.A, self.device
        nbytes = self.lib.msort_state_bytes(self._h)
        self.state = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self.obs = torch.zeros((n, D), dtype=torch.float32, device=dev)
        self.reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self.terminated = torch.zeros(n, dtype=torch.bool, device=dev)
        self.truncated = torch.zeros(n, dtype=torch.bool, device=dev)     # always False (ref: step returns False)
        self.mask = torch.zeros((n, A), dtype=torch.bool, device=dev)
        self.stats = torch.zeros(_abi.NUM_STATS, dtype=torch.float64, device=dev) if track_stats else None
        self._info = _abi.MsortInfoOut()
        self._info.struct_size = C.sizeof(self._info)
        self.info_buffers = {}
        if info_level not in ("none", "episode", "full"):
            raise ValueError("info_level must be 'none', 'episode' or 'full'")
        if info_level in ("episode", "full"):
            self.info_buffers["episode_return"] = torch.zeros(n, dtype=torch.float64, device=dev)
            self.info_buffers["episode_length"] = torch.zeros(n, dtype=torch.int32, device=dev)
            self.info_buffers["terminal_observation"] = torch.zeros((n, D), dtype=torch.float32, device=dev)
        if info_level == "full":
            self.info_buffers["action"] = torch.zeros(n, dtype=torch.int64, device=dev)
            self.info_buffers["overflow"] = torch.zeros(n, dtype=torch.bool, device=dev)
            self.info_buffers["overflow_material"] = torch.full((n,), -1, dtype=torch.int8, device=dev)
            self.info_buffers["sort_mode"] = torch.zeros(n, dtype=torch.uint8, device=dev)
            self.info_buffers["press_action"] = torch.zeros(n, dtype=torch.uint8, device=dev)
            self.info_buffers["invalid_action"] = torch.zeros(n, dtype=torch.bool, device=dev)
            # the two reward terms the reference logs per step (reward_data['Reward'], env_super.py:933)
            self.info_buffers["reward_sort"] = torch.zeros(n, dtype=torch.float32, device=dev)
            self.info_buffers["reward_press"] = torch.zeros(n, dtype=torch.float32, device=dev)
            # units sorted correctly this step, one byte per station (→ logged mean purity, env_super.py:605)
            self.info_buffers["sorted_true"] = torch.zeros(n, dtype=torch.int32, device=dev)
        b = self.info_buffers
        self._info.action = _ptr(b.get("action"))
        self._info.overflow = _ptr(b.get("overflow"))
        self._info.overflow_material = _ptr(b.get("overflow_material"))
        self._info.sort_mode = _ptr(b.get("sort_mode"))
        self._info.press_action = _ptr(b.get("press_action"))
        self._info.invalid_action = _ptr(b.get("invalid_action"))
        self._info.terminal_obs = _ptr(b.get("terminal_observation"))
        self._info.episode_return = _ptr(b.get("episode_return"))
        self._info.episode_length = _ptr(b.get("episode_length"))
        self._info.stats = _ptr(self.stats)
        self._info.reward_sort = _ptr(b.get("reward_sort"))
        self._info.reward_press = _ptr(b.get("reward_press"))
        self._info.sorted_true = _ptr(b.get("sorted_true"))
        self._trace = None
        self._has_info = info_level != "none" or track_stats
        self._was_reset = False
        self.sort_agent = None
        if sort_policy is not None:
            self.set_sort_policy(sort_policy)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.msort_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def unwrapped(self):
        return self

    @property
    def launch_count(self) -> int:
        return int(self.lib.msort_launch_count(self._h))

    @property
    def step_variant(self) -> str:
        """Which instantiation of the step kernel the last step() launched (include/msort.h MSORT_STEP_*)."""
        return ("none", "replay", "generic", "fast", "hot", "hot_persistent", "hot_tensor", "hot_fused", "hot_tensor_split")[int(self.lib.msort_step_variant(self._h))]

    def set_option(self, option: int, value: int):
        """Handle options of include/msort.h (MSORT_OPT_*), e.g. `_abi.OPT_TENSOR_POLICY`."""
        _abi.check(self.lib, self.lib.msort_set_option(self._h, int(option), int(value)), "msort_set_option")

    def get_option(self, option: int) -> int:
        v = C.c_int64(0)
        _abi.check(self.lib, self.lib.msort_get_option(self._h, int(option), C.byref(v)), "msort_get_option")
        return int(v.value)

    def set_flags(self, *, use_action_masking=None, check_overflow=None, auto_reset=None):
        f = int(self.cfg.flags)
        for val, bit in ((use_action_masking, _abi.F_ACTION_MASKING), (check_overflow, _abi.F_CHECK_OVERFLOW),
                         (auto_reset, _abi.F_AUTO_RESET)):
            if val is not None:
                f = (f | bit) if val else (f & ~bit)
        self.cfg.flags = f
        _abi.check(self.lib, self.lib.msort_set_flags(self._h, f), "msort_set_flags")

    # ------------------------------------------------------------------ Gymnasium surface
    def reset(self, seed: int | None = None, options=None, *, which=None, first_pattern=None):
        """ref: Env_X.reset(seed) → (obs, {}).  `seed` re-keys the Philox generator and restarts
        episode numbering; `seed=None` after the first reset keeps the streams running
        (env_super.py:377-378)."""
        flags = 0
        if seed is not None and which is not None:
            # the Philox key belongs to the whole handle and the PHILOX layouts recompute every env's sorter accuracies
            # from (key, env, episode, step): re-keying under a partial reset would change the envs that are NOT reset
            raise ValueError("reset(seed=..., which=...): a new seed re-keys every env of the batch; "
                             "reset the whole batch with the seed, or the subset without one")
        if seed is not None:
            self.seed = seed
            _abi.check(self.lib, self.lib.msort_set_seed(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF), "msort_set_seed")
        elif self._was_reset:
            flags = _abi.RESET_KEEP_STREAMS
        w = None if which is None else torch.as_tensor(which, device=self.device).to(torch.uint8).contiguous()
        fp = None if first_pattern is None else torch.as_tensor(first_pattern, device=self.device).to(torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            rc = self.lib.msort_reset(self._h, _ptr(self.state), _ptr(w), _ptr(fp), _ptr(self.obs),
                                      _ptr(self.mask), flags, self._stream())
        _abi.check(self.lib, rc, "msort_reset")
        self._was_reset = True
        if self._trace is not None and which is None:
            self._trace.start()
        return self.obs, {}

    def step(self, actions, replay: dict | None = None, out_obs: torch.Tensor | None = None,
             out_mask: torch.Tensor | None = None, env_range: tuple[int, int] | None = None):
        """ref: Env_X.step(action) → (obs, reward, terminated, truncated, info), batched.
        `actions`: int64 CUDA tensor [N].  Returned tensors are views of internal buffers that
        the next step() overwrites.  `out_obs` [N,D] f32 / `out_mask` [N,A] bool (contiguous CUDA
        tensors, e.g. slot t+1 of a rollout buffer) receive the new observation / action mask instead
        of the internal buffers — the kernel writes them in place, nothing is copied; `env.obs` /
        `env.mask` / `action_masks()` then refer to those tensors.  `env_range=(first, stop)` steps only the
        envs [first, stop) (`first` a multiple of 128; `msort_step_range`): every tensor is still the whole-batch
        one and only that slice of it is read / written, so disjoint ranges may run on different CUDA streams."""
        if not self._was_reset:
            # the reference raises AttributeError on step() before reset() (env_super.py:394,402)
            raise AttributeError("step() called before reset()")
        if not (isinstance(actions, torch.Tensor) and actions.dtype == torch.int64 and actions.is_cuda
                and actions.is_contiguous()):
            actions = torch.as_tensor(actions, device=self.device).to(torch.int64).contiguous()
        if actions.numel() != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {actions.numel()}")
        rp = None
        if self.rng_mode == "replay":
            rp, _keep = self._make_replay(replay or {})
        for name, t, shape, dt in (("out_obs", out_obs, (self.num_envs, self.D), torch.float32),
                                   ("out_mask", out_mask, (self.num_envs, self.A), torch.bool)):
            if t is not None and not (t.is_cuda and t.is_contiguous() and tuple(t.shape) == shape and t.dtype == dt
                                      and t.data_ptr() % 16 == 0):
                raise ValueError(f"{name} must be a contiguous, 16-byte aligned CUDA tensor of shape {shape}, dtype {dt}")
        if out_obs is not None:
            self.obs = out_obs
        if out_mask is not None:
            self.mask = out_mask
        with torch.cuda.device(self.device):
            if env_range is None:
                rc = self.lib.msort_step(self._h, _ptr(self.state), _ptr(actions), _ptr(self.obs),
                                         _ptr(self.reward), _ptr(self.terminated), _ptr(self.mask),
                                         C.byref(self._info) if self._has_info else None,
                                         C.byref(rp) if rp is not None else None, self._stream())
            else:
                first, stop = int(env_range[0]), int(env_range[1])
                rc = self.lib.msort_step_range(self._h, first, stop - first, _ptr(self.state), _ptr(actions), _ptr(self.obs),
                                               _ptr(self.reward), _ptr(self.terminated), _ptr(self.mask),
                                               C.byref(self._info) if self._has_info else None, self._stream())
        _abi.check(self.lib, rc, "msort_step")
        if self._trace is not None:
            self._trace.record()
        return self.obs, self.reward, self.terminated, self.truncated, self.info_buffers

    def attach_trace(self, env_ids, capacity: int):
        """Record per-step telemetry of the listed envs (the device stand-in for the reference's
        reward_data / press_actions_per_timestep / bale_count logs, env_super.py:928-946): returns a
        `telemetry.TraceRecorder` that `reset()` restarts and every `step()` appends to.  Needs
        `info_level='full'`.  `detach_trace()` stops recording."""
        from .telemetry import TraceRecorder
        self._trace = TraceRecorder(self, env_ids, capacity)
        if self._was_reset:
            self._trace.start()
        return self._trace

    def detach_trace(self):
        self._trace = None

    def action_masks(self):
        """ref: Env_X.action_masks() (env_super.py:869-898) on the current state: bool [N, A]."""
        return self.mask

    def sample_actions(self, seed: int = 0, t: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
        """Uniform random valid action per env under the current mask (one kernel launch).
        ref: Env_3 step(mode='random', use_action_masking=True) (env_monolith.py:152-158)."""
        if out is None:
            out = torch.empty(self.num_envs, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.msort_sample_actions(self._h, _ptr(self.mask), _ptr(out), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                               int(t) & 0xFFFFFFFF, self._stream())
        _abi.check(self.lib, rc, "msort_sample_actions")
        return out

    def generate_streams(self, episode: int, first_step: int, num_steps: int, draw_words: bool = False):
        """The state-independent random inputs of `num_steps` steps of one episode (`msort_generate_streams`): dict of CUDA
        tensors input_counts [T,N] int32 (packed bytes), noise_u [T,N,4] f64, first_pattern [N] uint8 and, on request,
        draw_words [T,N,12] int32 — in the form `step(..., replay=...)` of a REPLAY-mode env consumes.
        ref: SeasonalInputGenerator.generate_input (input_generator.py:37-64), rng_noise.uniform (env_super.py:508)."""
        T, n, dev = int(num_steps), self.num_envs, self.device
        out = dict(input_counts=torch.empty((T, n), dtype=torch.int32, device=dev),
                   noise_u=torch.empty((T, n, 4), dtype=torch.float64, device=dev),
                   first_pattern=torch.empty(n, dtype=torch.uint8, device=dev))
        if draw_words:
            out["draw_words"] = torch.empty((T, n, 12), dtype=torch.int32, device=dev)
        with torch.cuda.device(self.device):
            rc = self.lib.msort_generate_streams(self._h, int(episode), int(first_step), T, _ptr(out["input_counts"]),
                                                 _ptr(out["noise_u"]), _ptr(out.get("draw_words")), _ptr(out["first_pattern"]),
                                                 self._stream())
        _abi.check(self.lib, rc, "msort_generate_streams")
        return out

    def rule_based_actions(self, after_shift: bool = True, out: torch.Tensor | None = None) -> torch.Tensor:
        """The reference's heuristic policy for every env (one kernel launch): sorting_rules() +
        check_container_level() combined as Env_3.step(mode='rule_based') does
        (env_monolith.py:166-184, env_super.py:469-482, 689-720)."""
        if out is None:
            out = torch.empty(self.num_envs, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.msort_rule_based_actions(self._h, _ptr(self.state), 1 if after_shift else 0, _ptr(out),
                                                   self._stream())
        _abi.check(self.lib, rc, "msort_rule_based_actions")
        return out

    def policy_act(self, packed_weights: torch.Tensor, seed: int = 0, t: int = 0, deterministic: bool = False,
                   obs: torch.Tensor | None = None, mask: torch.Tensor | None = None, out=None,
                   env_range: tuple[int, int] | None = None):
        """Fused actor-critic inference + masked categorical draw for every env (one tcgen05 kernel
        launch; `msort_policy_act`): (actions int64 [N], log-prob f32 [N], value f32 [N]) for the
        current observation / action mask (or the given ones).  `packed_weights` comes from
        `ppo.pack_actor_critic`.  ref: MaskablePPO's policy forward during collect_rollouts
        (training.py:118-143, net_arch pi=[32,32], vf=[32,32]).  `env_range=(first, stop)`: only those envs
        (whole-batch tensors, that slice read / written; the draw is keyed by the global env id)."""
        obs = self.obs if obs is None else obs
        mask = self.mask if mask is None else mask
        if packed_weights.numel() != _abi.POLICY_ACT_WEIGHTS or packed_weights.dtype != torch.float32:
            raise ValueError(f"packed_weights must hold {_abi.POLICY_ACT_WEIGHTS} float32 values")
        if not (obs.is_contiguous() and mask.is_contiguous() and packed_weights.is_contiguous()):
            raise ValueError("policy_act needs contiguous tensors")
        if out is None:
            out = (torch.empty(self.num_envs, dtype=torch.int64, device=self.device),
                   torch.empty(self.num_envs, dtype=torch.float32, device=self.device),
                   torch.empty(self.num_envs, dtype=torch.float32, device=self.device))
        a, lp, v = out
        with torch.cuda.device(self.device):
            if env_range is None:
                rc = self.lib.msort_policy_act(self._h, _ptr(obs), _ptr(mask), _ptr(packed_weights),
                                               int(seed) & 0xFFFFFFFFFFFFFFFF, int(t) & 0xFFFFFFFF, 1 if deterministic else 0,
                                               _ptr(a), _ptr(lp), _ptr(v), self._stream())
            else:
                first, stop = int(env_range[0]), int(env_range[1])
                rc = self.lib.msort_policy_act_range(self._h, first, stop - first, _ptr(obs), _ptr(mask), _ptr(packed_weights),
                                                     int(seed) & 0xFFFFFFFFFFFFFFFF, int(t) & 0xFFFFFFFF,
                                                     1 if deterministic else 0, _ptr(a), _ptr(lp), _ptr(v), self._stream())
        _abi.check(self.lib, rc, "msort_policy_act")
        return a, lp, v

    def policy_eval(self, packed_weights: torch.Tensor, obs: torch.Tensor, mask: torch.Tensor | None = None, *, num_actions: int,
                    deterministic: bool = True, seed: int = 0, t: int = 0, out=None):
        """`policy_act`'s kernels on caller-given rows (`msort_policy_eval`): `obs` [K, D] f32 may be a column slice of a wider
        tensor (row stride > D, e.g. `env.obs[:, :13]`), `mask` [K, A] bool / uint8 likewise or None (every action valid).
        (D, A) must be one of (13, 2), (16, 11), (29, 22).  Returns (actions int64 [K], log-prob f32 [K], value f32 [K]).
        ref: the sort / press agents' `predict` inside Env_3.step(mode='model') (env_monolith.py:186-221)."""
        if obs.dim() != 2 or obs.dtype != torch.float32 or not obs.is_cuda or obs.stride(1) != 1:
            raise ValueError("policy_eval needs a 2-D float32 CUDA observation tensor with unit column stride")
        K, D = obs.shape
        if mask is not None and (mask.dim() != 2 or mask.shape != (K, num_actions) or mask.stride(1) != 1 or mask.element_size() != 1):
            raise ValueError("policy_eval: mask must be [K, A] bool / uint8 with unit column stride")
        if out is None:
            out = (torch.empty(K, dtype=torch.int64, device=self.device), torch.empty(K, dtype=torch.float32, device=self.device),
                   torch.empty(K, dtype=torch.float32, device=self.device))
        a, lp, v = out
        with torch.cuda.device(self.device):
            rc = self.lib.msort_policy_eval(self._h, int(D), int(num_actions), int(K), _ptr(obs), int(obs.stride(0)),
                                            _ptr(mask), 0 if mask is None else int(mask.stride(0)), _ptr(packed_weights),
                                            int(seed) & 0xFFFFFFFFFFFFFFFF, int(t) & 0xFFFFFFFF, 1 if deterministic else 0,
                                            _ptr(a), _ptr(lp), _ptr(v), self._stream())
        _abi.check(self.lib, rc, "msort_policy_eval")
        return a, lp, v

    def rollout_pack(self, params: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """The flat fp32 actor-critic parameter vector (`ppo.flatten_parameters`, the layout of `msort_ppo_*`) packed for the
        fused rollout kernel (`msort_rollout_pack`, one small kernel): `_abi.ROLLOUT_WEIGHTS` 32-bit words."""
        if out is None:
            out = torch.empty(_abi.ROLLOUT_WEIGHTS, dtype=torch.int32, device=self.device)
        if not (params.is_cuda and params.dtype == torch.float32 and params.is_contiguous()):
            raise ValueError("rollout_pack needs a contiguous float32 CUDA parameter vector")
        with torch.cuda.device(self.device):
            rc = self.lib.msort_rollout_pack(_ptr(params), _ptr(out), self._stream())
        _abi.check(self.lib, rc, "msort_rollout_pack")
        return out

    def rollout_policy(self, packed: torch.Tensor, seed: int = 0, t: int = 0, deterministic: bool = False,
                       obs: torch.Tensor | None = None, mask: torch.Tensor | None = None, out=None,
                       env_range: tuple[int, int] | None = None):
        """`policy_act` in the fused kernel's arithmetic and weights format (`rollout_pack`): actor-critic forward + masked
        categorical draw for every env of Env_3 (`msort_rollout_policy`: one CTA per 128 envs, 8 CTAs per SM).  Returns
        (actions int64 [N], log-prob f32 [N], value f32 [N])."""
        obs = self.obs if obs is None else obs
        mask = self.mask if mask is None else mask
        if not (obs.is_contiguous() and mask.is_contiguous() and packed.is_contiguous()):
            raise ValueError("rollout_policy needs contiguous tensors")
        if out is None:
            out = (torch.empty(self.num_envs, dtype=torch.int64, device=self.device),
                   torch.empty(self.num_envs, dtype=torch.float32, device=self.device),
                   torch.empty(self.num_envs, dtype=torch.float32, device=self.device))
        a, lp, v = out
        first, stop = (0, self.num_envs) if env_range is None else (int(env_range[0]), int(env_range[1]))
        with torch.cuda.device(self.device):
            rc = self.lib.msort_rollout_policy(self._h, first, stop - first, _ptr(obs), _ptr(mask), _ptr(packed),
                                               int(seed) & 0xFFFFFFFFFFFFFFFF, int(t) & 0xFFFFFFFF, 1 if deterministic else 0,
                                               _ptr(a), _ptr(lp), _ptr(v), self._stream())
        _abi.check(self.lib, rc, "msort_rollout_policy")
        return a, lp, v

    def rollout_step(self, actions: torch.Tensor, packed: torch.Tensor, seed: int, t: int, next_out,
                     out_obs: torch.Tensor | None = None, out_mask: torch.Tensor | None = None, deterministic: bool = False):
        """One env-step of the MaskablePPO rollout loop as ONE kernel (`msort_rollout_step`; Env_3, training configuration):
        step(actions) AND the actor-critic forward + masked categorical draw for the NEXT step on the observation / mask tile
        the step has just built on chip.  `next_out` = (actions int64 [N], log-prob f32 [N], value f32 [N]) receives what
        `policy_act(obs, mask, seed=seed, t=t)` would return for the new observation (up to the kernels' rounding).
        Returns what step() returns.  ref: training.py:118-143 (sb3 collect_rollouts)."""
        if not self._was_reset:
            raise AttributeError("step() called before reset()")
        if self.kind != "mono":
            raise ValueError("rollout_step exists for Env_3_Monolith only")
        for name, t_, shape, dt in (("out_obs", out_obs, (self.num_envs, self.D), torch.float32),
                                    ("out_mask", out_mask, (self.num_envs, self.A), torch.bool)):
            if t_ is not None and not (t_.is_cuda and t_.is_contiguous() and tuple(t_.shape) == shape and t_.dtype == dt
                                       and t_.data_ptr() % 16 == 0):
                raise ValueError(f"{name} must be a contiguous, 16-byte aligned CUDA tensor of shape {shape}, dtype {dt}")
        if out_obs is not None:
            self.obs = out_obs
        if out_mask is not None:
            self.mask = out_mask
        na, nlp, nv = next_out
        with torch.cuda.device(self.device):
            rc = self.lib.msort_rollout_step(self._h, _ptr(self.state), _ptr(actions), _ptr(self.obs), _ptr(self.reward),
                                             _ptr(self.terminated), _ptr(self.mask), C.byref(self._info) if self._has_info else None,
                                             _ptr(packed), int(seed) & 0xFFFFFFFFFFFFFFFF, int(t) & 0xFFFFFFFF,
                                             1 if deterministic else 0, _ptr(na), _ptr(nlp), _ptr(nv), self._stream())
        _abi.check(self.lib, rc, "msort_rollout_step")
        return self.obs, self.reward, self.terminated, self.truncated, self.info_buffers

    # ------------------------------------------------------------------ host-buffer surface
    def _host_buffers(self):
        if getattr(self, "_hb", None) is None:
            n, D = self.num_envs, self.D
            pin = dict(pin_memory=True)
            self._hb = dict(actions=torch.zeros(n, dtype=torch.uint8, **pin),
                            obs=torch.zeros((n, D), dtype=torch.float32, **pin),
                            reward=torch.zeros(n, dtype=torch.float32, **pin),
                            flags=torch.zeros(n, dtype=torch.int16, **pin))
            self._host_scratch = torch.zeros(int(self.lib.msort_host_scratch_bytes(self._h)), dtype=torch.uint8, device=self.device)
            assert self._host_scratch.data_ptr() % 256 == 0
            self._hio = _abi.MsortHostIO()
            self._hio.struct_size = C.sizeof(self._hio)
            self._hio.obs, self._hio.reward, self._hio.flags = (C.c_void_p(self._hb[k].data_ptr()) for k in ("obs", "reward", "flags"))
            self._hb_np = {k: v.numpy() for k, v in self._hb.items()}
            self._hb_np["flags"] = self._hb_np["flags"].view(np.uint16)
        return self._hb

    def step_host(self, actions, chunks: int = 0):
        """step() for callers that live on the host (SB3-style loops; the VecEnv contract of training.py:64-69).
        `actions`: numpy array / CPU tensor [N] of any integer dtype — a pinned uint8 or int64 tensor is used in place,
        anything else is converted into the env's pinned uint8 buffer.  One native call (`msort_step_host`) cuts the batch
        into `chunks` env ranges (0 = the library's default) and pipelines, on the library's own streams, the H2D copy of
        a range's actions, its step kernel and the D2H copies of its results, so both PCIe directions and the kernel
        overlap.  Back come numpy views of pinned host buffers: obs [N,D] f32, reward [N] f32, and two LAZY arrays —
        `terminated` [N] bool and the action mask [N,A] bool — that expand the 16-bit flag word per env the device sent
        (11 mask bits + the done bit) only when they are read (`np.asarray(x)`, indexing); `self.host_flags` is the raw
        uint16 array.  Returns (obs, reward, terminated, truncated, mask); `self.h2d_bytes` / `self.d2h_bytes` count the
        bytes that crossed PCIe."""
        if not self._was_reset:
            raise AttributeError("step() called before reset()")
        hb = self._host_buffers()
        a = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions))
        a = a.reshape(-1)
        if a.numel() != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {a.numel()}")
        io = self._hio
        io.chunks = int(chunks)
        io.dev_obs, io.dev_reward, io.dev_terminated, io.dev_mask = (C.c_void_p(t.data_ptr()) for t in (self.obs, self.reward, self.terminated, self.mask))
        if a.is_pinned() and a.is_contiguous() and a.dtype in (torch.uint8, torch.int64):
            src = a                                    # the caller already holds a pinned buffer
        else:
            hb["actions"].copy_(a)
            src = hb["actions"]
        if src.dtype == torch.uint8:
            io.actions_u8, io.actions_i64 = C.c_void_p(src.data_ptr()), None
        else:
            io.actions_u8, io.actions_i64 = None, C.c_void_p(src.data_ptr())
        with torch.cuda.device(self.device):
            rc = self.lib.msort_step_host(self._h, _ptr(self.state), _ptr(self._host_scratch), C.byref(io),
                                          C.byref(self._info) if self._has_info else None, self._stream())
        _abi.check(self.lib, rc, "msort_step_host")
        torch.cuda.current_stream(self.device).synchronize()
        self.h2d_bytes = src.numel() * src.element_size()
        self.d2h_bytes = sum(hb[k].numel() * hb[k].element_size() for k in ("obs", "reward", "flags"))
        flags = self.host_flags = self._hb_np["flags"]
        return (self._hb_np["obs"], self._hb_np["reward"], LazyHostArray(lambda: (flags >> 15).astype(bool), (self.num_envs,)),
                np.zeros(self.num_envs, dtype=bool), LazyHostArray(lambda: expand_mask_bits(flags, self.A), (self.num_envs, self.A)))

    def get_obs(self):
        with torch.cuda.device(self.device):
            rc = self.lib.msort_observe(self._h, _ptr(self.state), _ptr(self.obs), _ptr(self.mask), self._stream())
        _abi.check(self.lib, rc, "msort_observe")
        return self.obs

    def observe_after_shift(self, out: torch.Tensor | None = None) -> torch.Tensor:
        """The observation the agents of `Env_3_Monolith.step(mode='model')` are shown (env_monolith.py:114-115,
        186-221): the plant after update_environment has moved input -> belt -> sorting, before anything else
        of the coming step.  Returns a float32 [N, D] tensor of its own (`out` or a cached buffer); `env.obs`
        and the state are untouched."""
        if out is None:
            if getattr(self, "_obs_shift", None) is None:
                self._obs_shift = torch.empty((self.num_envs, self.D), dtype=torch.float32, device=self.device)
            out = self._obs_shift
        with torch.cuda.device(self.device):
            rc = self.lib.msort_observe_after_shift(self._h, _ptr(self.state), _ptr(out), None, self._stream())
        _abi.check(self.lib, rc, "msort_observe_after_shift")
        return out

    # ------------------------------------------------------------------ embedded sort agent (Env_2)
    def set_agents(self, sort_agent=None, **_):
        """ref: Env_2_Pressing.set_agents(sort_agent) (env_2_press.py:39-40).  Accepts a flat
        1570-float weight vector, or an object exposing SB3's `policy.state_dict()` layout."""
        self.set_sort_policy(sort_agent)

    def set_sort_policy(self, agent):
        from .policy import flatten_sort_policy
        self.sort_agent = agent
        if agent is None:
            self.set_flags_mlp(False)
            return
        w = flatten_sort_policy(agent).to(device=self.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(self.device):
            rc = self.lib.msort_set_policy(self._h, _ptr(w), 1, self._stream())
        _abi.check(self.lib, rc, "msort_set_policy")
        torch.cuda.current_stream(self.device).synchronize()   # `w` may be freed after return
        self.set_flags_mlp(True)

    def policy_logits_tensor(self, sort_obs: torch.Tensor) -> torch.Tensor:
        """Diagnostics: the two logits of the embedded sort policy for `sort_obs` [K,13] f32 (CUDA), computed by the
        tensor-core form the HOT_TENSOR step kernel uses (`msort_debug_policy_logits`)."""
        o = sort_obs.to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty((o.shape[0], 2), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.msort_debug_policy_logits(self._h, _ptr(o), o.shape[0], _ptr(out), self._stream())
        _abi.check(self.lib, rc, "msort_debug_policy_logits")
        return out

    def set_flags_mlp(self, on: bool):
        f = int(self.cfg.flags)
        f = (f | _abi.F_SORT_POLICY_MLP) if on else (f & ~_abi.F_SORT_POLICY_MLP)
        self.cfg.flags = f
        _abi.check(self.lib, self.lib.msort_set_flags(self._h, f), "msort_set_flags")

    # ------------------------------------------------------------------ replay / state exchange
    def _make_replay(self, r: dict):
        rp = _abi.MsortReplay()
        rp.struct_size = C.sizeof(rp)
        keep = []

        def dev(x, dtype, shape):
            t = torch.as_tensor(x, device=self.device)
            t = t.to(dtype).reshape(shape).contiguous()
            keep.append(t)
            return t
        n = self.num_envs
        if "noise_u" in r:
            rp.noise_u = _ptr(dev(r["noise_u"], torch.float64, (n, 4)))
        if "redis_u" in r:
            t = r["redis_u"]
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
                t = dev(t, torch.float64, (n, -1))
            rp.redis_u = _ptr(t)
            rp.redis_len = t.shape[1]
        if r.get("input_counts") is not None:
            ic = r["input_counts"]
            if isinstance(ic, torch.Tensor):
                ic = ic.detach().cpu().numpy()
            ic = np.ascontiguousarray(np.asarray(ic).astype(np.uint32)).view(np.int32).reshape(n)
            t = torch.from_numpy(ic.copy()).to(self.device)
            keep.append(t)
            rp.input_counts = _ptr(t)
        if r.get("press_choice") is not None:
            rp.press_choice = _ptr(dev(r["press_choice"], torch.uint8, (n,)))
        if r.get("sort_mode") is not None:
            rp.sort_mode = _ptr(dev(r["sort_mode"], torch.uint8, (n,)))
        self._replay_keep = keep
        return rp, keep

    def export_state(self) -> np.ndarray:
        """SoA device blob → numpy structured array of msort_env_state_t (synchronises)."""
        dt = _abi.env_state_dtype()
        buf = torch.zeros(self.num_envs * dt.itemsize, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.msort_export_state(self._h, _ptr(self.state), _ptr(buf), self._stream())
        _abi.check(self.lib, rc, "msort_export_state")
        return buf.cpu().numpy().view(dt).copy()

    def import_state(self, arr: np.ndarray):
        dt = _abi.env_state_dtype()
        a = np.ascontiguousarray(arr, dtype=dt)
        assert a.shape == (self.num_envs,)
        buf = torch.from_numpy(a.view(np.uint8).copy()).to(self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.msort_import_state(self._h, _ptr(self.state), _ptr(buf), self._stream())
        _abi.check(self.lib, rc, "msort_import_state")
        torch.cuda.current_stream(self.device).synchronize()
        self._was_reset = True
        self.get_obs()

    def state_stats(self) -> torch.Tensor:
        """16 state-wide sums (include/msort.h msort_reduce_stats) as a CUDA f64 tensor."""
        out = torch.zeros(_abi.NUM_STATS, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.msort_reduce_stats(self._h, _ptr(self.state), _ptr(out), self._stream())
        _abi.check(self.lib, rc, "msort_reduce_stats")
        return out

    def sync_check(self):
        _abi.check(self.lib, self.lib.msort_sync_check(self._h, self._stream()), "msort_sync_check")


class BatchedSortingEnv(BatchedEnv):
    """ref: Env_1_Sorting (env_1_sort.py:12-154): Box(13) / Discrete(2)."""
    kind = "sort"


class BatchedPressingEnv(BatchedEnv):
    """ref: Env_2_Pressing (env_2_press.py:12-165): Box(16) / Discrete(11)."""
    kind = "press"


class BatchedMonolithEnv(BatchedEnv):
    """ref: Env_3_Monolith (env_monolith.py:12-284): Box(29) / Discrete(22)."""
    kind = "mono"


ENV_CLASSES = {"sort": BatchedSortingEnv, "press": BatchedPressingEnv, "mono": BatchedMonolithEnv}
