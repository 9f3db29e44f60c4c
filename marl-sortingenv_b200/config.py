"""Plant configuration: the reference's `config.yml` keys + ctor arguments → msort_config_t.

ref: Env_Super.__init__ reads config.yml with yaml.safe_load and the ctor kwargs override
`noise` and `bale_standard_size` (env_super.py:27-133).  Only the LIVE keys
are consumed; dead keys are accepted and ignored, exactly as the reference does.
"""
from __future__ import annotations

import copy
import math

from . import _abi

# The reference's shipped parameter values (config.yml:1-59), restated as a dict so the
# package works without a yaml file; `load_config(path)` overlays a user's config.yml.
DEFAULT_CONFIG = {
    "simulation": {"input_occupancy_min": 60, "input_occupancy_max": 80, "input_batch_size": 100,
                   "steps_per_pattern": 20, "input_history_length": 10},
    "sorting_station": {"baseline_accuracy": [0.75, 0.75, 0.75, 0.75], "boost": 0.5,
                        "occupancy_reduction_factor": 0.2, "noise": 0.05, "stage_capacity": 100},
    "pressing_station": {"press_times": {1: 12, 2: 15}, "container_capacity": 700,
                         "bale_standard_size": 200, "bale_remainder_threshold": 0.5,
                         "bale_quality_thresholds": {"A": 0.9, "B": 0.9, "C": 0.9, "D": 0.9}},
    "rewards": {
        "sorting": {"purity_threshold_theta": 0.80, "decay_steepness_k": 170.0, "min_weight": 0.1,
                    "reward_scaling_factor": 2.0, "sorting_mode_change_penalty": -0.1,
                    "tanh_temperature": 0.5},
        "pressing": {"overflow_penalty_catastrophic": -1.0, "overflow_penalty_severe": -0.5,
                     "overflow_penalty_mild": -0.2, "bale_efficiency_factor": 1,
                     "full_bale_bonus": 0.1, "max_state_reward": 0.5},
        "overflow_termination_penalty": -10.0,
    },
}

# Seasonal patterns are hard-coded in the reference generator (input_generator.py:17-20).
PATTERN_RATIOS = {1: {"A": 0.40, "C": 0.35, "B": 0.15, "D": 0.10},
                  2: {"B": 0.40, "D": 0.35, "A": 0.15, "C": 0.10}}
# Every reset() rebuilds the generator with its ctor default (env_super.py:375,
# input_generator.py:15), and step() before reset() raises in the reference, so the
# config's simulation.steps_per_pattern never takes effect on the hot path.
STEPS_PER_PATTERN_AFTER_RESET = 20
MATERIALS = ("A", "B", "C", "D")


def _merge(base: dict, over: dict) -> dict:
    out = copy.deepcopy(base)
    for k, v in (over or {}).items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


def load_config(config_path: str | None = None) -> dict:
    """yaml.safe_load of a reference-format config.yml (env_super.py:27-29); None → defaults."""
    if config_path is None:
        return copy.deepcopy(DEFAULT_CONFIG)
    import yaml
    with open(config_path, "r") as f:
        user = yaml.safe_load(f)
    return _merge(DEFAULT_CONFIG, user)


def pattern_counts(batch_size: int):
    """floor(ratio*batch) per material for both patterns (input_generator.py:49), float64
    arithmetic identical to the reference's `int(np.floor(ratios[mat] * batchsize))`."""
    out = []
    for key in (1, 2):
        out.append([int(math.floor(PATTERN_RATIOS[key][m] * batch_size)) for m in MATERIALS])
    return out


def make_config(kind: str, num_envs: int, *, max_steps: int = 50, seed: int | None = None,
                noise_sorting: float | None = 0.05, balesize: int | None = 200,
                config: dict | None = None, config_path: str | None = None,
                use_action_masking: bool = True, check_overflow: bool = False,
                auto_reset: bool = True, rng_mode: str = "philox", sort_policy_mlp: bool = False,
                global_env_offset: int = 0) -> _abi.MsortConfig:
    cfgd = config if config is not None else load_config(config_path)
    sim, srt, prs, rew = (cfgd["simulation"], cfgd["sorting_station"], cfgd["pressing_station"],
                          cfgd["rewards"])
    c = _abi.MsortConfig()
    c.struct_size = _abi.C.sizeof(_abi.MsortConfig)
    c.env_kind = _abi.KIND_BY_NAME[kind]
    c.num_envs = int(num_envs)
    c.global_env_offset = int(global_env_offset)
    c.max_steps = int(max_steps)
    flags = 0
    if use_action_masking:
        flags |= _abi.F_ACTION_MASKING
    if check_overflow:
        flags |= _abi.F_CHECK_OVERFLOW
    if auto_reset:
        flags |= _abi.F_AUTO_RESET
    if sort_policy_mlp:
        flags |= _abi.F_SORT_POLICY_MLP
    c.flags = flags
    c.rng_mode = {"philox": _abi.RNG_PHILOX, "replay": _abi.RNG_REPLAY}[rng_mode]
    c.seed = int(seed or 0) & 0xFFFFFFFFFFFFFFFF          # set_seed: `seed or 0` (env_super.py:167)
    c.input_batch_size = int(sim["input_batch_size"])
    c.steps_per_pattern = STEPS_PER_PATTERN_AFTER_RESET
    pc = pattern_counts(c.input_batch_size)
    for p in range(2):
        for m in range(4):
            c.pattern_counts[p][m] = pc[p][m]
    for m in range(4):
        c.baseline_accuracy[m] = float(srt["baseline_accuracy"][m])
        c.quality_threshold[m] = float(prs["bale_quality_thresholds"][MATERIALS[m]])
    c.boost = float(srt["boost"])
    c.noise = float(noise_sorting if noise_sorting is not None else srt["noise"])  # env_super.py:71
    c.stage_capacity = int(srt["stage_capacity"])
    pt = prs["press_times"]
    c.press_time[0] = int(pt[1] if 1 in pt else pt["1"])
    c.press_time[1] = int(pt[2] if 2 in pt else pt["2"])
    c.container_capacity = int(prs["container_capacity"])
    c.bale_size = int(balesize if balesize is not None else prs["bale_standard_size"])  # :87
    c.bale_remainder_threshold = float(prs["bale_remainder_threshold"])
    c.purity_theta = float(rew["sorting"]["purity_threshold_theta"])
    c.purity_scaling = 2.0                                 # hard-coded at env_super.py:971
    c.tanh_temperature = float(rew["sorting"]["tanh_temperature"])
    pr = rew["pressing"]
    c.overflow_penalty_catastrophic = float(pr["overflow_penalty_catastrophic"])
    c.overflow_penalty_severe = float(pr["overflow_penalty_severe"])
    c.overflow_penalty_mild = float(pr["overflow_penalty_mild"])
    c.bale_efficiency_factor = float(pr.get("bale_efficiency_factor", 0.5))  # env_super.py:1059
    c.max_state_reward = float(pr["max_state_reward"])
    c.overflow_termination_penalty = float(rew["overflow_termination_penalty"])
    return c
