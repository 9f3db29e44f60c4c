"""Telemetry for a sampled subset of device envs.

The reference keeps unbounded per-step Python logs on every env: `reward_data`
(env_super.py:402-408, 928-946), `press_actions_per_timestep` (env_super.py:631-637, 729-736,
env_2_press.py:127-131, env_monolith.py:132-138) and the `bale_count` lists
(env_super.py:661-687); `utils/plotting.plot_env` (plotting.py:28-48) renders its dashboard from
them.  On the device those lists are fixed counters, so for the envs one wants to inspect a
`TraceRecorder` takes one snapshot per step instead: `msort_gather_state` copies the plain state
of the K traced envs into a preallocated `[capacity+1, K]` device buffer, next to the step's
info entries (action, applied sort mode / press action, invalid flag, the two reward terms).
`reference_logs()` turns the snapshots of one env back into the reference's own structures —
the per-step lists are read off the snapshots, press log entries and bale lists are rebuilt from
consecutive snapshots with the reference's rules — and `reference_view()` wraps them in an object
with the attributes `plot_env(env)` unpacks.

Rows are recorded after the step kernel, i.e. after an auto-reset: the row of a terminal step shows
the freshly reset plant (its terminal observation is in `info["terminal_observation"]`).  Dashboards
are normally rendered for single episodes (`auto_reset=False`, as `test_env` does, testing.py:28-68).

The conversion (`logs_from_arrays`) is plain numpy and has no device dependency.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np

from . import _abi

MATERIALS = ("A", "B", "C", "D")
ALL5 = MATERIALS + ("E",)
_INFO_KEYS = ("action", "sort_mode", "press_action", "invalid_action", "reward_sort", "reward_press", "sorted_true")


# ------------------------------------------------------------------------------------- conversion
def _press_bale(bales: list, n: int, qk: int, S: int, threshold: float) -> None:
    """ref: Env_Super.press_bale (env_super.py:661-687).  `qk` = the press job's quality in
    hundredths as the state stores it; the reference truncates `int(q*100)` (env_super.py:663)."""
    q = int(np.float64(qk) / np.float64(100.0) * np.float64(100.0))
    full, rem = divmod(int(n), S)
    for _ in range(full):
        bales.append((S, q))
    if rem > 0:
        if rem > S * threshold:
            bales.append((rem, q))
        elif bales:
            bales[-1] = (bales[-1][0] + rem, bales[-1][1])
        else:
            bales.append((rem, q))


def logs_from_arrays(kind: str, snaps: np.ndarray, info: dict, *, bale_size: int,
                     bale_remainder_threshold: float, batch: int = 100) -> dict:
    """Rebuild the reference's logs of ONE env from its snapshots.

    snaps : msort_env_state_t[T+1] — row 0 = state before the first recorded step (after
            reset), row t+1 = state after recorded step t.
    info  : dict of arrays [T]: action, sort_mode, press_action, invalid_action, reward_sort,
            reward_press, sorted_true.
    Returns {"reward_data", "press_actions_per_timestep", "bale_count"} in the reference's layout.
    """
    T = len(snaps) - 1
    rd = {"Accuracy": [], "Setting": [], "Belt_Occupancy": [], "Reward": [], "Belt_Proportions": [], "Total": []}
    for m in ALL5:
        rd[f"{m}_True"], rd[f"{m}_False"] = [], []
    press_log = []
    bales = {m: [] for m in ALL5}
    # bales already pressed before the first recorded step cannot be itemised: carry them as counts
    first = snaps[0]
    for mi, m in enumerate(ALL5):
        if int(first["bale_n"][mi]) > 0:
            raise ValueError("TraceRecorder must start at an episode boundary (bales already exist)")
    for t in range(T):
        prev, cur = snaps[t], snaps[t + 1]
        a = int(info["action"][t])
        pa = int(info["press_action"][t])
        invalid = bool(info["invalid_action"][t])
        # ---- press timers (check_press_status env_super.py:642-659): a press whose timer was 1 finishes
        #      now — unless Env_3 rejected the action, which skips press_action_rules entirely
        #      (env_monolith.py:132-138, 237-243)
        ticked = not (kind == "mono" and invalid)
        timer_after_tick = [0, 0]
        for p in range(2):
            tp = int(prev["press_timer"][p])
            timer_after_tick[p] = tp
            if ticked and tp > 0:
                timer_after_tick[p] = tp - 1
                if tp == 1:
                    _press_bale(bales[ALL5[int(prev["press_mat"][p])]], int(prev["press_n"][p]),
                                int(prev["press_q"][p]), bale_size, bale_remainder_threshold)
        # ---- press log entry of this step
        if invalid:                                   # sanitize_press_action (env_super.py:838-862)
            orig = a % 11 if kind == "mono" else a
            press_id, mat = (1 if orig <= 5 else 2), (orig - 1) % 5
            press_log.append((111 if press_id == 1 else 222, ALL5[mat]))
            if kind == "press":                       # Env_2 still runs press_action_rules((None, None))
                press_log.append((0, None))
        elif pa == 0:
            press_log.append((0, None))               # env_super.py:629-637
        else:
            press_id, mat = (1 if pa <= 5 else 2), (pa - 1) % 5      # press_discrete_to_action :804-809
            if timer_after_tick[press_id - 1] > 0:    # busy press: use_press logs and returns (:725-733)
                press_log.append((111 if press_id == 1 else 222, ALL5[mat]))
            else:
                press_log.append((press_id, mat))     # :736
        # ---- mean purity of this step's sort (sort_material env_super.py:605-607)
        total_input = sum(int(x) for x in cur["sorting"])             # the stage holds what was just sorted
        st4 = int(info["sorted_true"][t]) & 0xFFFFFFFF
        n_true = sum((st4 >> (8 * q)) & 0xFF for q in range(4))
        rd["Accuracy"].append(round(1 - ((total_input - n_true) / total_input), 2) if total_input > 0 else 0)
        # ---- _log_step_data (env_super.py:928-946)
        rs, rp = float(info["reward_sort"][t]), float(info["reward_press"][t])
        rd["Reward"].append((rs, rp))
        rd["Total"].append(rs + rp)
        rd["Setting"].append(int(info["sort_mode"][t]))
        belt = [int(x) for x in cur["belt"]]
        rd["Belt_Occupancy"].append(round(sum(belt) / 100, 2))          # env_super.py:441,456
        tot = sum(belt)
        rd["Belt_Proportions"].append({m: (belt[mi] / tot if tot > 0 else 0) for mi, m in enumerate(MATERIALS)})  # :199-210
        for mi, m in enumerate(MATERIALS):
            rd[f"{m}_True"].append(int(cur["cont_true"][mi]))
            rd[f"{m}_False"].append(int(cur["cont_false"][mi]))
        rd["E_True"].append(int(cur["cont_e"]))
        rd["E_False"].append(0)
    return {"reward_data": rd, "press_actions_per_timestep": press_log, "bale_count": bales}


def reference_view(kind: str, snaps: np.ndarray, info: dict, cfg: _abi.MsortConfig, seed=None) -> SimpleNamespace:
    """An object with the attributes `utils/plotting.plot_env(env)` unpacks (plotting.py:32-48),
    describing the traced env at its last recorded step."""
    logs = logs_from_arrays(kind, snaps, info, bale_size=int(cfg.bale_size),
                            bale_remainder_threshold=float(cfg.bale_remainder_threshold),
                            batch=int(cfg.input_batch_size))
    last = snaps[-1]
    cm = {}
    for mi, m in enumerate(MATERIALS):
        cm[m] = int(last["cont_true"][mi])
        cm[f"{m}_False"] = int(last["cont_false"][mi])
    cm["E"] = int(last["cont_e"])
    ps = {}
    for p in (1, 2):
        busy = int(last["press_timer"][p - 1]) > 0
        ps[f"press_{p}"] = int(last["press_timer"][p - 1])
        ps[f"material_{p}"] = ALL5[int(last["press_mat"][p - 1])] if busy else 0
        ps[f"n_{p}"] = int(last["press_n"][p - 1])
        ps[f"q_{p}"] = int(last["press_q"][p - 1]) / 100
    acc_prev = snaps[-2]["acc_belt"] if len(snaps) > 1 else last["acc_belt"]
    belt = [int(x) for x in last["belt"]]
    return SimpleNamespace(
        current_material_input=[int(x) for x in last["input"]],
        current_material_belt=belt,
        current_material_sorting=[int(x) for x in last["sorting"]],
        container_materials=cm,
        accuracy_belt=[float(x) for x in last["acc_belt"]],
        accuracy_sorter=[float(x) for x in acc_prev],
        sensor_current_setting=int(last["sensor_mode"]),
        reward_data=logs["reward_data"],
        belt_occupancy=round(sum(belt) / 100, 2),
        press_state=ps,
        bale_count=logs["bale_count"],
        bale_standard_size=int(cfg.bale_size),
        quality_thresholds={m: float(cfg.quality_threshold[i]) for i, m in enumerate(MATERIALS)},
        press_actions_per_timestep=logs["press_actions_per_timestep"],
        container_global_max=int(cfg.container_capacity),
        press_times={1: int(cfg.press_time[0]), 2: int(cfg.press_time[1])},
        seed=seed,
        current_step=int(last["step"]),
        material_names=list(MATERIALS),
    )


# ------------------------------------------------------------------------------------- recorder
class TraceRecorder:
    """Per-step snapshots of `env_ids` of a BatchedEnv, taken on the device (one small gather
    launch per step) into buffers preallocated for `capacity` steps."""

    def __init__(self, env, env_ids, capacity: int, growable: bool = False):
        import torch
        self.growable = growable
        missing = [k for k in _INFO_KEYS if k not in env.info_buffers]
        if missing:
            raise ValueError(f"TraceRecorder needs an env created with info_level='full' (missing {missing})")
        self.env = env
        self.capacity = int(capacity)
        self.ids = torch.as_tensor(np.asarray(env_ids, dtype=np.int64), device=env.device).contiguous()
        if self.ids.numel() == 0 or int(self.ids.min()) < 0 or int(self.ids.max()) >= env.num_envs:
            raise ValueError("env_ids must be a non-empty list of indices in [0, num_envs)")
        self.K = int(self.ids.numel())
        self._dt = _abi.env_state_dtype()
        dev = env.device
        self.snaps = torch.zeros((self.capacity + 1, self.K, self._dt.itemsize), dtype=torch.uint8, device=dev)
        self.info = {k: torch.zeros((self.capacity, self.K), dtype=env.info_buffers[k].dtype, device=dev)
                     for k in _INFO_KEYS}
        self.info["reward"] = torch.zeros((self.capacity, self.K), dtype=torch.float32, device=dev)
        self.info["terminated"] = torch.zeros((self.capacity, self.K), dtype=torch.bool, device=dev)
        self.t = 0
        self.started = False

    def _gather(self, row):
        env = self.env
        import torch
        with torch.cuda.device(env.device):
            rc = env.lib.msort_gather_state(env._h, C.c_void_p(env.state.data_ptr()), C.c_void_p(self.ids.data_ptr()),
                                            self.K, C.c_void_p(row.data_ptr()), env._stream())
        _abi.check(env.lib, rc, "msort_gather_state")

    def start(self):
        """Snapshot the traced envs as they are now (call right after `env.reset()`)."""
        self.t = 0
        self._gather(self.snaps[0])
        self.started = True

    def record(self):
        """Append the step that `env.step()` just executed."""
        import torch
        if not self.started:
            raise RuntimeError("TraceRecorder.record() before start()")
        if self.t >= self.capacity:
            if not self.growable:
                raise RuntimeError(f"TraceRecorder capacity ({self.capacity} steps) exhausted")
            self._grow()
        env = self.env
        self._gather(self.snaps[self.t + 1])
        for k in _INFO_KEYS:
            torch.index_select(env.info_buffers[k], 0, self.ids, out=self.info[k][self.t])
        torch.index_select(env.reward, 0, self.ids, out=self.info["reward"][self.t])
        torch.index_select(env.terminated, 0, self.ids, out=self.info["terminated"][self.t])
        self.t += 1

    def _grow(self):
        import torch
        new_cap = 2 * self.capacity
        snaps = torch.zeros((new_cap + 1, self.K, self._dt.itemsize), dtype=torch.uint8, device=self.snaps.device)
        snaps[: self.capacity + 1] = self.snaps
        self.snaps = snaps
        for k, v in self.info.items():
            nv = torch.zeros((new_cap, self.K), dtype=v.dtype, device=v.device)
            nv[: self.capacity] = v
            self.info[k] = nv
        self.capacity = new_cap

    # ---- host side
    def arrays(self, j: int):
        """(snapshots msort_env_state_t[t+1], info dict of [t] arrays) of traced env number j."""
        snaps = self.snaps[: self.t + 1, j].cpu().numpy().copy().view(self._dt).reshape(self.t + 1)
        info = {k: v[: self.t, j].cpu().numpy() for k, v in self.info.items()}
        return snaps, info

    def reference_logs(self, j: int) -> dict:
        snaps, info = self.arrays(j)
        cfg = self.env.cfg
        return logs_from_arrays(self.env.kind, snaps, info, bale_size=int(cfg.bale_size),
                                bale_remainder_threshold=float(cfg.bale_remainder_threshold),
                                batch=int(cfg.input_batch_size))

    def reference_view(self, j: int) -> SimpleNamespace:
        snaps, info = self.arrays(j)
        return reference_view(self.env.kind, snaps, info, self.env.cfg, seed=self.env.seed)
