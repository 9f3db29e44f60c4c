"""bench.py --impl reference (the CPU arm: the unmodified Python reference where a copy of it exists — /root/reference or
oracle/_ref — else the C oracle port; the port's number is always reported beside it) prints exactly ONE JSON line on
stdout with the contract's keys — also under a multi-rank launch, where only rank 0 may print."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "4", "--warmup", "3"],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == 4 and d["warmup"] == 3 and d["value"] > 0
    from oracle.ref_loader import reference_available
    assert d["cpu_baseline"]["kind"] == ("reference" if reference_available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_port"]["kind"] == "port" and d["cpu_port"]["value"] > 0
    if d["cpu_baseline"]["kind"] == "reference":
        assert d["cpu_port"]["value"] > 20 * d["value"]          # the C restatement is orders of magnitude faster than the Python reference
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["env"] == "Env_3_Monolith" and d["config"]["envs_per_gpu"] == 1048576 and d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""
