"""CPU suite: bench.py's reference arm (the reference's CPU formulation via the oracle) prints
one JSON line with the contract's keys; non-zero ranks print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + extra,
                         capture_output=True, text=True, env=e, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip()


def test_reference_arm_json_contract():
    line = _run(["--steps", "1", "--warmup", "0", "--n", "300", "--cpu-sample", "16"])
    d = json.loads(line.splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "moves/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    assert _run(["--steps", "1", "--warmup", "0", "--n", "100", "--cpu-sample", "4"],
                env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == ""


def test_scheduling_and_ttb_reference_arms():
    d = json.loads(_run(["--workload", "es50", "--steps", "1", "--warmup", "0"]).splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and "time_to_zero_hard" in d
    d = json.loads(_run(["--workload", "nq64", "--n", "16", "--steps", "2"]).splitlines()[-1])
    assert d["impl"] == "reference" and d["best_score"] >= 0 and d["value"] > 0
