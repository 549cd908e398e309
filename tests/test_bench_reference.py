"""The reference arm of bench.py (the reference's CPU path = the oracle port, no GPU involved) prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "iter/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "C3" in line["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_state_deviation_figure():
    """The parity figure bench.py adds to its line (`solution.max_rel_state_deviation_vs_cpu_port`): per state group, relative to
    the group's largest entry."""
    import numpy as np

    import bench
    from sadvio_b200 import abi

    a, b = abi.Delta.zeros(3, 2), abi.Delta.zeros(3, 2)
    b.dpose[:] = 1.0
    b.dv[:] = 1e-3
    a.dpose[:] = 1.0
    a.dv[:] = 1e-3
    assert bench.state_deviation(a, b) == 0.0
    a.dv[1, 2] += 1e-9                       # 1e-6 of the velocity group's scale, although tiny next to the poses
    assert abs(bench.state_deviation(a, b) - 1e-6) < 1e-12
    a.dlmk[:] = 5.0                          # landmarks are not part of the state figure
    assert abs(bench.state_deviation(a, b) - 1e-6) < 1e-12
