"""bench.py's reference arm runs on host cores only: check its JSON line against the driver contract here (CPU)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


import pytest


@pytest.mark.parametrize("port", [False, True])
def test_reference_arm_prints_one_contract_line(port):
    """kind 'reference' = the unmodified reference (present here, vendored to oracle/_ref/reference for the GPU box),
    kind 'port' = the oracle port (DPM_BENCH_PORT=1, or when the reference did not travel)"""
    env = dict(os.environ)
    if port:
        env["DPM_BENCH_PORT"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--points", "4096", "--cpu-frames", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    from oracle import ref_loader
    want = "port" if (port or ref_loader.ref_root() is None) else "reference"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == want and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
