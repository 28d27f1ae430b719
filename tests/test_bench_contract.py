"""bench.py's output contract, checked where it can run without a GPU: the reference arm (the unmodified reference, or
the oracle port, on the host cores) prints exactly ONE JSON line with the keys the driver reads, and the GPU arm
refuses to run -- loudly, not with a CPU fallback -- when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("workload", ["dxt1_rgba8", "etc1_rgb8"])
def test_reference_arm_prints_one_json_line(workload):
    res = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", workload)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout[:2000]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "Mpixels/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["config"]["workload"] == workload
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_gpu_arm_refuses_without_a_device():
    res = _run("--steps", "1", "--warmup", "1")
    assert res.returncode != 0 and res.stdout.strip() == ""
    assert "no CUDA device" in res.stderr and "no CPU path" in res.stderr
