"""bench.py's JSON contract where it can be checked without a GPU: the CPU (reference) arm prints one line with the keys the driver
reads and the SAME `config` object the GPU arm prints; the GPU arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_reference_arm_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "path-traced Msamples/s @1080p, 4 bounces" and line["unit"] == "Msamples/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 0 and abs(line["e2e"]["value"] - line["value"]) < 1e-9
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "C2 frame" in cb["sample"]
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.config_dict("c2", bench.CONFIGS["c2"]["passes"], "weak", 1)      # what the GPU arm prints at N = 1


def test_reference_arm_other_ranks_and_configs_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], cwd=ROOT, capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""
    p = _run("--impl", "reference", "--config", "c4")
    assert p.returncode == 0 and "unavailable" in json.loads(p.stdout.strip().splitlines()[-1])


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = _run("--steps", "1", "--warmup", "3")
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)
