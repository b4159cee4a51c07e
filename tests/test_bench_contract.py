"""The driver's bench contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line
with the agreed keys, and the B200 arm refuses to run (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    proc = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "audio_hours_per_sec_tokenized"
    assert line["unit"] == "audio-hours/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["steps"] == 1
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample", "host"}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "audio-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["gpu_launches"] == 0


def test_reference_arm_on_other_ranks_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, cwd=ROOT, env=env)
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("needs a machine without a CUDA device")
    proc = _run("--steps", "1", "--warmup", "3", "--no-e2e", "--no-cpu-baseline", timeout=300)
    assert proc.returncode != 0 and "CUDA" in proc.stderr
    assert proc.stdout.strip() == ""
