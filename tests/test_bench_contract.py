"""CPU: the bench.py contract that does not need a GPU -- the reference arm prints ONE JSON
line with the agreed keys on the same config as the GPU arm, and the GPU arm fails loudly
(no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args],
                          capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_json_line():
    res = run_bench("--impl", "reference", "--steps", "2", "--warmup", "3", "--preroll", "0")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["metric"] == "PT-TEMPO steps/s at dkmax=200, epsrel=1e-9"
    assert d["config"]["workload"].startswith("spin-boson PT-TEMPO process tensor build")
    assert d["config"]["timed_steps"] == [5, 6] and d["steps"] == 2 and d["warmup"] == 3
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["n_gpus"] == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_gpu_arm_fails_loudly_without_cuda():
    res = run_bench("--steps", "1", "--warmup", "3", "--preroll", "0")
    assert res.returncode != 0
    assert "no CPU fallback" in res.stderr
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
