"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the keys the
driver reads, and the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle", "_ref", "gcm3d_ref")


def _run(args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    if not os.path.exists(ORACLE):
        pytest.skip("oracle/_ref is not built here")
    out = _run(["--impl", "reference", "--ref-size", "16", "--steps", "1", "--warmup", "1"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "tet node-updates/sec" and d["unit"] == "node-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "reference"
    assert d["cpu_baseline"]["cores"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "sample" in d["config"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = _run(["--size", "8", "--zones", "1", "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-e2e"], timeout=300)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]
