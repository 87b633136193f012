"""CPU tests of bench.py's driver contract: the reference arm runs entirely on the host and prints
one well-formed JSON line; the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 3
    assert d["metric"] == "set-abstraction points/sec" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and abs(d["value"] - 32 * 4096 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"]
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_product_arm_refuses_to_run_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "1", timeout=300)
    assert r.returncode != 0
    assert "CUDA" in r.stderr and not r.stdout.strip()
