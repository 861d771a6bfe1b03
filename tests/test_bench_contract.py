"""bench.py contract on a machine without a GPU: the reference arm (the oracle's port of the training iteration on
the host cores) prints ONE well-formed JSON line, and the product arm refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"]


def run(*argv, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *argv], capture_output=True, text=True,
                          timeout=timeout, cwd=REPO, env=env)


def test_reference_arm_prints_one_contract_line():
    p = run("--impl", "reference", "--steps", "3", "--warmup", "3")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "img/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 3
    assert d["vs_baseline"] is None and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_product_arm_needs_cuda():
    p = run("--steps", "1", "--warmup", "3", "--no-cpu-baseline", timeout=300)
    assert p.returncode != 0
    assert "CUDA" in (p.stderr + p.stdout)
    assert not [l for l in p.stdout.strip().splitlines() if l.startswith("{")]
