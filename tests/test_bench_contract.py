"""CPU: the bench.py contract that can be checked without a GPU -- the reference arm (the oracle port on the host
cores) prints ONE JSON line with the agreed keys, and the product arm refuses to run without CUDA instead of falling
back to anything."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "train images/sec (fwd+bwd)" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None  # BASELINE.md holds no published number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_product_arm_needs_cuda():
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box the product arm is exercised by the driver itself
    r = _run("--steps", "1", "--warmup", "1", "--no-cpu", "--no-eager", timeout=300)
    assert r.returncode != 0  # no CPU fallback: it raises
    assert not any(l.startswith("{") and '"value"' in l for l in r.stdout.splitlines())
