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


def test_hbm_kernel_table_reports_the_fused_patch_embedding():
    """post-processing of the profiler's per-class times into achieved GB/s (pure host code): with the one-kernel patch
    embedding the row `embed_fused` replaces `im2col` / `tdl`, whose passes over HBM no longer exist"""
    sys.path.insert(0, str(ROOT))
    import bench

    rec = json.loads((ROOT / "profiles" / "r2_bench_n1_jumpcp.json").read_text().strip().splitlines()[-1])
    prof = {k: {"ms_per_step": v, "launches_per_step": 1.0} for k, v in rec["kernel_breakdown_ms_per_step"].items()}
    assert "im2col" not in prof  # the record was taken on the fused path
    t = bench.hbm_kernel_table(prof, [(32, 1569)], 12, 384, 16, 6549.0)
    for k in ("ln_fwd", "ln_bwd", "colsum", "attn_bwd_fin", "embed_bwd"):  # unchanged rows reproduce the record
        assert abs(t[k]["GB/s"] - rec["hbm_kernels"][k]["GB/s"]) < 0.01 * rec["hbm_kernels"][k]["GB/s"], k
    assert "im2col" not in t and "tdl" not in t and t["tdl_followup"]["us_per_step"] > 0
    # 32 x 1568 tokens x (256 px x (4 B read + 2 B hi-patch) + 384 x 4 B token) = 154 MB over the kernel's time
    ef = t["embed_fused"]
    assert abs(ef["GB/s"] - 154.14e6 / (prof["embed_gemm"]["ms_per_step"] * 1e-3) / 1e9) < 1.0
    assert 0.0 < ef["frac_of_hbm_peak"] < 1.0
    # three-kernel path (ViT-B, So2Sat, uint8 input): the old rows, no fused row
    prof3 = dict(prof, im2col={"ms_per_step": 0.03, "launches_per_step": 1.0})
    t3 = bench.hbm_kernel_table(prof3, [(32, 1569)], 12, 384, 16, 6549.0)
    assert "im2col" in t3 and "tdl" in t3 and "embed_fused" not in t3 and "tdl_followup" not in t3
    # CHAMMI-style step: several sub-batches with different token counts add up
    t2 = bench.hbm_kernel_table(prof, [(22, 589), (21, 785), (21, 981)], 12, 384, 16, 6549.0)
    assert t2["ln_fwd"]["GB/s"] > 0
