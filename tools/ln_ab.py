"""GPU: LayerNorm forward / backward of one library build (DCV_LIB=<variant> selects libdcvit_<variant>.so) on inputs
larger than L2 (rotating buffers), CUDA-event timed.  Usage: [DCV_LIB=head] python tools/ln_ab.py"""
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from diverse_channel_vit_b200 import kernels as K  # noqa: E402

tag = os.environ.get("DCV_LIB", "shipped")
for M, D in ((50208, 384), (25120, 384), (6304, 384), (50208, 768)):
    nbuf = 6
    xs = [torch.randn(M, D, device="cuda") for _ in range(nbuf)]
    g, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    for i in range(nbuf):
        K.ln_fwd(xs[i], g, b)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(30):
            K.ln_fwd(xs[i % nbuf], g, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 30 * 1e3)
    y, mean, rstd = K.ln_fwd(xs[0], g, b)
    ref = torch.nn.functional.layer_norm(xs[0], (D,), g, b, 1e-6)
    err = ((y.float() - ref).norm() / ref.norm()).item()
    print(f"{tag:8s} ln_fwd M={M:6d} D={D}: {best:7.2f} us  {M * D * 6 / best / 1e3:7.1f} GB/s  (incl. the output allocation of the "
          f"test binding; rel err vs torch {err:.2e})", flush=True)
