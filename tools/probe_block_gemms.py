"""Times every GEMM flavour of one transformer block at the JUMP-CP shape (M = 32*1569), L2-cold-ish
(rotating buffers larger than L2)."""
import sys
sys.path.insert(0, ".")
import torch
from diverse_channel_vit_b200 import kernels as K, _lib
import os
if os.environ.get('DCV_CM'): _lib.lib().dcv_debug_set_nt_cluster(int(os.environ['DCV_CM']))
M, D, F = 50208, 384, 1536
dev = "cuda"
def bf(*s): return (torch.randn(*s, device=dev) * 0.5).bfloat16()
NB = 3  # rotate buffers so inputs do not sit in L2
acts = {k: [bf(M, n) for _ in range(NB)] for k, n in (("d", D), ("f", F), ("q", 3 * D))}
res = [torch.randn(M, D, device=dev) for _ in range(NB)]
W = {"qkv": bf(3 * D, D), "proj": bf(D, D), "fc1": bf(F, D), "fc2": bf(D, F)}
bias = {k: torch.randn(v.shape[0], device=dev) for k, v in W.items()}
outs = {"d16": [torch.empty(M, D, device=dev, dtype=torch.bfloat16) for _ in range(NB)],
        "f16": [torch.empty(M, F, device=dev, dtype=torch.bfloat16) for _ in range(NB)],
        "f16b": [torch.empty(M, F, device=dev, dtype=torch.bfloat16) for _ in range(NB)],
        "q16": [torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16) for _ in range(NB)],
        "d32": [torch.empty(M, D, device=dev) for _ in range(NB)]}
def bench(name, fn, flops, iters=12):
    for i in range(3): fn(i % NB)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i % NB)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{name:34s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
fl = lambda n, k: 2.0 * M * n * k
bench("nt qkv   bias      N=1152 K=384", lambda i: K.gemm_nt(acts["d"][i], W["qkv"], K.EPI_BIAS, bias=bias["qkv"], out=outs["q16"][i]), fl(3 * D, D))
bench("nt proj  resid     N=384  K=384", lambda i: K.gemm_nt(acts["d"][i], W["proj"], K.EPI_BIAS_RESID, bias=bias["proj"], out=outs["d32"][i], resid=res[i]), fl(D, D))
bench("nt fc1   gelu      N=1536 K=384", lambda i: K.gemm_nt(acts["d"][i], W["fc1"], K.EPI_BIAS_GELU, bias=bias["fc1"], out=outs["f16"][i], out2=outs["f16b"][i]), fl(F, D))
bench("nt fc2   resid     N=384  K=1536", lambda i: K.gemm_nt(acts["f"][i], W["fc2"], K.EPI_BIAS_RESID, bias=bias["fc2"], out=outs["d32"][i], resid=res[i]), fl(D, F))
bench("nn dfc2  dgelu     N=1536 K=384", lambda i: K.gemm_nn(acts["d"][i], W["fc2"], K.EPI_DGELU, out=outs["f16"][i], aux=acts["f"][i]), fl(F, D))
bench("nn dfc1  plain     N=384  K=1536", lambda i: K.gemm_nn(acts["f"][i], W["fc1"], out=outs["d16"][i]), fl(D, F))
bench("nn dproj plain     N=384  K=384", lambda i: K.gemm_nn(acts["d"][i], W["proj"], out=outs["d16"][i]), fl(D, D))
bench("nn dqkv  plain     N=384  K=1152", lambda i: K.gemm_nn(acts["q"][i], W["qkv"], out=outs["d16"][i]), fl(D, 3 * D))
gW = {k: torch.zeros_like(v, dtype=torch.float32) for k, v in W.items()}
bench("tn wfc2  [384,1536]", lambda i: K.gemm_tn(acts["d"][i], acts["f"][i], out=gW["fc2"]), fl(D, F))
bench("tn wfc1  [1536,384]", lambda i: K.gemm_tn(acts["f"][i], acts["d"][i], out=gW["fc1"]), fl(D, F))
bench("tn wproj [384,384]", lambda i: K.gemm_tn(acts["d"][i], acts["d"][(i + 1) % NB], out=gW["proj"]), fl(D, D))
bench("tn wqkv  [1152,384]", lambda i: K.gemm_tn(acts["q"][i], acts["d"][i], out=gW["qkv"]), fl(D, 3 * D))
# row kernels
x = [torch.randn(M, D, device=dev) for _ in range(NB)]
g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
y, mean, rstd = K.ln_fwd(x[0], g, b)
bench("ln_fwd", lambda i: K.ln_fwd(x[i], g, b), M * D * 6 * 1e3)  # "TFLOP/s" column = GB/s here
dg, db_, ds = (torch.zeros(D, device=dev) for _ in range(3))
bench("ln_bwd", lambda i: K.ln_bwd(acts["d"][i], x[i], mean, rstd, g, res[i], dg, db_, ds), M * D * 16 * 1e3)
cs = torch.zeros(F, device=dev)
bench("colsum [M,1536]", lambda i: K.colsum_bf16(acts["f"][i], cs), M * F * 2 * 1e3)
