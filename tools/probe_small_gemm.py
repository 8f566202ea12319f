import sys
sys.path.insert(0, ".")
import torch
from diverse_channel_vit_b200 import kernels as K
dev = "cuda"
def bench(name, fn, iters=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for M in (6304, 12576, 25120, 50208):
    a = torch.randn(M, 384, device=dev).bfloat16(); w = torch.randn(1152, 384, device=dev).bfloat16()
    bias = torch.randn(1152, device=dev); out = torch.empty(M, 1152, device=dev, dtype=torch.bfloat16)
    t1 = bench("ours", lambda: K.gemm_nt(a, w, K.EPI_BIAS, bias=bias, out=out))
    t2 = bench("torch", lambda: torch.addmm(bias.bfloat16(), a, w.t(), out=out))
    fl = 2.0 * M * 1152 * 384
    x = torch.randn(M, 384, device=dev); g = torch.ones(384, device=dev); b = torch.zeros(384, device=dev)
    t3 = bench("ln", lambda: K.ln_fwd(x, g, b))
    print(f"M={M:6d} qkv gemm: ours {t1:6.1f} us ({fl/t1/1e6:6.1f} TF)  torch {t2:6.1f} us ({fl/t2/1e6:6.1f} TF)   ln_fwd {t3:5.1f} us", flush=True)
