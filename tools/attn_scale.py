import sys
sys.path.insert(0, ".")
import torch
from diverse_channel_vit_b200 import kernels as K
H = 6
for (B, L) in [(1, 1569), (4, 1569), (8, 1569), (16, 1569), (32, 785), (32, 1569)]:
    D = H * 64
    qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16(); do = torch.randn(B * L, D, device="cuda").bfloat16()
    o, lse = K.attn_fwd(qkv, B, L, H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): K.attn_bwd(qkv, o, do, lse, B, L, H)
    e1.record(); torch.cuda.synchronize()
    print(f"B{B} L{L}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us", flush=True)
