"""Minimal driver for ncu: a few attention fwd/bwd launches at the JUMP-CP shape."""
import sys
sys.path.insert(0, ".")
import torch
from diverse_channel_vit_b200 import kernels as K
B, L, H = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8, 1569, 6)))
D = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * L, 3 * D, device="cuda", generator=g).bfloat16()
do = torch.randn(B * L, D, device="cuda", generator=g).bfloat16()
for _ in range(3):
    o, lse = K.attn_fwd(qkv, B, L, H)
    dqkv = K.attn_bwd(qkv, o, do, lse, B, L, H)
torch.cuda.synchronize()
print("ok", float(dqkv.float().abs().mean()))
