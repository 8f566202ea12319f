"""Per-iteration clock64 timeline of one attention-forward CTA (build with DCV_NVCC_EXTRA=-DDCV_ATTN_TIMELINE)."""
import sys
sys.path.insert(0, ".")
import ctypes, torch
from diverse_channel_vit_b200 import kernels as K, _lib
B, L, H = 32, 1569, 6
D = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * L, 3 * D, device="cuda", generator=g).bfloat16()
K.attn_fwd(qkv, B, L, H)
buf = torch.zeros(4096 + 2 * 1024, dtype=torch.int64, device="cuda")
_lib.lib().dcv_debug_attn_timeline(ctypes.c_void_p(buf.data_ptr()))
K.attn_fwd(qkv, B, L, H)
torch.cuda.synchronize()
_lib.lib().dcv_debug_attn_timeline(None)
t = buf.cpu()[4096:].view(2, 128, 8)
t0 = int(t[1, 0, 0])
names = {0: ["mma:S_next", "mma:PV"], 1: ["sm:top", "sm:got_S", "sm:S_ld", "sm:max", "sm:exp", "sm:arrive"]}
for i in range(13):
    for role in (0, 1):
        print(f"it{i:2d} " + "  ".join(f"{names[role][p]}={int(t[role, i, p]) - t0:7d}" for p in range(len(names[role]))))
