import sys
sys.path.insert(0, ".")
import ctypes, torch
from diverse_channel_vit_b200 import kernels as K, _lib
B, L, H = 8, 1569, 6
D = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * L, 3 * D, device="cuda", generator=g).bfloat16()
do = torch.randn(B * L, D, device="cuda", generator=g).bfloat16()
o, lse = K.attn_fwd(qkv, B, L, H)
K.attn_bwd(qkv, o, do, lse, B, L, H)
buf = torch.zeros(3 * 1024, dtype=torch.int64, device="cuda")
_lib.lib().dcv_debug_attn_timeline(ctypes.c_void_p(buf.data_ptr()))
K.attn_bwd(qkv, o, do, lse, B, L, H)
torch.cuda.synchronize()
_lib.lib().dcv_debug_attn_timeline(None)
t = buf.cpu().view(3, 128, 8)
t0 = int(t[0, 0, 0])
names = {0: ["mma:wait_p", "mma:got_p", "mma:wait_ds", "mma:got_ds", "mma:got_dqE", "mma:end"],
         1: ["wg0:top", "wg0:got_S", "wg0:p_arr", "wg0:drained", "wg0:got_dP", "wg0:ds_arr"],
         2: ["wg1:top", "wg1:got_S", "wg1:p_arr", "wg1:drained", "wg1:got_dP", "wg1:ds_arr"]}
for i in range(13):
    for role in range(3):
        print(f"it{i:2d} " + "  ".join(f"{names[role][p]}={int(t[role, i, p]) - t0:7d}" for p in range(6)))
