import sys
sys.path.insert(0, ".")
import ctypes, os, torch
from diverse_channel_vit_b200 import kernels as K, _lib
B, L, H = 8, 1569, 6
_lib.lib().dcv_debug_set_attn_mode(-1, int(os.environ.get("BWD_MODE", "1")))
if hasattr(_lib.lib(), "dcv_debug_set_attn_ablate"):  # ablation builds only
    _lib.lib().dcv_debug_set_attn_ablate(int(os.environ.get("ABLATE", "0"), 0))
D = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * L, 3 * D, device="cuda", generator=g).bfloat16()
do = torch.randn(B * L, D, device="cuda", generator=g).bfloat16()
o, lse = K.attn_fwd(qkv, B, L, H)
K.attn_bwd(qkv, o, do, lse, B, L, H)
buf = torch.zeros(4 * 1024, dtype=torch.int64, device="cuda")
_lib.lib().dcv_debug_attn_timeline(ctypes.c_void_p(buf.data_ptr()))
K.attn_bwd(qkv, o, do, lse, B, L, H)
torch.cuda.synchronize()
_lib.lib().dcv_debug_attn_timeline(None)
t = buf.cpu().view(4, 128, 8)
t0 = int(t[0, 0, 0])
names = {0: ["front:S_next", "front:dP", "front:dp_consumed", "front:do_full"],
         3: ["back:phase", "back:dKdV", "back:dQ", "back:S", "back:dV", "back:dq_empty"],
         1: ["wg0:top", "wg0:S_ld", "wg0:got_dP", "wg0:chunks", "wg0:arrive", "wg0:pfree", "wg0:sfull", "wg0:tmemw"],
         2: ["drain:dq_full", "drain:ld0", "drain:smem_free", "drain:dq_empty", "drain:reduce"]}
t0 = int(t[1, 0, 0])
for i in range(14):
    for role in (0, 3, 1, 2):
        print(f"it{i:2d} " + "  ".join(f"{names[role][p]}={int(t[role, i, p]) - t0:7d}" for p in range(len(names[role]))))
