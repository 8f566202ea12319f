"""GPU: time the attention kernels of one library build (DCV_LIB=<variant> selects libdcvit_<variant>.so), for A/B
comparisons of two builds inside ONE gpurun call (box-to-box variation is ~15 %).
Usage: [DCV_LIB=head] python tools/attn_ab.py [bwd_mode ...]"""
import os
import sys

import torch

sys.path.insert(0, ".")
from diverse_channel_vit_b200 import _lib, kernels as K  # noqa: E402

lib = _lib.lib()
tag = os.environ.get("DCV_LIB", "shipped")


def bench(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


modes = [int(a) for a in sys.argv[1:]] or [-1]
for (B, L, H) in [(32, 1569, 6), (32, 785, 6), (32, 197, 6)]:
    D = H * 64
    torch.manual_seed(0)
    qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
    do = torch.randn(B * L, D, device="cuda").bfloat16()
    o = torch.empty(B * L, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, K.lpad(L), device="cuda")
    dqkv = torch.empty_like(qkv)
    delta = K.delta_ws(B, H, L, "cuda")
    acc = torch.empty(B, H, L, 64, device="cuda")
    for rep in range(2):
        us = bench(lambda: K.attn_fwd(qkv, B, L, H, o=o, lse2=lse))
        print(f"[{tag}] fwd B{B} L{L} rep{rep}: {us:.1f} us", flush=True)
        for m in modes:
            if m >= 0:
                lib.dcv_debug_set_attn_mode(-1, m)
            us = bench(lambda: K.attn_bwd(qkv, o, do, lse, B, L, H, dqkv=dqkv, delta=delta, dq_acc=acc))
            print(f"[{tag}] bwd mode {m} B{B} L{L} rep{rep}: {us:.1f} us (prep+main+finish)", flush=True)
