"""GPU: every instruction-stream variant of the attention kernels (dcv_debug_set_attn_mode) -- parity against an fp32
torch reference at a ragged shape, then CUDA-event timings at the benched shapes next to torch SDPA (flash / cuDNN
backends) forward and backward on the same box.  Usage: python tools/attn_modes.py [--quick]"""
import math
import sys

import torch

sys.path.insert(0, ".")
from diverse_channel_vit_b200 import _lib, kernels as K  # noqa: E402

dev = "cuda"
lib = _lib.lib()
torch.manual_seed(0)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def reference(qkv, do, B, L, H):
    D = H * 64
    x = qkv.float().requires_grad_(True)
    q, k, v = x.reshape(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    o = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * L, D)
    o.backward(do.float())
    return o.detach(), torch.logsumexp(s.detach(), -1) * math.log2(math.e), x.grad


def bench(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


FAST = "--fast" in sys.argv  # shipped modes only, no SDPA legs
FWD_MODES = [1, 2, 3, 4] if FAST else [0, 1, 2, 3, 4]
BWD_MODES = [1] if FAST else [0, 1]

# ---- parity ----
for (B, L, H, mag) in [(2, 589, 3, 2.0), (1, 1569, 2, 1.0), (3, 81, 3, 3.0), (2, 197, 2, 2.0), (2, 393, 2, 2.0),
                       (2, 256, 2, 2.0), (2, 129, 2, 2.0), (3, 17, 1, 2.0)]:
    D = H * 64
    qkv = (torch.randn(B * L, 3 * D, device=dev) * mag).bfloat16()
    do = torch.randn(B * L, D, device=dev).bfloat16()
    ro, rl, rg = reference(qkv, do, B, L, H)
    for fm in FWD_MODES:
        lib.dcv_debug_set_attn_mode(fm, -1)
        o, lse = K.attn_fwd(qkv, B, L, H)
        torch.cuda.synchronize()
        print(f"parity fwd mode {fm} B{B} L{L} H{H}: o {rel(o, ro):.2e} lse {rel(lse[:, :, :L], rl):.2e}", flush=True)
    lib.dcv_debug_set_attn_mode(1, -1)
    o, lse = K.attn_fwd(qkv, B, L, H)
    for bm in BWD_MODES:
        lib.dcv_debug_set_attn_mode(-1, bm)
        g = K.attn_bwd(qkv, o, do, lse, B, L, H)
        torch.cuda.synchronize()
        print(f"parity bwd mode {bm} B{B} L{L} H{H}: dq {rel(g[:, :D], rg[:, :D]):.2e} dk {rel(g[:, D:2*D], rg[:, D:2*D]):.2e} "
              f"dv {rel(g[:, 2*D:], rg[:, 2*D:]):.2e}", flush=True)

# ---- timing ----
shapes = [(32, 1569, 6), (16, 1569, 12), (32, 785, 6), (128, 289, 6), (32, 393, 6), (32, 197, 6)]
if "--quick" in sys.argv:
    shapes = shapes[:1]
for (B, L, H) in shapes:
    D = H * 64
    qkv = torch.randn(B * L, 3 * D, device=dev).bfloat16()
    do = torch.randn(B * L, D, device=dev).bfloat16()
    o = torch.empty(B * L, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, K.lpad(L), device=dev)
    dqkv = torch.empty_like(qkv)
    delta = K.delta_ws(B, H, L, dev)
    acc = torch.empty(B, H, L, 64, device=dev)
    ff, fb = 4.0 * B * H * L * L * 64, 8.0 * B * H * L * L * 64
    for rep in range(2):
        for fm in FWD_MODES:
            lib.dcv_debug_set_attn_mode(fm, -1)
            us = bench(lambda: K.attn_fwd(qkv, B, L, H, o=o, lse2=lse))
            print(f"time fwd mode {fm} B{B} L{L} H{H} rep{rep}: {us:.1f} us = {ff / us / 1e6:.0f} TF", flush=True)
        lib.dcv_debug_set_attn_mode(1, -1)
        K.attn_fwd(qkv, B, L, H, o=o, lse2=lse)
        for bm in BWD_MODES:
            lib.dcv_debug_set_attn_mode(-1, bm)
            us = bench(lambda: K.attn_bwd(qkv, o, do, lse, B, L, H, dqkv=dqkv, delta=delta, dq_acc=acc))
            print(f"time bwd mode {bm} B{B} L{L} H{H} rep{rep}: {us:.1f} us (prep+main+finish) = {fb / us / 1e6:.0f} TF", flush=True)
    # torch SDPA on the same box (library kernels): forward and backward
    q, k, v = (t.contiguous().requires_grad_(True) for t in qkv.reshape(B, L, 3, H, 64).permute(2, 0, 3, 1, 4))
    from torch.nn.attention import SDPBackend, sdpa_kernel

    for name, be in (() if FAST else (("flash", SDPBackend.FLASH_ATTENTION), ("cudnn", SDPBackend.CUDNN_ATTENTION),
                                      ("efficient", SDPBackend.EFFICIENT_ATTENTION))):
        try:
            with sdpa_kernel(be):
                usf = bench(lambda: torch.nn.functional.scaled_dot_product_attention(q.detach(), k.detach(), v.detach()))
                out = torch.nn.functional.scaled_dot_product_attention(q, k, v)
                gout = torch.randn_like(out)
                usb = bench(lambda: torch.autograd.grad(out, (q, k, v), gout, retain_graph=True))
            print(f"time sdpa {name} B{B} L{L} H{H}: fwd {usf:.1f} us = {ff / usf / 1e6:.0f} TF, bwd {usb:.1f} us = {fb / usb / 1e6:.0f} TF",
                  flush=True)
        except Exception as ex:
            print(f"sdpa {name} unavailable: {repr(ex)[:120]}", flush=True)
lib.dcv_debug_set_attn_mode(1, 1)
