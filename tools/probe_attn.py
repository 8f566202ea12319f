"""GPU probe for the tcgen05 attention kernels."""
import sys, math
import torch
sys.path.insert(0, ".")
from diverse_channel_vit_b200 import kernels as K, _lib
torch.manual_seed(0)
dev = "cuda"

def stats(name, got, ref, tol=2e-2):
    got = got.float(); ref = ref.float()
    err = (got - ref).abs()
    rel = (err.norm() / ref.norm().clamp_min(1e-30)).item()
    print(f"{name}: rel_l2={rel:.3e} max_abs={err.max().item():.3e} ref_absmax={ref.abs().max().item():.3e} nan={torch.isnan(got).any().item()}", flush=True)
    if not rel < tol:
        print("   got", got.flatten()[:8].tolist()); print("   ref", ref.flatten()[:8].tolist())
    return rel

def ref_attn(qkv, B, L, H):
    D = H * 64
    q, k, v = qkv.float().reshape(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    p = s.softmax(-1)
    o = (p @ v).transpose(1, 2).reshape(B * L, D)
    lse2 = torch.logsumexp(s, -1) * math.log2(math.e)
    return o, lse2

def fwd_case(B, L, H, mag=1.0):
    qkv = (torch.randn(B * L, 3 * H * 64, device=dev) * mag).bfloat16()
    o, lse = K.attn_fwd(qkv, B, L, H)
    torch.cuda.synchronize()
    ro, rl = ref_attn(qkv, B, L, H)
    a = stats(f"fwd o   B{B} L{L} H{H}", o, ro)
    b = stats(f"fwd lse B{B} L{L} H{H}", lse[:, :, :L], rl, 1e-3)
    return a < 2e-2 and b < 1e-3

ok = True
for (B, L, H, mag) in [(1, 128, 1, 1.0), (2, 197, 3, 1.0), (2, 589, 6, 2.0), (1, 1569, 6, 1.0), (3, 81, 3, 3.0), (2, 17, 3, 1.0)]:
    try:
        ok &= fwd_case(B, L, H, mag)
    except Exception as ex:
        print("FWD FAILED", (B, L, H), repr(ex)); ok = False; break
print("ATTN FWD ok" if ok else "ATTN FWD BAD", flush=True)

def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
if ok:
    for (B, L, H) in [(32, 1569, 6), (32, 785, 6), (128, 289, 6)]:
        qkv = torch.randn(B * L, 3 * H * 64, device=dev).bfloat16()
        o = torch.empty(B * L, H * 64, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, L, device=dev)
        ms = bench(lambda: K.attn_fwd(qkv, B, L, H, o=o, lse2=lse))
        fl = 4.0 * B * H * L * L * 64
        q, k, v = qkv.reshape(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
        ms_t = bench(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
        print(f"attn fwd B{B} L{L}: {ms*1e3:.1f} us = {fl/ms/1e9:.1f} TFLOP/s (torch sdpa {ms_t*1e3:.1f} us = {fl/ms_t/1e9:.1f})", flush=True)

# ---------------- backward ----------------
def bwd_case(B, L, H, mag=1.0):
    D = H * 64
    qkv = (torch.randn(B * L, 3 * D, device=dev) * mag).bfloat16()
    do = (torch.randn(B * L, D, device=dev)).bfloat16()
    o, lse = K.attn_fwd(qkv, B, L, H)
    dqkv = K.attn_bwd(qkv, o, do, lse, B, L, H)
    torch.cuda.synchronize()
    x = qkv.float().requires_grad_(True)
    q, k, v = x.reshape(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    ro = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * L, D)
    ro.backward(do.float())
    g = x.grad
    ok = True
    for nm, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        ok &= stats(f"bwd {nm} B{B} L{L} H{H}", dqkv[:, sl], g[:, sl], 3e-2) < 3e-2
    return ok

okb = True
for (B, L, H, mag) in [(1, 128, 1, 1.0), (2, 197, 3, 1.0), (2, 589, 6, 2.0), (1, 1569, 6, 1.0), (3, 81, 3, 2.0), (2, 17, 3, 1.0)]:
    try:
        okb &= bwd_case(B, L, H, mag)
    except Exception as ex:
        print("BWD FAILED", (B, L, H), repr(ex)); okb = False; break
print("ATTN BWD ok" if okb else "ATTN BWD BAD", flush=True)
if okb:
    for (B, L, H) in [(32, 1569, 6), (32, 785, 6), (128, 289, 6)]:
        D = H * 64
        qkv = torch.randn(B * L, 3 * D, device=dev).bfloat16(); do = torch.randn(B * L, D, device=dev).bfloat16()
        o, lse = K.attn_fwd(qkv, B, L, H)
        dqkv = torch.empty_like(qkv); delta = K.delta_ws(B, H, L, dev); acc = torch.empty(B, H, L, 64, device=dev)
        ms = bench(lambda: K.attn_bwd(qkv, o, do, lse, B, L, H, dqkv=dqkv, delta=delta, dq_acc=acc))
        fl = 8.0 * B * H * L * L * 64
        print(f"attn bwd B{B} L{L}: {ms*1e3:.1f} us = {fl/ms/1e9:.1f} TFLOP/s (algorithmic 8*L^2*D)", flush=True)
