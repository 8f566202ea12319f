"""Soak test: many back-to-back launches of the pipelined kernels at assorted shapes (no host sync in between) to
flush out protocol races (mbarrier phase aliasing, buffer reuse); every shape's last result must match its first."""
import sys, time
sys.path.insert(0, ".")
import torch
from diverse_channel_vit_b200 import kernels as K

def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()

torch.manual_seed(0)
t0 = time.time()
shapes = [(1, 1, 2), (2, 17, 3), (3, 81, 3), (2, 128, 6), (2, 197, 6), (4, 289, 6), (2, 589, 6), (8, 785, 6), (4, 1569, 6), (32, 1569, 6), (64, 197, 6), (2, 1569, 12)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
worst = 0.0
for (B, L, H) in shapes:
    D = H * 64
    qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
    dy = torch.randn(B * L, D, device="cuda").bfloat16()
    w = (torch.randn(D, D, device="cuda") * 0.05).bfloat16()
    first = None
    for it in range(reps):
        o, lse = K.attn_fwd(qkv, B, L, H)
        d_o, delta = K.gemm_nn_delta(dy, w, o, B, L)
        dbias = torch.zeros(3 * D, device="cuda")
        dqkv = K.attn_bwd(qkv, o, d_o, lse, B, L, H, delta=delta, delta_ready=True, dbias=dbias)
        dq2 = K.attn_bwd(qkv, o, d_o, lse, B, L, H)
        if first is None:
            first = (o.clone(), dqkv.clone(), dbias.clone())
    torch.cuda.synchronize()
    e = max(rel(o, first[0]), rel(dqkv, first[1]), rel(dbias, first[2]), rel(dq2, first[1]))
    worst = max(worst, e)
    print(f"B{B} L{L} H{H}: {reps} x (fwd, dgrad+delta, bwd fused, bwd plain) ok, drift {e:.2e}", flush=True)
# GEMM epilogues, ragged M
for (M, N, Kd) in [(50, 192, 64), (777, 384, 384), (6304, 1152, 384), (6304, 1536, 384), (6304, 384, 1536), (50208, 384, 384)]:
    a = torch.randn(M, Kd, device="cuda").bfloat16(); b = (torch.randn(N, Kd, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda"); res = torch.randn(M, N, device="cuda")
    first = None
    for it in range(reps):
        h, g = K.gemm_nt(a, b, K.EPI_BIAS_GELU, bias=bias)
        r = K.gemm_nt(a, b, K.EPI_BIAS_RESID, bias=bias, resid=res)
        dg = K.gemm_nt(a, b, K.EPI_DGELU, aux=h)
        if first is None:
            first = (g.clone(), r.clone(), dg.clone())
    torch.cuda.synchronize()
    assert torch.equal(g, first[0]) and torch.equal(r, first[1]) and torch.equal(dg, first[2]), (M, N, Kd)
    print(f"gemm M{M} N{N} K{Kd}: {reps} x (gelu, resid, dgelu) bit-identical", flush=True)
assert worst < 5e-3, worst
print(f"SOAK OK in {time.time() - t0:.1f} s, worst drift {worst:.2e}")
