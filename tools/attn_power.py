"""GPU diagnostic: SM clock and board power while one attention kernel runs back to back for ~2 s each -- tells a
pipeline bound from a power-cap bound (clock64-based timelines count SM cycles; wall-clock microseconds do not).
Usage: python tools/attn_power.py [bwd_mode]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from bench import ClockSampler  # noqa: E402
from diverse_channel_vit_b200 import _lib, kernels as K  # noqa: E402

lib = _lib.lib()
B, L, H = 32, 1569, 6
D = H * 64
torch.manual_seed(0)
qkv = torch.randn(B * L, 3 * D, device="cuda").bfloat16()
do = torch.randn(B * L, D, device="cuda").bfloat16()
o, lse = K.attn_fwd(qkv, B, L, H)
dqkv = torch.empty_like(qkv)
delta = K.delta_ws(B, H, L, "cuda")
acc = torch.empty(B, H, L, 64, device="cuda")
if len(sys.argv) > 1:
    lib.dcv_debug_set_attn_mode(-1, int(sys.argv[1]))
q, k, v = (t.contiguous() for t in qkv.reshape(B, L, 3, H, 64).permute(2, 0, 3, 1, 4))
big_a = torch.randn(8192, 8192, device="cuda").bfloat16()
big_b = torch.randn(8192, 8192, device="cuda").bfloat16()


def run(name, f, seconds=2.0, flops=None):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    cs = ClockSampler(0)
    time.sleep(0.3)
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(20):
            f()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    us = e0.elapsed_time(e1) / n * 1e3
    c = cs.stop(t0 + 0.5, t1)
    tf = f" {flops / us / 1e6:7.0f} TF" if flops else ""
    print(f"{name:28s} {us:8.1f} us{tf}  sm {c['sm_mhz']} / {c['sm_max_mhz']} MHz  power max {c['power_w_max']} W  {c['reasons']}", flush=True)


from torch.nn.attention import SDPBackend, sdpa_kernel  # noqa: E402

ff, fb = 4.0 * B * H * L * L * 64, 8.0 * B * H * L * L * 64
run("attn_fwd (ours)", lambda: K.attn_fwd(qkv, B, L, H, o=o, lse2=lse), flops=ff)
run("attn_bwd (ours, 3 kernels)", lambda: K.attn_bwd(qkv, o, do, lse, B, L, H, dqkv=dqkv, delta=delta, dq_acc=acc), flops=fb)
with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
    run("sdpa cudnn fwd", lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), flops=ff)
    qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
    out = torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
    g = torch.randn_like(out)
    run("sdpa cudnn bwd", lambda: torch.autograd.grad(out, (qq, kk, vv), g, retain_graph=True), flops=fb)
run("cublas 8192^3 bf16", lambda: torch.mm(big_a, big_b), flops=2.0 * 8192 ** 3)
