"""GPU diagnostic: per-parameter gradient error of the CUDA path against the fp32 CPU oracle at the BENCHED sizes
(C3 ViT-S/16 8ch 224^2, C4 ViT-S/8 18ch 32^2, C5 ViT-B/16 8ch 224^2), next to the error the reference's own bf16 path
(torch.autocast on the same GPU, the reference's use_amp switch) makes against the same oracle -- the bf16 noise floor
the tolerances in tests/test_model_gpu.py are set against.  Usage: python tools/grad_diag.py [c3|c4|c5|tiny ...]"""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from tests.util import O, build_cuda_model, cases, cuda_step, make_inputs, rel_l2  # noqa: E402


def full_cases():
    ch8 = [f"c{i}" for i in range(8)]
    c3 = O.OracleConfig(pretrained_model_name="small", img_size=224, patch_size=16, in_channel_names=ch8, num_classes=161,
                        proxy_loss_lambda=0.001, ortho_loss_v1_lambda=0.001, gamma_s=1.0, gamma_d=4.0,
                        reverse_pos_pairs=True, hcs_sampling_temp=1000.0)
    c4 = O.OracleConfig(pretrained_model_name="small", img_size=32, patch_size=8,
                        in_channel_names=[f"c{i}" for i in range(18)], num_classes=17, proxy_loss_lambda=0.001,
                        ortho_loss_v1_lambda=0.1, gamma_s=0.5, gamma_d=4.0, reverse_pos_pairs=True, hcs_sampling_temp=0.01)
    c5 = O.OracleConfig(pretrained_model_name="base", img_size=224, patch_size=16, in_channel_names=ch8, num_classes=161,
                        proxy_loss_lambda=0.0, ortho_loss_v1_lambda=0.0)
    return {
        "c3": (c3, {"train": list(range(8))}, "train", True, 2, 71, 72, 1.0, None),
        "c3s": (c3, {"train": list(range(8))}, "train", True, 2, 71, 72, 1.0, [5, 0, 3]),
        "c4": (c4, {"train": list(range(18))}, "train", True, 8, 73, 74, 1.0, None),
        "c5": (c5, {"train": list(range(8))}, "train", True, 2, 75, 76, 1.0, None),
    }


def block_forward_sdpa(x, p, prefix, heads):
    """oracle block with the attention done by torch SDPA (flash-style library kernel: O saved in bf16, backward
    recomputes P and uses delta = rowsum(dO * O)) instead of the reference's materialised softmax"""
    B, L, D = x.shape
    hd = D // heads
    u = F.layer_norm(x, (D,), p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], eps=1e-6)
    qkv = F.linear(u, p[prefix + "attn.qkv.weight"], p[prefix + "attn.qkv.bias"])
    qkv = qkv.reshape(B, L, 3, heads, hd).permute(2, 0, 3, 1, 4)
    o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(B, L, D)
    x = x + F.linear(o, p[prefix + "attn.proj.weight"], p[prefix + "attn.proj.bias"])
    v2 = F.layer_norm(x, (D,), p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], eps=1e-6)
    h = F.gelu(F.linear(v2, p[prefix + "mlp.fc1.weight"], p[prefix + "mlp.fc1.bias"]))
    return x + F.linear(h, p[prefix + "mlp.fc2.weight"], p[prefix + "mlp.fc2.bias"])


def run(name, case):
    oc, mapper, chunk, has_head, B, wseed, iseed, xlam, indices = case
    weights = O.make_weights(oc, has_head, wseed)
    x, y = make_inputs(oc, B, len(mapper[chunk]), oc.num_classes, iseed)
    t0 = time.time()
    o_loss, o, o_grads = O.loss_and_grads(x, y, weights, oc, mapper[chunk], has_head, indices=indices, extra_loss_lambda=xlam)
    t_or = time.time() - t0
    model = build_cuda_model(oc, mapper, weights)
    model._debug_keep_token_grad = True
    out, extra, loss, grads = cuda_step(model, x.cuda(), y.cuda(), chunk, has_head, xlam, indices=indices)
    torch.cuda.synchronize()
    # dLoss / d(tokens entering block 0): fp32 oracle on the CPU, autocast on the GPU, ours
    def token_grad(params, xin, yin, autocast):
        cap = {}
        keep_bf = O.block_forward

        def first(xx, pp, prefix, heads):
            if "x" not in cap:
                xx.retain_grad()
                cap["x"] = xx
            return keep_bf(xx, pp, prefix, heads)

        O.block_forward = first
        try:
            pr = {k: v.detach().to(xin.device).clone().requires_grad_(True) for k, v in params.items() if k != "adaptive_interface.0"}
            import contextlib
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
            with ctx:
                oo_ = O.forward(xin, pr, oc, mapper[chunk], training=True, has_head=has_head, indices=indices)
                l_ = O.train_loss(oo_, yin, pr, oc, has_head, xlam)
            l_.backward()
        finally:
            O.block_forward = keep_bf
        return cap["x"].grad.detach().double().cpu()

    G_or = token_grad(weights, x, y, False)
    G_amp = token_grad(weights, x.cuda(), y.cuda(), True)
    G_us = model._last_token_grad.double().cpu()

    def stats(G, tag):
        e = G - G_or
        tok = (e[:, 1:].norm(dim=-1) / G_or[:, 1:].norm(dim=-1).clamp_min(1e-30))
        print(f"   token grads {tag}: patch tokens rel-L2 {(e[:, 1:].norm() / G_or[:, 1:].norm()).item():.2e} (median per token {tok.median().item():.2e}), "
              f"CLS {(e[:, 0].norm() / G_or[:, 0].norm()).item():.2e}; sum over batch of patch tokens rel {(e[:, 1:].sum(0).norm() / G_or[:, 1:].sum(0).norm()).item():.2e}; "
              f"mean error vector / mean |token grad| = {(e[:, 1:].mean((0, 1)).norm() / G_or[:, 1:].norm(dim=-1).mean()).item():.2e}")

    stats(G_us, "ours")
    stats(G_amp, "amp ")
    # the positional-embedding gradient recomputed in fp64 from each set of token gradients
    from diverse_channel_vit_b200.dichavit import bicubic_pos_matrix
    n_pos = weights["feature_extractor.pos_embed"].shape[1] - 1
    Cs = len(indices) if indices is not None else len(mapper[chunk])
    Np = (G_or.shape[1] - 1) // Cs
    Mp = bicubic_pos_matrix(int(round(n_pos ** 0.5)), oc.img_size, oc.img_size, oc.patch_size).double() if Cs * Np != n_pos else torch.eye(n_pos).double()
    ref_pos = o_grads["feature_extractor.pos_embed"].double().cpu()[0, 1:]
    for tag, G in (("oracle G", G_or), ("our G", G_us), ("amp G", G_amp)):
        dp = Mp.t() @ G[:, 1:].reshape(G.shape[0], Cs, Np, -1).sum((0, 1))
        print(f"   pos patch rows from {tag:9s} in fp64: rel to oracle grad {((dp - ref_pos).norm() / ref_pos.norm()).item():.2e}")
    print(f"   pos patch rows, kernel output: rel to oracle grad {((grads['feature_extractor.pos_embed'].double().cpu()[0, 1:] - ref_pos).norm() / ref_pos.norm()).item():.2e}")
    print(f"   |G| patch tokens {G_or[:, 1:].norm().item():.3e}, |sum_b G| {G_or[:, 1:].sum(0).norm().item():.3e}, CLS {G_or[:, 0].norm().item():.3e}")
    # the reference's own bf16 path: oracle restatement under autocast on the GPU
    p = {k: v.detach().cuda().clone().requires_grad_(True) for k, v in weights.items() if k != "adaptive_interface.0"}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        oo = O.forward(x.cuda(), p, oc, mapper[chunk], training=True, has_head=has_head, indices=indices)
        l16 = O.train_loss(oo, y.cuda(), p, oc, has_head, xlam)
    l16.backward()
    # ... and the same with a flash-style library attention (what torch users get from SDPA under autocast)
    ps = {k: v.detach().cuda().clone().requires_grad_(True) for k, v in weights.items() if k != "adaptive_interface.0"}
    keep = O.block_forward
    O.block_forward = block_forward_sdpa
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            os_ = O.forward(x.cuda(), ps, oc, mapper[chunk], training=True, has_head=has_head, indices=indices)
            ls = O.train_loss(os_, y.cuda(), ps, oc, has_head, xlam)
        ls.backward()
    finally:
        O.block_forward = keep
    rows = []
    for k, g in o_grads.items():
        if g is None or g.abs().max() == 0:
            continue
        e_ours = rel_l2(grads[k], g)
        e_amp = rel_l2(p[k].grad, g) if p[k].grad is not None else float("nan")
        rows.append((e_ours, e_amp, k))
    rows.sort(reverse=True)
    ll = model.last_losses
    print(f"== {name}: oracle {t_or:.1f}s | out rel {rel_l2(out, o.out):.2e} (amp {rel_l2(oo.out.float(), o.out):.2e}) | "
          f"tdl rel {abs(ll['tdl'].item() - o.tdl.item()) / max(abs(o.tdl.item()), 1e-30):.2e} | "
          f"cdl rel {abs(ll['cdl'].item() - o.cdl.item()) / max(abs(o.cdl.item()), 1e-30):.2e} | "
          f"extra rel {abs(extra.item() - o.extra_loss.item()) / max(abs(o.extra_loss.item()), 1e-30):.2e}", flush=True)
    print(f"   grads: worst ours {rows[0][0]:.2e}, median ours {rows[len(rows) // 2][0]:.2e}, worst amp {max(r[1] for r in rows):.2e}, "
          f"median amp {sorted(r[1] for r in rows)[len(rows) // 2]:.2e}; > 1e-2 ours: {sum(r[0] > 1e-2 for r in rows)} of {len(rows)}")
    for e_ours, e_amp, k in rows[:12]:
        print(f"   {k:60s} ours {e_ours:.2e}   amp {e_amp:.2e}")
    k = "feature_extractor.pos_embed"
    g, go, ga = o_grads[k].double().cpu()[0], grads[k].double().cpu()[0], p[k].grad.double().cpu()[0]
    print(f"   pos_embed rows: CLS row ours {((go[0]-g[0]).norm()/g[0].norm()).item():.2e} amp {((ga[0]-g[0]).norm()/g[0].norm()).item():.2e}; "
          f"patch rows ours {((go[1:]-g[1:]).norm()/g[1:].norm()).item():.2e} amp {((ga[1:]-g[1:]).norm()/g[1:].norm()).item():.2e}; "
          f"|grad| CLS row {g[0].norm().item():.3e} patch rows {g[1:].norm().item():.3e}")
    gs = ps[k].grad.double().cpu()[0]
    print(f"   pos_embed with SDPA under autocast (library flash-style attention): all rows {((gs-g).norm()/g.norm()).item():.2e}, "
          f"CLS row {((gs[0]-g[0]).norm()/g[0].norm()).item():.2e}, patch rows {((gs[1:]-g[1:]).norm()/g[1:].norm()).item():.2e}")
    eo, ea = (go - g), (ga - g)
    print(f"   error direction: cos(err_ours, err_amp) = {(eo.flatten() @ ea.flatten() / (eo.norm() * ea.norm())).item():.3f}; "
          f"per-row mean of err_ours / row norm: {[round(float(eo[i].mean() / g[i].norm()), 4) for i in range(min(5, g.shape[0]))]}")
    del model, p
    torch.cuda.empty_cache()


def _modes():
    import os
    from diverse_channel_vit_b200 import _lib
    fm, bm = int(os.environ.get("DCV_FWD_MODE", "-1")), int(os.environ.get("DCV_BWD_MODE", "-1"))
    if fm >= 0 or bm >= 0:
        _lib.lib().dcv_debug_set_attn_mode(fm, bm)
        print(f"attention modes: fwd {fm} bwd {bm}")


if __name__ == "__main__":
    _modes()
    want = sys.argv[1:] or ["tiny", "c3", "c3s", "c4", "c5"]
    fc = full_cases()
    for w in want:
        if w == "tiny":
            for n, c in cases().items():
                run(n, (*c, None))
        else:
            run(w, fc[w])
