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


def run(name, case):
    oc, mapper, chunk, has_head, B, wseed, iseed, xlam, indices = case
    weights = O.make_weights(oc, has_head, wseed)
    x, y = make_inputs(oc, B, len(mapper[chunk]), oc.num_classes, iseed)
    t0 = time.time()
    o_loss, o, o_grads = O.loss_and_grads(x, y, weights, oc, mapper[chunk], has_head, indices=indices, extra_loss_lambda=xlam)
    t_or = time.time() - t0
    model = build_cuda_model(oc, mapper, weights)
    out, extra, loss, grads = cuda_step(model, x.cuda(), y.cuda(), chunk, has_head, xlam, indices=indices)
    torch.cuda.synchronize()
    # the reference's own bf16 path: oracle restatement under autocast on the GPU
    p = {k: v.detach().cuda().clone().requires_grad_(True) for k, v in weights.items() if k != "adaptive_interface.0"}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        oo = O.forward(x.cuda(), p, oc, mapper[chunk], training=True, has_head=has_head, indices=indices)
        l16 = O.train_loss(oo, y.cuda(), p, oc, has_head, xlam)
    l16.backward()
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
    del model, p
    torch.cuda.empty_cache()


if __name__ == "__main__":
    want = sys.argv[1:] or ["tiny", "c3", "c3s", "c4", "c5"]
    fc = full_cases()
    for w in want:
        if w == "tiny":
            for n, c in cases().items():
                run(n, (*c, None))
        else:
            run(w, fc[w])
