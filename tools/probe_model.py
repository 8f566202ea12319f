"""GPU probe: one training step of the drop-in module vs the CPU oracle on the golden cases."""
import sys, time
sys.path.insert(0, ".")
import torch
from tests.util import *

names = sys.argv[1:] or ["tiny_chammi_hpa", "tiny_jumpcp", "tiny_flags_rp0_sq1"]
torch.manual_seed(0)
for name in names:
    oc, mapper, chunk, has_head, B, wseed, iseed, xlam = cases()[name]
    weights = O.make_weights(oc, has_head, wseed)
    x, y = make_inputs(oc, B, len(mapper[chunk]), oc.num_classes, iseed)
    o_loss, o_out, o_grads = O.loss_and_grads(x, y, weights, oc, mapper[chunk], has_head, extra_loss_lambda=xlam)
    model = build_cuda_model(oc, mapper, weights)
    out, extra, loss, grads = cuda_step(model, x.cuda(), y.cuda(), chunk, has_head, xlam)
    torch.cuda.synchronize()
    ll = model.last_losses
    print(f"== {name}: out rel {rel_l2(out, o_out.out):.3e}  loss {loss.item():.6f} vs {o_loss.item():.6f}  "
          f"extra {extra.item():.6f} vs {o_out.extra_loss.item():.6f}  tdl {ll['tdl'].item():.6f} vs {o_out.tdl.item():.6f}  "
          f"cdl {ll['cdl'].item():.6f} vs {o_out.cdl.item():.6f}", flush=True)
    worst = []
    for k, g in o_grads.items():
        cg = grads.get(k)
        if g is None:
            continue
        if cg is None:
            print("   MISSING grad", k); continue
        worst.append((rel_l2(cg, g), k, g.norm().item()))
    worst.sort(reverse=True)
    for r, k, n in worst[:12]:
        print(f"   grad rel {r:.3e}  |g|={n:.3e}  {k}")
print("launches", __import__("diverse_channel_vit_b200._lib", fromlist=["x"]).launch_count())
