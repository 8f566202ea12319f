"""One tiny training step + eval forward (used under compute-sanitizer)."""
import sys
sys.path.insert(0, ".")
import torch
from tests.util import O, build_cuda_model, cases, cuda_step, make_inputs
oc, mapper, chunk, has_head, B, wseed, iseed, xlam = cases()["tiny_chammi_hpa"]
weights = O.make_weights(oc, has_head, wseed)
x, y = make_inputs(oc, B, len(mapper[chunk]), oc.num_classes, iseed)
m = build_cuda_model(oc, mapper, weights)
out, extra, loss, grads = cuda_step(m, x.cuda(), y.cuda(), chunk, has_head, xlam, indices=[2, 0, 3])
torch.cuda.synchronize()
m.eval()
with torch.no_grad():
    o2 = m(x.cuda(), chunk)
torch.cuda.synchronize()
print("ok", float(loss), float(o2.abs().mean()))
