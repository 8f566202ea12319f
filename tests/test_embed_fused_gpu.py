"""GPU: the fused patch embedding (csrc/embed_fused.cu: TMA from the fp32 image -> bf16 hi/lo split in shared memory ->
tcgen05 -> tokens + TDL partial sums in the epilogue) against the three-kernel path it replaces (gather / im2col,
EPI_EMBED GEMM, tdl_sum) on the same module, weights and inputs, at the benched JUMP-CP shape (8 x 224 x 224, P = 16,
D = 384) with all channels and with a sampled, unsorted channel draw.  Reference: models/dichavit.py:210, :377-389,
:409-411.  (Both paths are compared with the fp32 oracle in tests/test_fullsize_gpu.py / test_model_gpu.py.)"""
import pytest
import torch

from diverse_channel_vit_b200 import _lib
from tests.util import O, build_cuda_model, cases, cuda_step, make_inputs, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("indices", [None, [5, 0, 3], [2]])
def test_fused_patch_embedding_equals_three_kernel_path(indices, mode):
    """mode 1: one tile per CTA; mode 2: the persistent, cross-tile pipelined kernel."""
    oc, mapper, chunk, has_head, B, wseed, iseed, xlam = cases()["full_c3"]
    weights = O.make_weights(oc, has_head, wseed)
    x, y = make_inputs(oc, B, len(mapper[chunk]), oc.num_classes, iseed)
    res = {}
    lib = _lib.lib()
    try:
        for fused in (mode, 0):
            lib.dcv_debug_set_embed_fused(fused)
            model = build_cuda_model(oc, mapper, weights)
            out, extra, loss, grads = cuda_step(model, x.cuda(), y.cuda(), chunk, has_head, xlam, indices=indices)
            torch.cuda.synchronize()
            res[fused] = (out.detach().clone(), extra.detach().clone(), {k: v.item() for k, v in model.last_losses.items()},
                          {k: g.detach().clone() for k, g in grads.items() if g is not None})
    finally:
        lib.dcv_debug_set_embed_fused(-1)
    (o1, e1, l1, g1), (o0, e0, l0, g0) = res[mode], res[0]
    # the losses are fp32 reductions of the fp32 projection: only the summation order differs
    assert abs(l1["tdl"] - l0["tdl"]) <= 1e-5 * abs(l0["tdl"]) + 1e-9
    assert abs(l1["cdl"] - l0["cdl"]) <= 1e-6 * abs(l0["cdl"]) + 1e-9
    assert abs(e1.item() - e0.item()) <= 1e-5 * abs(e0.item()) + 1e-9
    # the tokens agree to fp32 rounding; 12 bf16 blocks amplify last-bit differences to the bf16 noise floor
    assert rel_l2(o1, o0) < 5e-3
    for k in ("feature_extractor.patch_embed.proj.weight", "feature_extractor.patch_embed.proj.bias",
              "feature_extractor.patch_embed.channel_embed.weight", "feature_extractor.pos_embed"):
        if g0[k].abs().max() > 0:
            assert rel_l2(g1[k], g0[k]) < 1e-2, k
