"""CPU: the oracle restatement (oracle/dichavit_oracle.py) against the golden vectors produced by the
unmodified reference module (oracle/make_golden.py, run in the build container)."""
import math
import random

import numpy as np
import pytest
import torch

from tests.util import CHAMMI_MAPPER, O, cases, load_golden, make_inputs

FAST = [n for n in cases() if n.startswith("tiny")]


@pytest.mark.parametrize("name", FAST + ["small_c1", "full_c3", "full_c4", "full_c5"])
def test_oracle_matches_reference_golden(name):
    oc, mapper, chunk, has_head, B, wseed, iseed, xlam = cases()[name]
    g = load_golden(name)
    weights = O.make_weights(oc, has_head, wseed)
    x, y = make_inputs(oc, B, len(mapper[chunk]), oc.num_classes, iseed)
    loss, o, grads = O.loss_and_grads(x, y, weights, oc, mapper[chunk], has_head, extra_loss_lambda=xlam)
    np.testing.assert_allclose(o.out.detach().numpy(), g["out"], rtol=1e-4, atol=2e-5)
    assert abs(o.extra_loss.item() - float(g["extra"])) <= 1e-5 * max(1.0, abs(float(g["extra"])))
    assert abs(loss.item() - float(g["loss"])) <= 2e-5 * max(1.0, abs(float(g["loss"])))
    assert abs(o.tdl.item() - float(g["tdl"])) <= 1e-6 and abs(o.cdl.item() - float(g["cdl"])) <= 1e-5
    for k in g.files:
        if not k.startswith("gstat:"):
            continue
        p = k[len("gstat:"):]
        gr = grads[p].detach().reshape(-1).double()
        norm, total = g[k]
        assert abs(gr.norm().item() - norm) <= 2e-4 * max(norm, 1e-6), p
        gen = torch.Generator().manual_seed(gr.numel())
        pick = torch.randint(0, gr.numel(), (16,), generator=gen)
        np.testing.assert_allclose(gr[pick].numpy(), g["gsamp:" + p], rtol=5e-3, atol=1e-4 * max(norm, 1e-6))
    # eval mode == train-mode output when nothing is sampled (dichavit.py:856-861)
    with torch.no_grad():
        oe = O.forward(x, weights, oc, mapper[chunk], training=False, has_head=has_head)
    np.testing.assert_allclose(oe.out.numpy(), g["out_eval"], rtol=1e-4, atol=2e-5)


def test_dcs_indices_match_reference():
    """DCS draws (python random x2 + torch.multinomial on the CPU generator) bit-exact with the reference."""
    g = load_golden("dcs_indices")
    names12 = [f"c{i}" for i in range(12)]
    oc = O.OracleConfig(pretrained_model_name="tiny", img_size=16, patch_size=8, in_channel_names=names12,
                        num_classes=14, enable_sample=True, hcs_sampling="lowest_cosine_prob", proxy_loss_lambda=0.1)
    weights = O.make_weights(oc, False, 51)
    ce_all = weights["feature_extractor.patch_embed.channel_embed.weight"]
    n = 0
    for key in g.files:
        tname, chunk = key.split(":")
        temp = {"t01": 0.1, "t1000": 1000.0, "t001": 0.01}[tname]
        ce = ce_all[torch.tensor(CHAMMI_MAPPER[chunk])]
        for row in g[key]:
            seed, c_new, anchor = int(row[0]), int(row[1]), int(row[2])
            idx = [int(v) for v in row[3:3 + c_new]]
            random.seed(seed)
            torch.manual_seed(seed + 2)
            got = O.dcs_select(ce, temp, "lowest_cosine_prob")
            assert got == (c_new, anchor, idx), (key, seed)
            assert anchor in idx and len(set(idx)) == c_new
            n += 1
    assert n == 360


def test_pos_interpolation_matches_reference():
    g = load_golden("pos_interp")
    for grid, img, P in ((14, 224, 16), (4, 32, 8), (2, 16, 8)):
        pos = torch.from_numpy(g[f"pos_{grid}"])
        out = O.interpolate_pos(pos, 3 * grid * grid, img, img, P, 3)[:, : 1 + grid * grid]
        assert torch.equal(out, torch.from_numpy(g[f"out_{grid}"]))
        # C' == 1 keeps the raw parameter (dichavit.py:529-530)
        assert O.interpolate_pos(pos, grid * grid, img, img, P, 1) is pos


def test_bicubic_matrix_reproduces_interpolate():
    """The explicit [N,N] matrix the CUDA path multiplies with equals F.interpolate."""
    from diverse_channel_vit_b200.dichavit import bicubic_pos_matrix

    for grid, img, P in ((14, 224, 16), (4, 32, 8)):
        pos = torch.randn(1, grid * grid + 1, 24, generator=torch.Generator().manual_seed(3))
        ref = O.interpolate_pos(pos, 2 * grid * grid, img, img, P, 2)[0, 1: 1 + grid * grid]
        W = bicubic_pos_matrix(grid, img, img, P)
        assert W.shape == (grid * grid, grid * grid)
        torch.testing.assert_close(W @ pos[0, 1:], ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("rp,sq", [(False, False), (False, True), (True, False), (True, True)])
def test_tdl_closed_form_equals_gram(rp, sq):
    """The per-channel-sum closed form used by the CUDA kernels == the reference's masked Gram sums."""
    torch.manual_seed(0)
    B, C, N, D = 3, 4, 9, 16
    y = torch.randn(B, C * N, D, dtype=torch.float64)
    labels = torch.arange(C).repeat_interleave(N)
    ref = O.tdl_loss(y, labels, 0.5, 4.0, rp, sq)
    f = torch.nn.functional.normalize(y, dim=-1).reshape(B, C, N, D)
    S = f.sum(2)
    a = (S ** 2).sum((1, 2))
    q = (f ** 2).sum((1, 2, 3))
    e = (S.sum(1) ** 2).sum(1)
    pos = (a - q) / (C * N * (N - 1) + 1e-6)
    neg = (e - a) / ((C * N) ** 2 - C * N * N + 1e-6)
    if sq:
        neg = neg ** 2
    if rp:
        loss = 0.5 * (pos ** 2 if sq else pos) + 4.0 * neg
    else:
        loss = 0.5 * (1 - pos) + 4.0 * neg
    assert abs(loss.mean().item() - ref.item()) < 1e-12


def test_tdl_single_channel_edge():
    """C' == 1: no negative pairs, neg = 0 / 1e-6 = 0 (loss_fn.py:44-48)."""
    y = torch.randn(2, 5, 8)
    v = O.tdl_loss(y, torch.zeros(5, dtype=torch.long), 1.0, 4.0, True, False)
    f = torch.nn.functional.normalize(y, dim=-1)
    pos = ((f.sum(1) ** 2).sum(-1) - 5) / (5 * 4 + 1e-6)
    assert abs(v.item() - pos.mean().item()) < 1e-5


def test_leave_one_out_matches_reference():
    """Eval-time channel-token synthesis for channels unseen in training (dichavit.py:219-374)."""
    from oracle.make_golden import LOO_MAPPER, LOO_MODES

    g = load_golden("leave_one_out")
    oc = O.OracleConfig(pretrained_model_name="tiny", img_size=32, patch_size=8,
                        in_channel_names=[f"c{i}" for i in range(7)], num_classes=6, proxy_loss_lambda=0.1,
                        ortho_loss_v1_lambda=0.5)
    weights = O.make_weights(oc, True, 61)
    x, _ = make_inputs(oc, 3, 4, oc.num_classes, 62)
    w = weights["feature_extractor.patch_embed.channel_embed.weight"]
    for mode in LOO_MODES:
        ce = O.leave_one_out_channel_tokens(w, LOO_MAPPER, "test", "train", mode)
        with torch.no_grad():
            oo = O.forward(x, weights, oc, LOO_MAPPER["test"], training=False, has_head=True, channel_embed_override=ce)
        np.testing.assert_allclose(oo.out.numpy(), g[mode], rtol=1e-4, atol=2e-5)
    # host-side mirror in the drop-in module produces the same token matrix
    from diverse_channel_vit_b200.dichavit import dichavit
    from tests.util import ref_cfg

    m = dichavit(ref_cfg(oc), mapper=LOO_MAPPER)
    m.load_state_dict({k: weights[k] for k in m.state_dict()})
    pe = m.feature_extractor.patch_embed
    for mode in LOO_MODES:
        got = pe.leave_one_out_tokens("test", "train", mode)
        assert torch.equal(got, O.leave_one_out_channel_tokens(w, LOO_MAPPER, "test", "train", mode))
    assert pe.leave_one_out_tokens("train", "train", None) is None
    with pytest.raises(ValueError):
        pe.leave_one_out_tokens("test", "train", "dynamic_input_corr_1")
    with pytest.raises(ValueError):
        pe.leave_one_out_tokens("test", "train", "nonsense")


def test_bicubic_backward_is_what_autograd_applies_not_the_adjoint():
    """reference models/dichavit.py:531-549: the positional grid goes through F.interpolate(scale_factor=(w0+0.1)/grid,
    mode="bicubic").  ATen's backward of that call is NOT the transpose of its forward: it derives the sampling scale
    from the tensor sizes.  At every training shape (w // P == grid) the forward matrix is off the identity by ~2-3 %
    while the backward IS the identity -- the drop-in has to reproduce that, not the exact adjoint."""
    import torch.nn.functional as F

    from diverse_channel_vit_b200.dichavit import bicubic_pos_backward_matrix, bicubic_pos_matrix

    for grid, w, patch in ((4, 32, 8), (14, 224, 16), (2, 32, 16)):
        fwd = bicubic_pos_matrix(grid, w, w, patch)
        assert (fwd - torch.eye(grid * grid)).abs().max() > 5e-3          # the forward resamples ...
        assert bicubic_pos_backward_matrix(grid, w, w, patch) is None      # ... the backward does not
        x = torch.randn(1, 3, grid, grid, dtype=torch.float64, requires_grad=True)
        y = F.interpolate(x, scale_factor=((w // patch + 0.1) / grid,) * 2, mode="bicubic")
        gy = torch.randn_like(y)
        y.backward(gy)
        assert torch.equal(x.grad, gy)
    # another image size (evaluation with a gradient): whatever autograd applies is what the matrix holds
    grid, w, patch = 4, 64, 8
    mb = bicubic_pos_backward_matrix(grid, w, w, patch)
    assert mb is not None and mb.shape == (64, 16)
    x = torch.randn(1, 5, grid, grid, requires_grad=True)
    y = F.interpolate(x, scale_factor=((w // patch + 0.1) / grid,) * 2, mode="bicubic")
    gy = torch.randn_like(y)
    y.backward(gy)
    want = x.grad[0].reshape(5, 16).t()                                    # [N_in, D]
    got = mb.t() @ gy[0].reshape(5, 64).t()
    assert torch.allclose(got, want, atol=1e-5)
