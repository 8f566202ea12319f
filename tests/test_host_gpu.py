"""GPU: host-side behaviour of the drop-in module around the kernels -- optimiser semantics against torch.optim.AdamW
(the reference's optimizers.py:20-21), state carried by the autograd node, copies of the module, gradient accumulation,
the learnable temperature, device-resident schedules."""
import copy
import pickle

import pytest
import torch
import torch.nn.functional as F

from tests.util import CHAMMI_MAPPER, O, build_cuda_model, cases, cuda_step, make_inputs, ref_cfg, rel_l2
from diverse_channel_vit_b200.dichavit import dichavit
from diverse_channel_vit_b200.optim import CosineLRSchedule, CosineWDSchedule, FusedAdamW
from diverse_channel_vit_b200.trainer_glue import training_loss

pytestmark = pytest.mark.gpu


def _pair(name, **cfg_over):
    oc, mapper, chunk, has_head, B, wseed, iseed, xlam = cases()[name]
    weights = O.make_weights(oc, has_head, wseed)
    x, y = make_inputs(oc, B, len(mapper[chunk]), oc.num_classes, iseed)

    def make():
        cfg = ref_cfg(oc)
        cfg.update(cfg_over)
        m = dichavit(cfg, mapper=mapper)
        m.load_state_dict({k: weights[k].clone() for k in m.state_dict() if k in weights}, strict=False)
        return m.cuda()

    return make, x.cuda(), y.cuda(), chunk, has_head, xlam


@pytest.mark.parametrize("direct", [False, True])
def test_fused_adamw_trains_trainer_owned_proxies(direct):
    """CHAMMI (no classifier head): `proxies` gets its gradient from the trainer's proxy loss through torch autograd,
    not from the kernels -- its .grad is a separate tensor.  FusedAdamW must pick it up (round-1 bug: it only decayed)."""
    make, x, y, chunk, has_head, xlam = _pair("tiny_chammi_hpa")
    assert not has_head
    ma, mb = make(), make()
    ma.direct_grad = direct
    oa = FusedAdamW(ma, lr=1e-3, weight_decay=0.05)
    ob = torch.optim.AdamW(mb.parameters(), lr=1e-3, weight_decay=0.05)
    p0 = ma.proxies.detach().clone()
    for _ in range(3):
        oa.zero_grad()
        cuda_step(ma, x, y, chunk, has_head, xlam) if not direct else None
        if direct:
            ma.train()
            out, extra = ma(x, chunk)
            training_loss(ma, out, extra, y, has_head, xlam).backward()
        assert ma.proxies.grad is not None and ma.proxies.grad.abs().max() > 0
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            pb.grad = None if pa.grad is None else pa.grad.detach().clone()
        oa.step()
        ob.step()
    assert rel_l2(ma.proxies, p0) > 1e-4  # it moved ...
    for (ka, pa), (kb, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        if pb.grad is not None:
            assert rel_l2(pa, pb) < 1e-5, (ka, rel_l2(pa, pb))  # ... exactly like torch's AdamW moves it


def test_fused_adamw_leaves_frozen_and_unused_parameters_alone():
    """torch / timm AdamW skip parameters without a gradient: no update and no weight decay.  freeze_channel_emb=True
    freezes the channel tokens (reference dichavit.py:88-89); `proxies` is unused on a model with a classifier head."""
    make, x, y, chunk, has_head, xlam = _pair("tiny_jumpcp", freeze_channel_emb=True)
    assert has_head
    for direct in (False, True):
        m = make()
        m.direct_grad = direct
        ce = m.feature_extractor.patch_embed.channel_embed.weight
        assert not ce.requires_grad
        ce0, px0 = ce.detach().clone(), m.proxies.detach().clone()
        w0 = m.feature_extractor.blocks[0].mlp.fc1.weight.detach().clone()
        opt = FusedAdamW(m, lr=1e-2, weight_decay=0.5, clip_grad_norm=0.1)
        for _ in range(2):
            opt.zero_grad()
            m.train()
            out, extra = m(x, chunk)
            training_loss(m, out, extra, y, has_head, xlam).backward()
            assert ce.grad is None and m.proxies.grad is None
            opt.step()
        assert torch.equal(ce, ce0) and torch.equal(m.proxies, px0)
        assert not torch.equal(m.feature_extractor.blocks[0].mlp.fc1.weight, w0)
    # the clipping norm ignores them too: compare one clipped step with torch on the trainable parameters
    ma, mb = make(), make()
    oa = FusedAdamW(ma, lr=1e-3, weight_decay=0.05, clip_grad_norm=0.05)
    ob = torch.optim.AdamW([p for p in mb.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.05)
    cuda_step(ma, x, y, chunk, has_head, xlam)
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        pb.grad = None if pa.grad is None else pa.grad.detach().clone()
    oa.step()
    torch.nn.utils.clip_grad_norm_([p for p in mb.parameters() if p.grad is not None], 0.05)
    ob.step()
    for (ka, pa), (kb, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert rel_l2(pa, pb) < 1e-5, ka


def test_backward_uses_the_state_of_its_own_forward():
    """Two forwards before one backward, and a train() -> eval() toggle in between, must not change the gradients:
    loss weights, input dtype, pixel statistics travel with the autograd node (round-1: read from the module)."""
    make, x, y, chunk, has_head, xlam = _pair("tiny_jumpcp")
    ref = make()
    _, _, _, g_ref = cuda_step(ref, x, y, chunk, has_head, xlam)
    m = make()
    m.train()
    out, extra = m(x, chunk)
    loss = training_loss(m, out, extra, y, has_head, xlam)
    xu = torch.randint(0, 256, x.shape, dtype=torch.uint8, device="cuda")
    out2, extra2 = m(xu, chunk, pixel_mean=torch.full((x.shape[1],), 120.0), pixel_std=torch.full((x.shape[1],), 50.0))
    m.eval()            # lambda_tdl / lambda_cdl are 0 in eval mode: backward must still use the training values
    with torch.no_grad():
        m(x, chunk)     # and an inference forward in between must not disturb the saved state
    loss.backward()
    for k, p in m.named_parameters():
        if g_ref[k] is None:
            continue
        assert rel_l2(p.grad, g_ref[k]) < 1e-3, k  # same kernels; atomics reorder the last bits
    ce = m.feature_extractor.patch_embed.channel_embed.weight.grad
    assert ce.abs().max() > 0


def test_module_can_be_deep_copied_and_pickled_after_a_forward():
    """AveragedModel (SWA, trainer.py:243) deep-copies the model; the engine caches hold raw-pointer ctypes structs."""
    make, x, y, chunk, has_head, xlam = _pair("tiny_jumpcp")
    m = make()
    out, _, _, _ = cuda_step(m, x, y, chunk, has_head, xlam)
    c = copy.deepcopy(m)
    blob = pickle.dumps(m)
    u = pickle.loads(blob)
    for other in (c, u):
        other.train()
        o2, _ = other(x, chunk)
        assert rel_l2(o2, out) < 1e-6
        with torch.no_grad():  # the copy owns its parameters: writing them must not touch the original
            other.classifer_head.weight.zero_()
        assert m.classifer_head.weight.abs().max() > 0
    avg = torch.optim.swa_utils.AveragedModel(m)
    avg.update_parameters(m)


def test_direct_grad_accumulates_in_place_over_several_backwards():
    """CHAMMI: three chunks, one optimiser step (trainer.py:846-931).  In direct_grad mode the second and third
    backward accumulate on top of the flat gradient buffer of the first (no per-parameter adds, no gather)."""
    oc, mapper, chunk, has_head, B, wseed, iseed, xlam = cases()["tiny_chammi_hpa"]
    weights = O.make_weights(oc, has_head, wseed)
    xs = {c: make_inputs(oc, 3, len(ch), oc.num_classes, 90 + i) for i, (c, ch) in enumerate(CHAMMI_MAPPER.items())}
    sep = build_cuda_model(oc, CHAMMI_MAPPER, weights)
    total = {}
    for c, (x, y) in xs.items():
        _, _, _, g = cuda_step(sep, x.cuda(), y.cuda(), c, has_head, xlam)
        for k, v in g.items():
            if v is not None:
                total[k] = total.get(k, 0) + v.detach().clone()
    m = build_cuda_model(oc, CHAMMI_MAPPER, weights)
    m.direct_grad = True
    m.train()
    m.zero_grad(set_to_none=True)
    ptrs = set()
    for c, (x, y) in xs.items():
        out, extra = m(x.cuda(), c)
        training_loss(m, out, extra, y.cuda(), has_head, xlam).backward()
        ptrs.add(m._last_gflat.data_ptr())
    assert len(ptrs) == 1  # one buffer for the whole step
    for k, p in m.named_parameters():
        assert rel_l2(p.grad, total[k]) < 2e-3, k
    opt = FusedAdamW(m, lr=1e-3)
    g, ranges = opt._collect()
    assert g.data_ptr() == m._last_gflat.data_ptr()


def test_learnable_temperature_end_to_end():
    """learnable_temp=True (reference dichavit.py:807-808, trainer.py:876-883): the module owns `logit_scale` instead
    of `scale`; the trainer's proxy loss uses exp(logit_scale); its gradient and the optimiser update match torch."""
    oc, mapper, chunk, has_head, B, wseed, iseed, xlam = cases()["tiny_chammi_hpa"]
    weights = O.make_weights(oc, has_head, wseed)
    cfg = ref_cfg(oc)
    cfg["learnable_temp"] = True
    m = dichavit(cfg, mapper=mapper)
    assert "logit_scale" in m.state_dict() and not hasattr(m, "scale")
    assert float(m.logit_scale.detach()) == pytest.approx(float(torch.log(torch.tensor(1 / oc.temperature))), rel=1e-6)
    m.load_state_dict({k: weights[k].clone() for k in weights}, strict=False)
    m = m.cuda().train()
    x, y = make_inputs(oc, B, len(mapper[chunk]), oc.num_classes, iseed)
    out, extra = m(x.cuda(), chunk)
    training_loss(m, out, extra, y.cuda(), has_head, xlam).backward()
    # oracle: same forward, the trainer glue with a learnable scale, autograd on the CPU
    p = {k: v.clone().requires_grad_(True) for k, v in weights.items() if k != "adaptive_interface.0"}
    ls = m.logit_scale.detach().cpu().clone().requires_grad_(True)
    oo = O.forward(x, p, oc, mapper[chunk], training=True, has_head=has_head)
    (O.proxy_loss(p["proxies"], oo.out, y, ls.exp()) + oo.extra_loss * xlam).backward()
    assert rel_l2(m.logit_scale.grad, ls.grad) < 2e-2
    assert rel_l2(m.proxies.grad, p["proxies"].grad) < 2e-2
    before = float(m.logit_scale)
    opt = FusedAdamW(m, lr=1e-2, weight_decay=0.0)
    opt.step()
    assert abs(float(m.logit_scale) - before) == pytest.approx(1e-2, rel=1e-3)  # first Adam step: lr * sign(grad)


@pytest.mark.parametrize("t_in_epochs", [True, False])
def test_device_resident_schedule_equals_reference_training_loop(t_in_epochs):
    """device_schedule=True: lr / weight decay / bias corrections are evaluated by a one-thread kernel from a device
    update counter (CUDA-graph friendly).  Sequence == the reference's loop (oracle restatement of timm + utils.py),
    parameters == the host-driven FusedAdamW fed the same gradients."""
    make, x, y, chunk, has_head, xlam = _pair("tiny_jumpcp")
    upe, epochs = 3, 4
    kw = dict(lr_min=1e-6, warmup_t=1 if t_in_epochs else 2, warmup_lr_init=1e-5, cycle_decay=0.5, cycle_limit=1)
    t_initial = epochs if t_in_epochs else epochs * upe
    want = O.trainer_lr_wd_sequence(upe * epochs, upe, epochs, 4e-3, 0.04, 0.4, dict(t_initial=t_initial, **kw), t_in_epochs)
    ma, mb = make(), make()

    def mk(m, dev):
        return FusedAdamW(m, lr=4e-3, weight_decay=0.04, clip_grad_norm=1.0, updates_per_epoch=upe, device_schedule=dev,
                          lr_schedule=CosineLRSchedule(4e-3, t_initial, t_in_epochs=t_in_epochs, **kw),
                          wd_schedule=CosineWDSchedule(0.04, 0.4, epochs, upe))

    oa, ob = mk(ma, True), mk(mb, False)
    mb._ensure_flat(x.device)  # mb never runs a forward here: it is fed ma's gradients
    u = 0
    for epoch in range(1, epochs + 1):
        ob.step_epoch(epoch)
        for bid in range(1, upe + 1):
            u += 1
            cuda_step(ma, x, y, chunk, has_head, xlam)
            for pa, pb in zip(ma.parameters(), mb.parameters()):
                pb.grad = None if pa.grad is None else pa.grad.detach().clone()
            oa.step()
            ob.step()
            ob.step_update(u)
            st = oa.device_state()
            assert st["num_updates"] == u
            assert st["lr"] == pytest.approx(want[u - 1][0], rel=1e-5)
            assert st["wd"] == pytest.approx(want[u - 1][1], rel=1e-5)
    for (ka, pa), (kb, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        if pb.grad is not None:
            assert rel_l2(pa, pb) < 1e-5, ka
