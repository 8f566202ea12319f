"""CPU: the learning-rate / weight-decay schedules of the fused optimiser against the oracle's
restatement of timm's CosineLRScheduler (third-party, pinned 0.8.3.dev0 by the reference) and of the reference's
utils.cosine_scheduler, driven in the order the reference's training loop uses (trainer.py:344-348, :1006-1019)."""
import types

import pytest

from oracle import dichavit_oracle as O
from diverse_channel_vit_b200.optim import CosineLRSchedule, CosineWDSchedule, FusedAdamW

COSINE_YAML = dict(lr_min=1e-6, warmup_t=3, warmup_lr_init=1e-5, warmup_prefix=False, cycle_decay=0.5, cycle_limit=1,
                   k_decay=1.0)  # reference configs/scheduler/cosine.yaml


@pytest.mark.parametrize("kw", [COSINE_YAML, dict(COSINE_YAML, warmup_prefix=True), dict(COSINE_YAML, warmup_t=0),
                                dict(COSINE_YAML, cycle_limit=2), dict(COSINE_YAML, k_decay=1.5)])
def test_cosine_lr_equals_timm_restatement(kw):
    s = CosineLRSchedule(4e-4, 17, **kw)
    for t in range(0, 60):
        want = O.timm_cosine_lr(t, 4e-4, 17, **kw)
        assert s.value(t) == pytest.approx(want, rel=1e-12, abs=0), t


def test_weight_decay_table_equals_reference_cosine_scheduler():
    table = O.cosine_scheduler(0.04, 0.4, 7, 13)  # configs/optimizer/adamw_jumpcp.yaml: 0.04 -> 0.4
    w = CosineWDSchedule(0.04, 0.4, 7, 13)
    for u in range(1, 7 * 13 + 20):
        assert w.after_update(u) == pytest.approx(table[min(u - 1, len(table) - 1)], rel=1e-12)


@pytest.mark.parametrize("t_in_epochs", [True, False])
def test_host_driven_sequence_equals_reference_training_loop(t_in_epochs):
    """lr / wd in force at every update when the trainer calls step_epoch / step / step_update where the reference
    calls scheduler.step(epoch) / optimizer.step() / scheduler.step_update(num_updates)."""
    upe, epochs = 5, 6
    t_initial = epochs if t_in_epochs else epochs * upe
    kw = dict(COSINE_YAML, warmup_t=COSINE_YAML["warmup_t"] * (1 if t_in_epochs else upe))
    want = O.trainer_lr_wd_sequence(upe * epochs, upe, epochs, 4e-4, 0.04, 0.4, dict(t_initial=t_initial, **kw), t_in_epochs)
    opt = FusedAdamW(types.SimpleNamespace(_layout=[], _flat=None), lr=4e-4, weight_decay=0.04,
                     lr_schedule=CosineLRSchedule(4e-4, t_initial, t_in_epochs=t_in_epochs, **kw),
                     wd_schedule=CosineWDSchedule(0.04, 0.4, epochs, upe), updates_per_epoch=upe)
    got = []
    for epoch in range(1, epochs + 1):
        opt.step_epoch(epoch)
        for bid in range(1, upe + 1):
            num_updates = (epoch - 1) * upe + bid
            got.append((opt.lr, opt.weight_decay))  # what optimizer.step() would use now
            opt.step_update(num_updates)
    for u, (g, w) in enumerate(zip(got, want), 1):
        assert g[0] == pytest.approx(w[0], rel=1e-12), u
        assert g[1] == pytest.approx(w[1], rel=1e-12), u
