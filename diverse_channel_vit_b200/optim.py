"""Fused AdamW for the drop-in DiChaViT: one kernel over the module's flat fp32 parameter buffer
instead of a per-tensor multi-tensor apply.  Semantics = timm.optim.AdamW / torch.optim.AdamW as built by reference
optimizers.py:20-21 (single parameter group, decoupled weight decay), optional global-norm clipping as
trainer.py:1003-1004, and the per-update cosine schedule of timm's CosineLRScheduler (lr_schedulers.py:6-9)."""
from __future__ import annotations

import math
from ctypes import c_float, c_longlong, c_void_p
from typing import Optional

import torch

from . import _lib
from ._lib import DcvError, check


class FusedAdamW:
    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 clip_grad_norm: Optional[float] = None):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        self.clip_grad_norm = clip_grad_norm
        self.step_count = 0
        self.exp_avg = None
        self.exp_avg_sq = None
        self._clip = None

    def zero_grad(self, set_to_none: bool = True):
        if set_to_none:  # nn.Module.zero_grad walks named_parameters(): 0.7 ms of host time per step for 150 tensors
            for p, _, _ in self.model._layout or [(q, 0, 0) for q in self.model.parameters()]:
                p.grad = None
        else:
            self.model.zero_grad(set_to_none=False)

    def _flat_grad(self) -> torch.Tensor:
        m = self.model
        g = getattr(m, "_last_gflat", None)
        first = next((p for p, _, _ in m._layout if p.grad is not None), None)
        if first is None:
            raise DcvError("FusedAdamW.step(): no gradients (call backward first)")
        if g is not None and first.grad.data_ptr() == g.data_ptr() + 4 * m._off[id(first)]:
            return g  # the views autograd stored in .grad alias the flat buffer of the last backward
        g = torch.zeros_like(m._flat)  # gradients came from elsewhere (accumulation, DDP): gather them
        for p, off, n in m._layout:
            if p.grad is not None:
                g[off:off + n].copy_(p.grad.reshape(-1))
        return g

    @torch.no_grad()
    def step(self):
        m = self.model
        if m._flat is None:
            raise DcvError("FusedAdamW.step(): the module has not run on a CUDA device yet")
        g = self._flat_grad()
        if self.exp_avg is None or self.exp_avg.numel() != m._flat.numel() or self.exp_avg.device != m._flat.device:
            self.exp_avg = torch.zeros_like(m._flat)
            self.exp_avg_sq = torch.zeros_like(m._flat)
        self.step_count += 1
        lib = _lib.lib()
        st = _lib.stream_ptr()
        n = m._flat.numel()
        clip_ptr = None
        if self.clip_grad_norm is not None:
            if self._clip is None or self._clip.device != g.device:
                self._clip = torch.zeros(2, dtype=torch.float32, device=g.device)
            self._clip.zero_()
            self._clip[1] = float(self.clip_grad_norm)
            check(lib.dcv_sumsq_f32(c_void_p(g.data_ptr()), c_longlong(n), c_void_p(self._clip.data_ptr()), st), "dcv_sumsq_f32")
            clip_ptr = c_void_p(self._clip.data_ptr())
        check(lib.dcv_adamw_step(c_void_p(m._flat.data_ptr()), c_void_p(g.data_ptr()), c_void_p(self.exp_avg.data_ptr()),
                                 c_void_p(self.exp_avg_sq.data_ptr()), c_void_p(m._bflat.data_ptr()), c_longlong(n),
                                 c_float(self.lr), c_float(self.betas[0]), c_float(self.betas[1]), c_float(self.eps),
                                 c_float(self.weight_decay), self.step_count, clip_ptr, st), "dcv_adamw_step")
        m._bflat_version = m._param_version()  # the kernel refreshed the bf16 operand copy (no torch-side version bump)


def cosine_lr(num_updates: int, base_lr: float, t_initial: int, lr_min: float = 0.0, warmup_t: int = 0,
              warmup_lr_init: float = 0.0) -> float:
    """timm CosineLRScheduler.step_update value for one parameter group (cycle_limit=1, no noise), as configured by
    reference lr_schedulers.py:6-9 and called per update at trainer.py:1009-1011."""
    if num_updates < warmup_t:
        return warmup_lr_init + num_updates * (base_lr - warmup_lr_init) / warmup_t
    t = num_updates - warmup_t if False else num_updates
    if t >= t_initial:
        return lr_min
    return lr_min + 0.5 * (base_lr - lr_min) * (1 + math.cos(math.pi * t / t_initial))
