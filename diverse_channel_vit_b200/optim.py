"""Fused AdamW for the drop-in DiChaViT: one kernel per contiguous trainable range of the module's
flat fp32 parameter buffer instead of a per-tensor multi-tensor apply.

Semantics = timm.optim.AdamW / torch.optim.AdamW as built by reference optimizers.py:20-21 (single parameter group,
decoupled weight decay; parameters without a gradient or with requires_grad=False are left untouched: no update, no
decay, not part of the clipping norm), optional global-norm clipping as trainer.py:1003-1004, the cosine learning-rate
schedule of timm's CosineLRScheduler (lr_schedulers.py:6-9, stepped per epoch trainer.py:344-348 or per update
trainer.py:1009-1010) and the per-update cosine weight-decay schedule (utils.py:563-574, trainer.py:1011-1019).

Two ways to drive the schedules:
  * host scalars (default): `opt.step()` then `opt.step_update(num_updates)` / `opt.step_epoch(epoch)` exactly where
    the reference's trainer calls its scheduler; lr / wd travel as kernel arguments (no copy, no sync);
  * `device_schedule=True`: the update counter, lr, wd and Adam's bias corrections live in a 32-byte device struct
    advanced by a one-thread kernel each step, so that a CUDA graph of the whole training step (graphs.py) replays
    with the right values without re-capture.
"""
from __future__ import annotations

import ctypes
import math
from ctypes import Structure, byref, c_float, c_int, c_longlong, c_void_p
from typing import List, Optional, Tuple

import torch

from . import _lib
from ._lib import DcvError, check


class CosineLRSchedule:
    """timm CosineLRScheduler._get_lr for one parameter group (cycle_mul == 1, no noise: the reference's
    configs/scheduler/cosine.yaml).  `t` is an epoch or an update count -- the caller decides, as timm's t_in_epochs."""

    def __init__(self, base_lr: float, t_initial: int, lr_min: float = 0.0, warmup_t: int = 0, warmup_lr_init: float = 0.0,
                 warmup_prefix: bool = False, cycle_decay: float = 1.0, cycle_limit: int = 1, k_decay: float = 1.0,
                 cycle_mul: float = 1.0, t_in_epochs: bool = True):
        if cycle_mul != 1.0:
            raise NotImplementedError("cycle_mul != 1 is outside the reference's configuration")
        if t_initial <= 0:
            raise ValueError("t_initial must be positive")
        self.base_lr, self.t_initial, self.lr_min = float(base_lr), int(t_initial), float(lr_min)
        self.warmup_t, self.warmup_lr_init, self.warmup_prefix = int(warmup_t), float(warmup_lr_init), bool(warmup_prefix)
        self.cycle_decay, self.cycle_limit, self.k_decay = float(cycle_decay), int(cycle_limit), float(k_decay)
        self.t_in_epochs = bool(t_in_epochs)

    def initial_lr(self) -> float:
        """the value timm's constructor leaves in the parameter group"""
        return self.warmup_lr_init if self.warmup_t > 0 else self.base_lr

    def value(self, t: int) -> float:
        if t < self.warmup_t:
            return self.warmup_lr_init + t * ((self.base_lr - self.warmup_lr_init) / self.warmup_t)
        if self.warmup_prefix:
            t = t - self.warmup_t
        i = t // self.t_initial
        t_curr = t - self.t_initial * i
        lr_max = self.base_lr * self.cycle_decay ** i
        if i < self.cycle_limit:
            return self.lr_min + 0.5 * (lr_max - self.lr_min) * (
                1 + math.cos(math.pi * t_curr ** self.k_decay / self.t_initial ** self.k_decay))
        return self.lr_min


class CosineWDSchedule:
    """utils.cosine_scheduler(weight_decay, weight_decay_end, epochs, updates_per_epoch) (utils.py:563-574) read the
    way trainer.py:1011-1019 reads it: after update u the group's weight decay becomes table[min(u - 1, len - 1)]."""

    def __init__(self, wd_base: float, wd_end: float, epochs: int, updates_per_epoch: int):
        self.wd_base, self.wd_end, self.total = float(wd_base), float(wd_end), int(epochs) * int(updates_per_epoch)
        if self.total <= 0:
            raise ValueError("epochs * updates_per_epoch must be positive")

    def after_update(self, num_updates: int) -> float:
        idx = min(num_updates - 1, self.total - 1)
        return self.wd_end + 0.5 * (self.wd_base - self.wd_end) * (1 + math.cos(math.pi * idx / self.total))


class _Sched(Structure):  # dcv_sched, include/dcvit.h
    _fields_ = [("base_lr", c_float), ("lr_min", c_float), ("warmup_lr_init", c_float), ("t_initial", c_int),
                ("warmup_t", c_int), ("warmup_prefix", c_int), ("cycle_limit", c_int), ("cycle_decay", c_float),
                ("k_decay", c_float), ("updates_per_epoch", c_int), ("wd_base", c_float), ("wd_end", c_float),
                ("wd_total", c_int), ("beta1", c_float), ("beta2", c_float)]


class FusedAdamW:
    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 clip_grad_norm: Optional[float] = None, lr_schedule: Optional[CosineLRSchedule] = None,
                 wd_schedule: Optional[CosineWDSchedule] = None, updates_per_epoch: int = 0,
                 device_schedule: bool = False):
        self.model = model
        self.base_lr, self.betas, self.eps, self.base_wd = float(lr), betas, float(eps), float(weight_decay)
        self.lr_schedule, self.wd_schedule = lr_schedule, wd_schedule
        self.lr = lr_schedule.initial_lr() if lr_schedule is not None else float(lr)
        self.weight_decay = float(weight_decay)
        self.clip_grad_norm = clip_grad_norm
        self.updates_per_epoch = int(updates_per_epoch)
        self.device_schedule = bool(device_schedule)
        if device_schedule and lr_schedule is not None and lr_schedule.t_in_epochs and self.updates_per_epoch <= 0:
            raise ValueError("a per-epoch schedule evaluated on the device needs updates_per_epoch")
        self.step_count = 0
        self.exp_avg = None
        self.exp_avg_sq = None
        self._clip = None
        self._state = None  # device dcv_optim_state
        self._gbuf = None   # flat gradient buffer used when gradients have to be gathered

    # ------------------------------------------------------------------ schedules driven from the host
    def step_epoch(self, epoch: int) -> None:
        """`scheduler.step(epoch)` of trainer.py:344-348 (acts only on a schedule counted in epochs)."""
        if self.lr_schedule is not None and self.lr_schedule.t_in_epochs:
            self.lr = self.lr_schedule.value(epoch)

    def step_update(self, num_updates: int) -> None:
        """`scheduler.step_update(num_updates)` + the weight-decay table lookup of trainer.py:1009-1019, called
        after `step()` like the reference does.  As there, the weight decay only moves when a scheduler exists."""
        if self.lr_schedule is None:
            return
        if not self.lr_schedule.t_in_epochs:
            self.lr = self.lr_schedule.value(num_updates)
        if self.wd_schedule is not None:
            self.weight_decay = self.wd_schedule.after_update(num_updates)

    def device_state(self) -> Optional[dict]:
        """{num_updates, lr, wd} read back from the device struct (synchronises; for logging / tests)."""
        if self._state is None:
            return None
        raw = self._state.cpu()
        return {"num_updates": int(raw.view(torch.int32)[0]), "lr": float(raw[1]), "wd": float(raw[2])}

    def _sched_struct(self) -> _Sched:
        s, w = self.lr_schedule, self.wd_schedule
        upe = self.updates_per_epoch if (s is not None and s.t_in_epochs) else 0
        return _Sched(self.base_lr if s is None else s.base_lr, 0.0 if s is None else s.lr_min,
                      0.0 if s is None else s.warmup_lr_init, 0 if s is None else s.t_initial,
                      0 if s is None else s.warmup_t, 0 if s is None else int(s.warmup_prefix),
                      1 if s is None else s.cycle_limit, 1.0 if s is None else s.cycle_decay,
                      1.0 if s is None else s.k_decay, upe, self.base_wd, 0.0 if w is None else w.wd_end,
                      0 if (w is None or s is None) else w.total, float(self.betas[0]), float(self.betas[1]))

    # ------------------------------------------------------------------ gradients
    def zero_grad(self, set_to_none: bool = True):
        if set_to_none:  # nn.Module.zero_grad walks named_parameters(): 0.7 ms of host time per step for 150 tensors
            for p, _, _ in self.model._layout or [(q, 0, 0) for q in self.model.parameters()]:
                p.grad = None
        else:
            self.model.zero_grad(set_to_none=False)

    def _collect(self) -> Tuple[torch.Tensor, List[Tuple[int, int]]]:
        """Flat gradient buffer + the merged element ranges [start, end) of the parameters to update: those with
        requires_grad and a gradient (torch / timm AdamW skip the others entirely).  A gradient that is not already
        the corresponding view of the flat buffer of the last backward (accumulated over several backwards, produced
        by torch autograd for the trainer-owned `proxies` / `logit_scale`, or written by DDP) is copied in."""
        m = self.model
        g = getattr(m, "_last_gflat", None)
        if g is None or g.numel() != m._flat.numel() or g.device != m._flat.device:
            if self._gbuf is None or self._gbuf.numel() != m._flat.numel() or self._gbuf.device != m._flat.device:
                self._gbuf = torch.zeros_like(m._flat)
            g = self._gbuf
        base = g.data_ptr()
        ranges: List[List[int]] = []
        any_grad = False
        for p, off, n in m._layout:
            gr = p.grad
            if gr is None or not p.requires_grad:
                continue
            any_grad = True
            if gr.data_ptr() != base + 4 * off or not gr.is_contiguous():
                g[off:off + n].copy_(gr.reshape(-1))
                if m.grad_allreduce and id(p) in m._external_ids:
                    m._allreduce_external(g[off:off + n])
            end = off + (n + 63) // 64 * 64  # parameters are laid out 64-element aligned; the pad stays zero
            if ranges and ranges[-1][1] == off:
                ranges[-1][1] = end
            else:
                ranges.append([off, end])
        if not any_grad:
            raise DcvError("FusedAdamW.step(): no gradients (call backward first)")
        total = m._flat.numel()
        return g, [(a, min(b, total)) for a, b in ranges]

    # ------------------------------------------------------------------ update
    @torch.no_grad()
    def step(self, ranges: Optional[List[Tuple[int, int]]] = None, grad: Optional[torch.Tensor] = None):
        """One AdamW update.  `ranges` / `grad`: skip the per-parameter gradient inspection and update these element
        ranges of the flat buffers from this flat gradient (used by the captured-graph step, where they are fixed)."""
        m = self.model
        if m._flat is None:
            raise DcvError("FusedAdamW.step(): the module has not run on a CUDA device yet")
        if ranges is None or grad is None:
            g, ranges = self._collect()
        else:
            g = grad
        dev = m._flat.device
        if self.exp_avg is None or self.exp_avg.numel() != m._flat.numel() or self.exp_avg.device != dev:
            self.exp_avg = torch.zeros_like(m._flat)
            self.exp_avg_sq = torch.zeros_like(m._flat)
        self.step_count += 1
        lib = _lib.lib()
        st = _lib.stream_ptr()
        clip_ptr = None
        if self.clip_grad_norm is not None:
            if self._clip is None or self._clip.device != dev:
                self._clip = torch.zeros(2, dtype=torch.float32, device=dev)
                self._clip_init = torch.tensor([0.0, float(self.clip_grad_norm)], dtype=torch.float32, device=dev)
            self._clip.copy_(self._clip_init)
            for a, b in ranges:
                check(lib.dcv_sumsq_f32(c_void_p(g.data_ptr() + 4 * a), c_longlong(b - a), c_void_p(self._clip.data_ptr()), st),
                      "dcv_sumsq_f32")
            clip_ptr = c_void_p(self._clip.data_ptr())
            if getattr(m, "grad_allreduce", False):
                # The sum of squares is accumulated with fp32 atomics: its last bit depends on the order the CTAs finish
                # in, so two replicas holding bit-identical averaged gradients can get clip factors one ulp apart and
                # drift away from each other for good (measured: 138 of 154 parameters off by an ulp after 8 steps on
                # 2 GPUs).  One 4-byte MAX all-reduce makes every rank use the same value (DDP keeps replicas
                # bit-identical, reference trainer.py:1185).
                import torch.distributed as dist

                if dist.is_available() and dist.is_initialized():
                    dist.all_reduce(self._clip[0:1], op=dist.ReduceOp.MAX, group=getattr(m, "_pg", None))
        if self.device_schedule:
            if self._state is None or self._state.device != dev:
                self._state = torch.zeros(8, dtype=torch.float32, device=dev)
            sc = self._sched_struct()
            check(lib.dcv_optim_sched_step(c_void_p(self._state.data_ptr()), byref(sc), st), "dcv_optim_sched_step")
        fp, gp = m._flat.data_ptr(), g.data_ptr()
        mp, vp, bp = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), m._bflat.data_ptr()
        for a, b in ranges:
            if self.device_schedule:
                check(lib.dcv_adamw_step_dev(c_void_p(fp + 4 * a), c_void_p(gp + 4 * a), c_void_p(mp + 4 * a),
                                             c_void_p(vp + 4 * a), c_void_p(bp + 2 * a), c_longlong(b - a),
                                             c_float(self.betas[0]), c_float(self.betas[1]), c_float(self.eps),
                                             c_void_p(self._state.data_ptr()), clip_ptr, st), "dcv_adamw_step_dev")
            else:
                check(lib.dcv_adamw_step(c_void_p(fp + 4 * a), c_void_p(gp + 4 * a), c_void_p(mp + 4 * a),
                                         c_void_p(vp + 4 * a), c_void_p(bp + 2 * a), c_longlong(b - a), c_float(self.lr),
                                         c_float(self.betas[0]), c_float(self.betas[1]), c_float(self.eps),
                                         c_float(self.weight_decay), self.step_count, clip_ptr, st), "dcv_adamw_step")
        m._bflat_version = m._param_version()  # the kernel refreshed the bf16 operand copy (no torch-side version bump)
