"""CUDA-graph training step for the drop-in DiChaViT: one graph launch per (chunk, batch shape, C') bucket instead of
~250 kernel launches enqueued from Python.

What is captured: gradient-buffer clear, the device half of Diverse Channel Sampling (cosine / softmax /
torch.multinomial on the CUDA generator -- graph-safe Philox state, same draws as the eager path), the kernels' forward,
the trainer's loss glue (plain torch), the kernels' backward (+ the NCCL gradient all-reduce in data-parallel mode) and
the fused AdamW update with its device-resident schedule.  What stays on the host: the two `random` draws of DCS
(reference dichavit.py:150-153) -- they pick WHICH graph to launch (C' fixes the sequence length); the anchor channel
travels through a one-element device tensor written by a fill kernel before the launch.

Buckets: JUMP-CP 8 (C' = 1..8), So2Sat 18, CHAMMI 3 + 4 + 5; each is captured the first time it is drawn (one extra
eager forward/backward as warm-up, RNG state restored afterwards).  All graphs share one memory pool: activations are
dead between steps, so the pool's size is the largest bucket's, not the sum.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import DcvError
from .optim import FusedAdamW
from .trainer_glue import training_loss


class _Bucket:
    __slots__ = ("graph", "n_kernels", "loss", "anchor", "pos")

    def __init__(self):
        self.graph = None
        self.n_kernels = 0
        self.loss = None
        self.anchor = None
        self.pos = None


class GraphedTrainStep:
    """step = GraphedTrainStep(model, FusedAdamW(model, ..., device_schedule=True))
    loss = step(x, y, "train")                                  # one optimiser step
    step(x1, y1, "Allen", last=False); step(x2, y2, "HPA", last=False); loss = step(x3, y3, "CP")   # CHAMMI

    `x`, `y`: device tensors (copied into the graph's static input buffers; upload from pinned host memory straight
    into `step.input_buffers(chunk, x.shape, x.dtype)` to skip that copy).  Returns the step's loss as a 0-dim device
    tensor that the NEXT call overwrites.  `loss_fn(model, out, extra, y) -> loss`: the trainer's glue; default =
    trainer.py:986-995 / :912-914 through trainer_glue.training_loss."""

    def __init__(self, model, optimizer: FusedAdamW, loss_fn: Optional[Callable] = None, extra_loss_lambda: float = 1.0):
        if not optimizer.device_schedule:
            raise DcvError("GraphedTrainStep needs FusedAdamW(device_schedule=True): a captured update must read lr / "
                           "weight decay / bias corrections from device memory")
        self.model, self.opt = model, optimizer
        has_head = isinstance(model.classifer_head, torch.nn.Linear)
        self.loss_fn = loss_fn or (lambda m, out, extra, y: training_loss(m, out, extra, y, has_head, extra_loss_lambda))
        self.buckets: Dict[tuple, _Bucket] = {}
        self.inputs: Dict[tuple, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.pool = None
        self.first = True            # next call starts an optimiser step (clears the gradient buffer)
        self.kernel_launches = 0     # kernels of this library executed through graph replays
        self.graph_launches = 0
        self._ready = False

    # ------------------------------------------------------------------ static state
    def _prepare(self, device) -> None:
        m = self.model
        m.train()
        m.direct_grad = True
        m._ensure_flat(device)
        m._static_gflat = torch.zeros_like(m._flat)
        m._last_gflat = m._static_gflat
        # every trainable parameter's .grad becomes its view of the static flat buffer: the kernels accumulate into it
        # directly, torch autograd (proxies / logit_scale from the trainer's loss glue) accumulates in place
        for (p, off, n) in m._layout:
            p.grad = m._static_gflat[off:off + n].view(p.shape) if p.requires_grad else None
        self.pool = torch.cuda.graph_pool_handle()
        self._stream = torch.cuda.Stream(device=device)
        self._ready = True

    def input_buffers(self, chunk_name: str, shape, dtype=torch.float32, device=None):
        key = (chunk_name, tuple(shape), dtype)
        buf = self.inputs.get(key)
        if buf is None:
            dev = device or self.model._flat.device
            buf = (torch.zeros(tuple(shape), dtype=dtype, device=dev), torch.zeros(shape[0], dtype=torch.int64, device=dev))
            self.inputs[key] = buf
        return buf

    def _update_ranges(self):
        m = self.model
        ext = m._external_ids
        ranges = []
        for p, off, n in m._layout:
            if not p.requires_grad:
                continue
            if id(p) in ext and not self._external_used(p):
                continue  # e.g. `proxies` on a model with a classifier head: never receives a gradient -> left alone
            end = min(off + (n + 63) // 64 * 64, m._flat.numel())
            if ranges and ranges[-1][1] == off:
                ranges[-1][1] = end
            else:
                ranges.append([off, end])
        return [tuple(r) for r in ranges]

    def _external_used(self, p) -> bool:
        return id(p) in self._ext_seen

    # ------------------------------------------------------------------ one micro-step
    def _run(self, x, y, chunk_name, draw, bucket: Optional[_Bucket], first: bool, last: bool, capture: bool):
        """the work of one micro-step on the current stream (eager warm-up and graph capture run the same code)"""
        m, pe = self.model, self.model.feature_extractor.patch_embed
        if first:
            m._static_gflat.zero_()
        n_in = x.shape[1]
        if draw is not None:
            sel = pe.select_device(chunk_name, x.device, draw, anchor_dev=bucket.anchor, pos_dev=bucket.pos)
            pe._prefetched = ((chunk_name, n_in, str(x.device), pe.training, pe.enable_sample), sel)
        out, extra = m(x, chunk_name)
        loss = self.loss_fn(m, out, extra, y)
        if m.grad_allreduce and not last:
            with m.no_sync():
                loss.backward()
        else:
            loss.backward()
        if last:
            if m.grad_allreduce:
                for p in (m.proxies, getattr(m, "logit_scale", None)):
                    if p is not None and id(p) in self._ext_seen:
                        m._allreduce_external(p.grad)
            self.opt.step(ranges=self._ranges, grad=m._static_gflat)
        bucket.loss.copy_(loss.detach())

    def __call__(self, x: torch.Tensor, y: torch.Tensor, chunk_name: str, last: bool = True,
                 eager: bool = False) -> torch.Tensor:
        """One micro-step (`last=True`: the optimiser update follows the backward).  `eager=True` runs the identical
        sequence as ordinary launches instead of a graph replay (profiling with the built-in per-kernel profiler,
        debugging): same host RNG consumption, same static buffers."""
        m = self.model
        if not x.is_cuda:
            raise DcvError("GraphedTrainStep needs device tensors (upload into step.input_buffers(...))")
        if not self._ready:
            self._prepare(x.device)
        pe = m.feature_extractor.patch_embed
        if not m.training:
            m.train()
        xs, ys = self.input_buffers(chunk_name, x.shape, x.dtype, x.device)
        if x.data_ptr() != xs.data_ptr():
            xs.copy_(x, non_blocking=True)
        if y.data_ptr() != ys.data_ptr():
            ys.copy_(y, non_blocking=True)
        draw = pe.draw_host(chunk_name, x.shape[1])  # python RNG, in the reference's order
        first = self.first
        self.first = last
        b = self._bucket(xs, ys, chunk_name, draw, first, last, make=not eager)
        if draw is not None:
            if draw["mode"] == "none":
                b.pos.copy_(torch.tensor(draw["pos"], dtype=torch.int32), non_blocking=False)
            else:
                b.anchor.fill_(draw["anchor"])  # a fill kernel with an immediate: no pinned staging, no race with replays
        if eager or b.graph is None:
            self._run(xs, ys, chunk_name, draw, b, first, last, capture=False)
            return b.loss
        b.graph.replay()
        self.kernel_launches += b.n_kernels
        self.graph_launches += 1
        return b.loss

    def _bucket(self, xs, ys, chunk_name, draw, first, last, make: bool) -> _Bucket:
        cs = draw["c_new"] if draw is not None else xs.shape[1]
        key = (chunk_name, tuple(xs.shape), xs.dtype, cs, draw["mode"] if draw else None, first, last)
        b = self.buckets.get(key)
        if b is None and make:
            b = self._capture(key, xs, ys, chunk_name, draw, first, last)
        if b is None:  # eager only: static tensors of the bucket without a graph
            b = self._eager_buckets.get(key) if hasattr(self, "_eager_buckets") else None
            if b is None:
                b = self._new_bucket(xs.device, draw)
                self.__dict__.setdefault("_eager_buckets", {})[key] = b
                if not hasattr(self, "_ext_seen"):
                    self._ext_seen = set()
                    self._probe_external(xs, ys, chunk_name, draw, b)
                    self._ranges = self._update_ranges()
        return b

    @staticmethod
    def _new_bucket(dev, draw) -> _Bucket:
        b = _Bucket()
        b.loss = torch.zeros((), dtype=torch.float32, device=dev)
        if draw is not None:
            if draw["mode"] == "none":
                b.pos = torch.tensor(draw["pos"], dtype=torch.int32, device=dev)
            else:
                b.anchor = torch.full((1,), draw["anchor"], dtype=torch.int64, device=dev)
        return b

    def precapture(self, x: torch.Tensor, y: torch.Tensor, chunk_name: str, first: bool = True, last: bool = True) -> int:
        """Capture every C' bucket of this (chunk, shape) now instead of on first use (a capture costs one eager
        step + graph instantiation: keep it out of timed regions).  Consumes no RNG, changes no state."""
        if not self._ready:
            self._prepare(x.device)
        m, pe = self.model, self.model.feature_extractor.patch_embed
        m.train()
        xs, ys = self.input_buffers(chunk_name, x.shape, x.dtype, x.device)
        xs.copy_(x)
        ys.copy_(y)
        n_in = x.shape[1]
        mode = pe.cfg.hcs_sampling
        draws = [None]
        if pe.enable_sample:
            if mode in (None, "none"):
                chans = list(pe.mapper[chunk_name])
                draws = [dict(mode="none", c_new=c, pos=list(range(c)), cur=chans[:c]) for c in range(1, n_in + 1)]
            else:
                draws = [dict(mode=mode, c_new=c, anchor=0) for c in range(1, n_in + 1)]
        n = 0
        for d in draws:
            cs = d["c_new"] if d is not None else n_in
            key = (chunk_name, tuple(xs.shape), xs.dtype, cs, d["mode"] if d else None, first, last)
            if key not in self.buckets:
                self._capture(key, xs, ys, chunk_name, d, first, last)
                n += 1
        return n

    # ------------------------------------------------------------------ capture
    def _capture(self, key, xs, ys, chunk_name, draw, first, last) -> _Bucket:
        m = self.model
        dev = xs.device
        b = self._new_bucket(dev, draw)
        if not hasattr(self, "_ext_seen"):
            # which trainer-owned parameters the loss glue actually uses: probe once with a throw-away graph
            self._ext_seen = set()
            self._probe_external(xs, ys, chunk_name, draw, b)
            self._ranges = self._update_ranges()
        torch.cuda.synchronize(dev)
        cpu_state, cuda_state = torch.get_rng_state(), torch.cuda.get_rng_state(dev)
        opt_state = None if self.opt._state is None else self.opt._state.clone()
        # ---- warm-up: the same micro-step once, eagerly, on the capture stream (lazy initialisation of kernel
        # attributes, cuBLAS handles, the autograd thread); everything it changed is put back afterwards
        snap = (m._flat.clone(), m._static_gflat.clone(), None if self.opt.exp_avg is None else self.opt.exp_avg.clone(),
                None if self.opt.exp_avg_sq is None else self.opt.exp_avg_sq.clone())
        counter = getattr(m.feature_extractor.patch_embed, "counter", None)
        csnap = None if counter is None or counter._dev is None else counter._dev.clone()
        self._stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._stream):
            self._run(xs, ys, chunk_name, draw, b, first, last, capture=False)
        torch.cuda.current_stream(dev).wait_stream(self._stream)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            m._flat.copy_(snap[0])
            m._static_gflat.copy_(snap[1])
            if snap[2] is not None:
                self.opt.exp_avg.copy_(snap[2])
                self.opt.exp_avg_sq.copy_(snap[3])
            elif self.opt.exp_avg is not None:
                self.opt.exp_avg.zero_()
                self.opt.exp_avg_sq.zero_()
            if opt_state is not None:
                self.opt._state.copy_(opt_state)
            elif self.opt._state is not None:
                self.opt._state.zero_()
            if csnap is not None:
                counter._dev.copy_(csnap)
            elif counter is not None and counter._dev is not None:
                counter._dev.zero_()
        m.mark_params_dirty()
        self._refresh_operand_copy()
        torch.set_rng_state(cpu_state)
        torch.cuda.set_rng_state(cuda_state, dev)
        torch.cuda.synchronize(dev)
        # ---- capture
        k0 = _lib.launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=self.pool, stream=self._stream):
            self._run(xs, ys, chunk_name, draw, b, first, last, capture=True)
        b.graph = g
        b.n_kernels = _lib.launch_count() - k0
        torch.cuda.set_rng_state(cuda_state, dev)  # capture itself must not advance the generator
        self.buckets[key] = b
        return b

    def _refresh_operand_copy(self) -> None:
        """bf16 operand copy of the parameters current BEFORE capture, so that the cast is not captured into the graph
        (the captured AdamW kernel keeps it current from then on)"""
        import ctypes

        m = self.model
        _lib.check(_lib.lib().dcv_cast_f32_bf16(ctypes.c_void_p(m._flat.data_ptr()), ctypes.c_void_p(m._bflat.data_ptr()),
                                                ctypes.c_longlong(m._flat.numel()), _lib.stream_ptr()), "dcv_cast_f32_bf16")
        m._bflat_version = m._param_version()

    def _probe_external(self, xs, ys, chunk_name, draw, b) -> None:
        """Does the trainer's loss glue use `proxies` / `logit_scale`?  (CHAMMI: yes; classifier-head models: no.)
        Decided by autograd on a detached stand-in for the model output -- no kernel runs."""
        m = self.model
        has_head = isinstance(m.classifer_head, torch.nn.Linear)
        ncol = m.classifer_head.out_features if has_head else m.dim
        out = torch.zeros((xs.shape[0], ncol), device=xs.device, requires_grad=True)
        extra = torch.zeros((), device=xs.device, requires_grad=True)
        cand = [p for p in (m.proxies, getattr(m, "logit_scale", None)) if p is not None and p.requires_grad]
        if not cand:
            return
        ramp = torch.linspace(-1.0, 1.0, out.numel(), device=xs.device).reshape(out.shape)  # no RNG consumed
        loss = self.loss_fn(m, out + ramp, extra, ys)
        grads = torch.autograd.grad(loss, cand, allow_unused=True)
        for p, g in zip(cand, grads):
            if g is not None:
                self._ext_seen.add(id(p))
