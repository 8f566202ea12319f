#!/usr/bin/env python
"""bench.py -- DiChaViT training hot path on B200 (metric of BASELINE.json: train images/sec, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload jumpcp|chammi|so2sat|vitb]

Workload at N GPUs (weak scaling, SURVEY.md section 8(d) "C3"): DiChaViT ViT-S/16, JUMP-CP shape (8 channels, 224x224,
161 classes, 1568 channel-patch tokens at full channels), 32 images per GPU, DCS channel sampling
(lowest_cosine_prob, temp 1000) + CDL (lambda 0.001) + TDL (lambda 0.001, gamma 1/4), CE loss, seed 2025.
A step = zero_grad + forward + loss + backward (+ gradient all-reduce for N > 1) + AdamW update, on synthetic
randn images / randint labels with the module's own random init (datasets and checkpoints are offline).

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM.  `e2e`: same step driven from pinned HOST
buffers through the public module API (H2D of the batch and D2H of the loss inside the timed region).
`roofline`: the dominant kernel class, timed live with CUDA events by the library's built-in profiler in a
separate pass of the same steps.  `cpu_baseline`: the oracle restatement of the reference (torch fp32, all
host cores) on a bounded sample.  `--impl reference` prints the reference arm (same CPU path) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "train images/sec (fwd+bwd)"
UNIT = "images/s"


class Cfg(dict):
    __getattr__ = dict.__getitem__


WORKLOADS = {
    # name: (model size, img, patch, channels, classes, per-GPU batch, hcs temp, lambda_cdl, lambda_tdl, gamma_s, gamma_d, temperature, chammi)
    "jumpcp": dict(size="small", img=224, patch=16, channels=8, classes=161, batch=32, hcs_temp=1000.0, l_cdl=0.001,
                   l_tdl=0.001, gs=1.0, gd=4.0, sample=True,
                   desc="DiChaViT ViT-S/16 JUMP-CP 8ch 224x224 (<=1569 tokens), 32 img/GPU, DCS+CDL+TDL, CE-161"),
    "so2sat": dict(size="small", img=32, patch=8, channels=18, classes=17, batch=128, hcs_temp=0.01, l_cdl=0.001,
                   l_tdl=0.1, gs=0.5, gd=4.0, sample=True,
                   desc="DiChaViT ViT-S/8 So2Sat 18ch 32x32 (<=289 tokens), 128 img/GPU, DCS+CDL+TDL, CE-17"),
    # BASELINE configs[1]: one optimiser step = three fwd+bwd (one per chunk, batch 64 split 22/21/21), proxy loss
    "chammi": dict(size="small", img=224, patch=16, channels=12, classes=14, batch=64, hcs_temp=0.1, l_cdl=0.1,
                   l_tdl=1.0, gs=0.5, gd=2.0, sample=True, temperature=0.07,
                   chunks={"Allen": ([0, 1, 2], 22), "HPA": ([3, 4, 5, 6], 21), "CP": ([7, 8, 9, 10, 11], 21)},
                   desc="DiChaViT ViT-S/16 CHAMMI mixed 3/4/5-channel chunks (Allen/HPA/CP, 22+21+21 img), DCS+CDL+TDL, "
                        "proxy loss, 3 fwd+bwd per optimiser step"),
    "vitb": dict(size="base", img=224, patch=16, channels=8, classes=161, batch=16, hcs_temp=0.1, l_cdl=0.0,
                 l_tdl=0.0, gs=1.0, gd=0.5, sample=False,
                 desc="DiChaViT ViT-B/16 JUMP-CP 8ch full channels (1569 tokens), 16 img/GPU, CE-161"),
}


def model_cfg(w) -> Cfg:
    return Cfg(name="dichavit", pretrained=False, pretrained_model_name=w["size"], in_dim=None, num_classes=w["classes"],
               pooling="avg", temperature=w.get("temperature", 0.11111), learnable_temp=False, unfreeze_last_n_layers=-1,
               unfreeze_first_layer=True, init_first_layer=None, reset_last_n_unfrozen_layers=False,
               enable_sample=w["sample"], in_channel_names=[f"c{i}" for i in range(w["channels"])],
               new_channel_inits=None, use_channelvit_channels=True, patch_size=w["patch"],
               orthogonal_channel_emb_init=True, dropout_tokens_hcs="none", freeze_channel_emb=False, keep_rate=None,
               block_type="block", hcs_sampling="lowest_cosine_prob", hcs_sampling_temp=w["hcs_temp"],
               proxy_loss_lambda=w["l_cdl"], ortho_loss_v1_lambda=w["l_tdl"], drop_path_rate=0.0, gamma_s=w["gs"],
               gamma_d=w["gd"], reverse_pos_pairs=True, use_square=False, img_size=[w["img"]])


def set_seeds(seed: int, cuda: bool):
    """reference utils.py:394-401"""
    import numpy as np
    import torch

    random.seed(seed)
    np.random.seed(seed + 1)
    torch.manual_seed(seed + 2)
    if cuda:
        torch.cuda.manual_seed_all(seed + 4)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 - 0.05 or t > t1 + 0.05:
                continue
            f = [v.strip() for v in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# the CPU (reference) path: oracle restatement of the reference, torch fp32, all host threads
# ------------------------------------------------------------------------------------------------
def hbm_kernel_table(prof, subs, depth, D, patch, hbm_peak):
    """Memory-bound kernel classes: algorithmic bytes per STEP / summed CUDA-event time, against the
    measured HBM copy bandwidth.  `prof`: {tag: {"ms_per_step", "launches_per_step"}} of the built-in profiler for one
    full-channel step; `subs`: [(images, tokens per image)] of the step's sub-batches.  Full-size calls only are counted
    (the last block's CLS-row calls move KBs).  When the one-kernel patch embedding ran (no `im2col` launches) the row
    `embed_fused` replaces `im2col` + the embedding GEMM + the TDL token pass: fp32 image in, fp32 tokens + the bf16
    hi-patches (operand of the conv weight gradient) out; the `tdl` tag then only holds two KB-sized follow-up launches
    and is reported in us, not as a bandwidth."""
    by = {k: 0.0 for k in ("ln_fwd", "ln_bwd", "colsum", "attn_bwd_fin", "im2col", "tdl", "embed_bwd", "embed_fused")}
    for nb, Lc in subs:
        M = nb * Lc
        Tc = Lc - 1
        by["ln_fwd"] += (2 * (depth - 1) + 1) * M * D * (4 + 2)
        by["ln_bwd"] += (2 * (depth - 1) + 1) * M * D * 16
        by["colsum"] += nb * Tc * D * 2  # patch-embed bias gradient only: the block bias gradients come out of GEMM / attention epilogues
        by["attn_bwd_fin"] += depth * M * D * (4 + 2)
        by["im2col"] += nb * Tc * (patch ** 2) * (4 + 6)
        by["tdl"] += nb * Tc * D * 4
        by["embed_bwd"] += nb * (2 * Lc * D * 4 + Tc * D * 2) + nb * Lc * D * 4  # dY pass + batch sum of the token gradient
        by["embed_fused"] += nb * Tc * ((patch ** 2) * (4 + 2) + D * 4)
    fused = "embed_gemm" in prof and not prof.get("im2col", {}).get("ms_per_step", 0.0)
    src = {"embed_fused": "embed_gemm"}
    out = {}
    for k, nbytes in by.items():
        if k == "embed_fused" and not fused:
            continue
        if fused and k in ("im2col", "tdl"):
            continue
        pk = src.get(k, k)
        if pk in prof and prof[pk]["ms_per_step"] > 0:
            gbs = nbytes / (prof[pk]["ms_per_step"] * 1e-3) / 1e9
            out[k] = {"GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm_peak, 3),
                      "ms_per_step": round(prof[pk]["ms_per_step"], 4)}
    if fused and "tdl" in prof:
        out["tdl_followup"] = {"us_per_step": round(prof["tdl"]["ms_per_step"] * 1e3, 1), "bound": "latency (two KB-sized launches)"}
    return out


def cpu_reference_run(wname: str, steps: int, warmup: int, batch: int, seed: int = 2025):
    """Times fwd+loss+bwd(+AdamW) of the reference algorithm on the host.  Returns (img/s, seconds, cores, sample)."""
    import torch
    import torch.nn.functional as F

    from oracle import dichavit_oracle as O

    w = WORKLOADS[wname]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oc = O.OracleConfig(pretrained_model_name=w["size"], img_size=w["img"], patch_size=w["patch"],
                        in_channel_names=[f"c{i}" for i in range(w["channels"])], num_classes=w["classes"],
                        enable_sample=w["sample"], hcs_sampling="lowest_cosine_prob", hcs_sampling_temp=w["hcs_temp"],
                        proxy_loss_lambda=w["l_cdl"], ortho_loss_v1_lambda=w["l_tdl"], gamma_s=w["gs"], gamma_d=w["gd"],
                        reverse_pos_pairs=True, use_square=False, temperature=w.get("temperature", 0.11111))
    set_seeds(seed, False)
    chunks = w.get("chunks")
    has_head = not chunks
    weights = O.make_weights(oc, has_head, seed)
    params = {k: v.clone().requires_grad_(True) for k, v in weights.items() if k != "adaptive_interface.0"}
    opt = torch.optim.AdamW([p for p in params.values()], lr=4e-4, weight_decay=0.04)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, w["channels"], w["img"], w["img"], generator=g)
    y = torch.randint(0, w["classes"], (batch,), generator=g)
    # CHAMMI: the sample batch is split over the three chunks like the real one (trainer.py:846-931)
    parts = [(list(range(w["channels"])), 0, batch)]
    if chunks:
        parts, off = [], 0
        nb = max(1, batch // len(chunks))
        for chs, _ in chunks.values():
            parts.append((chs, off, nb))
            off += nb
        batch = off
    times = []
    for it in range(warmup + steps):
        if it == warmup:
            random.seed(seed)  # the C' sequence of the timed steps == the one our arm times (bench `timed(..., 2025)`)
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        for chs, off, nb in parts:
            o = O.forward(x[off:off + nb, :len(chs)], params, oc, chs, training=True, has_head=has_head)
            if has_head:
                loss = F.cross_entropy(o.out, y[off:off + nb]) + o.extra_loss * 1.0
            else:
                loss = O.proxy_loss(params["proxies"], o.out, y[off:off + nb], (1.0 / oc.temperature) ** 0.5) + o.extra_loss * 1.0
            loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return batch * len(times) / total, total, cores, f"{len(times)} steps of B={batch} (seeded DCS draws), after {warmup} warm-up"


def gpu_eager_run(wname: str, batch: int, steps: int, autocast: bool, seed: int = 2025):
    """The reference algorithm (oracle restatement = the reference's own ATen call sequence) run eagerly by PyTorch
    ON THE GPU: cuBLAS / cuDNN / ATen library kernels, the "existing Blackwell kernels" bar of SURVEY 8(d) (R0: fp32
    with TF32 off as the authors train; R1: torch.autocast(bfloat16), the reference's own use_amp path).  Full channels
    (no sampling) so that the work per step is fixed.  Returns images/s."""
    import torch
    import torch.nn.functional as F

    from oracle import dichavit_oracle as O

    w = WORKLOADS[wname]
    dev = torch.device("cuda")
    oc = O.OracleConfig(pretrained_model_name=w["size"], img_size=w["img"], patch_size=w["patch"],
                        in_channel_names=[f"c{i}" for i in range(w["channels"])], num_classes=w["classes"],
                        enable_sample=False, proxy_loss_lambda=w["l_cdl"], ortho_loss_v1_lambda=w["l_tdl"],
                        gamma_s=w["gs"], gamma_d=w["gd"], reverse_pos_pairs=True, use_square=False)
    weights = O.make_weights(oc, True, seed)
    params = {k: v.to(dev).requires_grad_(True) for k, v in weights.items() if k != "adaptive_interface.0"}
    opt = torch.optim.AdamW(list(params.values()), lr=4e-4, weight_decay=0.04, fused=True)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, w["channels"], w["img"], w["img"], generator=g).to(dev)
    y = torch.randint(0, w["classes"], (batch,), generator=g).to(dev)
    channels = list(range(w["channels"]))

    def one():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            o = O.forward(x, params, oc, channels, training=True, has_head=True)
            loss = F.cross_entropy(o.out.float(), y) + o.extra_loss.float()
        loss.backward()
        opt.step()

    for _ in range(2):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    del params, opt
    torch.cuda.empty_cache()
    return batch * steps / (e0.elapsed_time(e1) / 1000.0)


def bench_config(w, world: int) -> dict:
    """the `config` object of the JSON line: identical in both arms (the reference arm's bounded sample is described in
    its cpu_baseline.sample, not here)"""
    B = w["batch"]
    return {"workload": w["desc"], "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
            "step": "zero_grad + fwd + CE/extra loss + bwd (+NCCL grad all-reduce) + AdamW update",
            "dcs": "seeded random C' in 1..C per step (python random seed 2025 at the start of the timed steps), same "
                   "draws on every rank and in both arms",
            "l2": "each step streams >5 GB of activations (>> 126 MB L2); no explicit flush"}


def reference_arm(args, wname):
    """The reference's own CPU path (oracle port of the PyTorch module, fp32, every host thread): exactly --steps timed
    steps after --warmup untimed ones; each step is a bounded sample of the workload's batch (B = 4 at 224x224, 32 at
    32x32) so that the run ends within minutes.  The DCS draws of the timed steps are the ones our arm times."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[wname]
    batch = 4 if w["img"] >= 224 else 32
    if w["size"] == "base":
        batch = 2
    steps, warm = max(1, args.steps), max(0, args.warmup)
    ips, secs, cores, sample = cpu_reference_run(wname, steps, warm, batch)
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1000.0 * secs / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": bench_config(w, args.gpus),
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def algorithmic_flops_attn(B, L, D, bwd: bool):
    return (8.0 if bwd else 4.0) * L * L * D * B


def ours(args, wname):
    import ctypes

    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    from diverse_channel_vit_b200 import _lib
    from diverse_channel_vit_b200.dichavit import dichavit
    from diverse_channel_vit_b200.graphs import GraphedTrainStep
    from diverse_channel_vit_b200.optim import CosineLRSchedule, CosineWDSchedule, FusedAdamW
    from diverse_channel_vit_b200.trainer_glue import proxy_loss as _proxy_loss  # plain torch loss glue of the trainer

    w = WORKLOADS[wname]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass

    B = w["batch"]
    set_seeds(2025, True)
    chunks = w.get("chunks")
    mapper = {k: v[0] for k, v in chunks.items()} if chunks else {"train": list(range(w["channels"]))}
    model = dichavit(model_cfg(w), mapper=mapper).to(dev)
    model.train()
    model.direct_grad = True  # gradients land in .grad as views of one flat buffer, no per-tensor autograd nodes
    if world > 1:
        model.enable_data_parallel(overlap=os.environ.get("DCV_DP_OVERLAP", "1") != "0")

    # The public training-step API: one CUDA-graph launch per (chunk, batch, C') bucket (graphs.py).  DCV_GRAPHS=0
    # falls back to ordinary launches enqueued from Python (the round-1 path) for A/B comparison.
    use_graphs = os.environ.get("DCV_GRAPHS", "1") != "0"
    # AdamW as the reference configures it for JUMP-CP (configs/optimizer/adamw_jumpcp.yaml): lr 4e-4, weight decay
    # 0.04 -> 0.4 on a per-update cosine, cosine learning rate per epoch (configs/scheduler/cosine.yaml); all scalars
    # are evaluated on the device so that the captured step replays with the current values
    upe = 100
    opt = FusedAdamW(model, lr=4e-4, weight_decay=0.04, device_schedule=True, updates_per_epoch=upe,
                     lr_schedule=CosineLRSchedule(4e-4, 100, lr_min=1e-6, warmup_t=3, warmup_lr_init=1e-5, cycle_decay=0.5),
                     wd_schedule=CosineWDSchedule(0.04, 0.4, 100, upe))
    gen = torch.Generator(device="cpu").manual_seed(2025 + rank)
    max_ch = max(len(v[0]) for v in chunks.values()) if chunks else w["channels"]
    x_host = torch.randn(B, max_ch, w["img"], w["img"], generator=gen).pin_memory()
    y_host = torch.randint(0, w["classes"], (B,), generator=gen).pin_memory()
    x_dev = x_host.to(dev)
    y_dev = y_host.to(dev)
    copy_stream = torch.cuda.Stream(device=dev)
    x_buf = [torch.empty_like(x_dev) for _ in range(2)]
    y_buf = [torch.empty_like(y_dev) for _ in range(2)]

    def loss_glue(m, out, extra, y):
        if chunks:
            return _proxy_loss(m.proxies, out, y, m.scale) + extra * 1.0  # trainer.py:912-914
        return F.cross_entropy(out, y) + extra * 1.0  # trainer.py:986-995

    gstep = GraphedTrainStep(model, opt, loss_fn=loss_glue)
    mode = {"eager": not use_graphs}

    def step(x, y):
        if chunks:  # trainer.py:846-931: one forward/backward per chunk, one optimiser step
            off, names = 0, list(chunks)
            for i, name in enumerate(names):
                chs, nb = chunks[name]
                xs, ys = gstep.input_buffers(name, (nb, len(chs), w["img"], w["img"]), x.dtype, dev)
                xs.copy_(x[off:off + nb, :len(chs)])
                ys.copy_(y[off:off + nb])
                loss = gstep(xs, ys, name, last=i == len(names) - 1, eager=mode["eager"])
                off += nb
            return loss
        return gstep(x, y, "train", eager=mode["eager"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, e2e: bool, seed: int):
        """returns max-over-ranks elapsed ms (CUDA events)"""
        random.seed(seed)  # DCS draws: identical on every rank and in every pass
        torch.manual_seed(seed + 2)
        torch.cuda.manual_seed_all(seed + 4)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if e2e:
            # the input pipeline a trainer would run: the H2D copy of step i+1 (pinned host -> device, on a copy
            # stream, double-buffered) overlaps the compute of step i; every step's copy is inside the timed region
            main = torch.cuda.current_stream()
            ready = [torch.cuda.Event(), torch.cuda.Event()]
            freed = [torch.cuda.Event(), torch.cuda.Event()]

            def upload(i):
                sl = i & 1
                copy_stream.wait_event(freed[sl])
                with torch.cuda.stream(copy_stream):
                    x_buf[sl].copy_(x_host, non_blocking=True)
                    y_buf[sl].copy_(y_host, non_blocking=True)
                    ready[sl].record(copy_stream)

            for sl in range(2):
                freed[sl].record(main)
            upload(0)
            for i in range(nsteps):
                sl = i & 1
                main.wait_event(ready[sl])
                if i + 1 < nsteps:
                    upload(i + 1)
                loss = step(x_buf[sl], y_buf[sl])
                freed[sl].record(main)
                loss.item()  # D2H read of the step's result
        else:
            for _ in range(nsteps):
                step(x_dev, y_dev)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return ms.item()

    def launches():
        return _lib.launch_count() + gstep.kernel_launches

    K, W = args.steps, args.warmup
    pe = model.feature_extractor.patch_embed
    # warm-up: one full-channel step first (largest shape: activation / workspace arenas reach their final size,
    # every kernel attribute is set), every (chunk, C') graph captured, then W >= 3 steps of the workload itself
    pe.enable_sample = False
    timed(1, False, 1)
    pe.enable_sample = w["sample"]
    t_cap = time.time()
    n_graphs = 0
    if use_graphs:
        if chunks:
            names = list(chunks)
            off = 0
            for i, name in enumerate(names):
                chs, nb = chunks[name]
                n_graphs += gstep.precapture(x_dev[off:off + nb, :len(chs)].contiguous(), y_dev[off:off + nb], name,
                                             first=i == 0, last=i == len(names) - 1)
                off += nb
        else:
            n_graphs += gstep.precapture(x_dev, y_dev, "train")
    t_cap = time.time() - t_cap
    timed(max(W, 3), False, 1)
    clk = ClockSampler(local) if rank == 0 else None
    l0, g0 = launches(), gstep.graph_launches
    t0 = time.time()
    ms = timed(K, False, 2025)
    t1 = time.time()
    n_launch, n_glaunch = launches() - l0, gstep.graph_launches - g0
    clocks = clk.stop(t0, t1) if clk else None
    value = world * B * K / (ms / 1000.0)

    timed(2, True, 1)
    ms_e2e = timed(K, True, 2025)
    e2e_value = world * B * K / (ms_e2e / 1000.0)

    # sustained pass: the same workload for >= 3 s (the --steps pass is a 0.2 s burst at boost clocks)
    ks = max(K, int(3000.0 / (ms / K)) + 1)
    clk2 = ClockSampler(local) if rank == 0 else None
    t0 = time.time()
    ms_sus = timed(ks, False, 2025)
    t1 = time.time()
    clocks_sus = clk2.stop(t0, t1) if clk2 else None
    sustained = {"value": world * B * ks / (ms_sus / 1000.0), "unit": UNIT, "steps": ks, "ms_per_step": ms_sus / ks,
                 "seconds": ms_sus / 1000.0, "clocks": clocks_sus}

    # full-channel (C' = C, no sampling) reference point
    pe.enable_sample = False
    timed(2, False, 1)
    kf = max(3, K // 2)
    ms_full = timed(kf, False, 2025)
    full_value = world * B * kf / (ms_full / 1000.0)

    # ---- per-kernel-class breakdown, live CUDA events around every launch (the built-in profiler sits in the host
    # launchers, so these passes run the identical step sequence as ordinary launches, not as a graph replay) ----
    lib = _lib.lib()
    ntags = lib.dcv_profile_num_tags()
    lib.dcv_profile_tag_name.restype = ctypes.c_char_p
    names_t = [lib.dcv_profile_tag_name(i).decode() for i in range(ntags)]
    msb = (ctypes.c_double * ntags)()
    cnt = (ctypes.c_longlong * ntags)()
    kp = max(2, min(K, 5))
    mode["eager"] = True
    timed(1, False, 1)
    barrier()
    lib.dcv_profile_start()
    ms_prof = timed(kp, False, 2025)  # still full channels: L fixed, algorithmic work per launch exact
    lib.dcv_profile_stop(msb, cnt, ntags)
    prof = {names_t[i]: {"ms_per_step": msb[i] / kp, "launches_per_step": cnt[i] / kp} for i in range(ntags) if cnt[i]}
    pe.enable_sample = w["sample"]
    # same breakdown over the timed workload itself (seeded DCS draws, variable L)
    barrier()
    lib.dcv_profile_start()
    timed(K, False, 2025)
    lib.dcv_profile_stop(msb, cnt, ntags)
    prof_dcs = {names_t[i]: round(msb[i] / K, 4) for i in range(ntags) if cnt[i]}
    # eager (no graphs) end-to-end pass of the same workload: what the graph launch buys
    ms_e2e_eager = timed(K, True, 2025)
    mode["eager"] = not use_graphs

    D = model.dim
    npatch = (w["img"] // w["patch"]) ** 2
    L = 1 + (max_ch if chunks else w["channels"]) * npatch
    depth = len(model.feature_extractor.blocks)
    heads = model.feature_extractor.num_heads
    # sub-batches of one (full-channel) step: (images, tokens per image)
    subs = [(nb, 1 + len(chs) * npatch) for chs, nb in chunks.values()] if chunks else [(B, L)]
    # algorithmic FLOPs per STEP and kernel class for the work actually performed (the last block only
    # runs one query tile of attention and CLS-row proj / MLP; recomputation of S in the backward is not counted)
    fl = {"attn_fwd": 0.0, "attn_bwd": 0.0, "gemm_nt": 0.0}
    for nb, Lc in subs:
        att = 4.0 * Lc * Lc * D * nb * ((depth - 1) + min(1.0, 128.0 / Lc))
        fl["attn_fwd"] += att
        fl["attn_bwd"] += 2.0 * att
        fl["gemm_nt"] += (depth - 1) * 2.0 * nb * Lc * D * 12 * D + 2.0 * nb * Lc * D * 3 * D + 2.0 * nb * D * 9 * D
    fl["gemm_nn"] = fl["gemm_nt"]  # dgrad
    fl["gemm_tn"] = fl["gemm_nt"]  # wgrad
    top = max(prof.items(), key=lambda kv: kv[1]["ms_per_step"])[0] if prof else None
    roof = None
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json, sustained)" if peaks else "fallback"
    traffic = None
    try:
        tj = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        if top in tj and "B" in tj[top] and tj[top].get("L") == L and not chunks:
            traffic = tj[top]["dram_bytes"] * B / tj[top]["B"]  # ncu capture at another batch size, linear in B
    except Exception:
        pass
    if top in fl:
        ms_top = prof[top]["ms_per_step"]
        ach = fl[top] / (ms_top * 1e-3) / 1e12
        roof = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": ach / peak_tf, "traffic": traffic, "peak_source": peak_src,
                "avg_launch_ms": ms_top / prof[top]["launches_per_step"],
                "launches_per_step": prof[top]["launches_per_step"], "share_of_step": ms_top / (ms_prof / kp),
                "note": "full-channel pass; achieved = algorithmic FLOPs of all launches of this kernel class in one step / "
                        "their summed CUDA-event time; work actually performed (last block: one query tile, CLS-row "
                        "proj/MLP); recompute not counted"}
    elif top is not None:
        roof = {"kernel": top, "bound": "hbm", "achieved": None, "peak": float(peaks.get("hbm_gbs", 6650.0)),
                "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src}
    # memory-bound kernel classes: algorithmic bytes per STEP / summed CUDA-event time, against the
    # measured HBM copy bandwidth.  Full-size calls only are counted (the last block's CLS-row calls move KBs).
    hbm_peak = float(peaks.get("hbm_gbs", 6550.0))
    hbm_kernels = hbm_kernel_table(prof, subs, depth, D, w["patch"], hbm_peak)
    attn_ms = sum(prof.get(k, {}).get("ms_per_step", 0.0) for k in ("attn_fwd", "attn_bwd", "attn_bwd_prep", "attn_bwd_fin"))
    attn_tf = (fl["attn_fwd"] + fl["attn_bwd"]) / (attn_ms * 1e-3) / 1e12 if attn_ms else None
    gemm_ms = sum(prof.get(k, {}).get("ms_per_step", 0.0) for k in ("gemm_nt", "gemm_nn", "gemm_tn"))
    gemm_tf = 3.0 * fl["gemm_nt"] / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None

    if rank != 0:
        _leave(world)
        return

    # ---- attention kernels next to the library kernel on the same box (torch SDPA, cuDNN backend), at this
    # workload's full-channel shape: CUDA events, 20 launches each after 5 warm-ups ----
    sdpa = None
    if world == 1 and not args.no_eager:
        try:
            sdpa = attention_vs_sdpa(subs[-1][0], subs[-1][1], heads)
        except Exception as ex:
            sdpa = {"error": repr(ex)[:200]}

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0, N == 1 only ----
    cpu = None
    if world == 1 and not args.no_cpu:
        bb = 4 if w["img"] >= 224 else 32
        ips, secs, cores, sample = cpu_reference_run(wname, 12, 1, bb)  # ~10-15 s of host work
        cpu = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    # ---- PyTorch eager on the same GPU (library kernels), rank 0, N == 1 only ----
    eager = None
    if world == 1 and not args.no_eager:
        try:
            del gstep, model, opt
            torch.cuda.empty_cache()
            eb = B
            while True:  # the reference materialises [B,H,L,L] fp32 probabilities per layer: halve the batch on OOM
                try:
                    r32, r16 = gpu_eager_run(wname, eb, 3, False), gpu_eager_run(wname, eb, 3, True)
                    break
                except torch.cuda.OutOfMemoryError:
                    torch.cuda.empty_cache()
                    if eb <= 4:
                        raise
                    eb //= 2
            eager = {"what": "the reference algorithm run eagerly by PyTorch on this GPU (cuBLAS/cuDNN/ATen), full channels; "
                             "compare with full_channels.value",
                     "source": "port (oracle restatement of the reference's ATen call sequence: the reference checkout and "
                               "its timm / omegaconf dependencies do not exist on the GPU box)",
                     "batch": eb, "fp32_images_per_s": r32, "bf16_autocast_images_per_s": r16}
        except Exception as ex:  # report, do not fail the bench
            eager = {"error": repr(ex)[:200]}

    h2d = x_host.numel() * 4 + y_host.numel() * 8
    cfg = bench_config(w, world)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": cfg,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / K, "without_graphs_ms_per_step": ms_e2e_eager / K},
        "gpu_launches": int(n_launch),
        "graph_launches": int(n_glaunch),
        "graphs": {"enabled": use_graphs, "captured": n_graphs, "capture_seconds": round(t_cap, 2),
                   "note": "one cudaGraphLaunch per (chunk, C') bucket and micro-step; gpu_launches counts the kernels "
                           "of this library executed by those replays"},
        "clocks": clocks,
        "sustained": sustained,
        "full_channels": {"value": full_value, "unit": UNIT, "ms_per_step": ms_full / kf, "tokens": L,
                          "model_tflops": None},
        "roofline": roof,
        "kernel_breakdown_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms_per_step"])},
        "kernel_breakdown_dcs_ms_per_step": dict(sorted(prof_dcs.items(), key=lambda kv: -kv[1])),
        "attn_tflops": attn_tf, "attn_frac_of_peak": (attn_tf / peak_tf) if attn_tf else None,
        "gemm_tflops": gemm_tf, "gemm_frac_of_peak": (gemm_tf / peak_tf) if gemm_tf else None,
        "hbm_kernels": hbm_kernels,
        "attention_vs_sdpa": sdpa,
        "cpu_baseline": cpu,
        "torch_eager_gpu": eager,
    }
    # algorithmic model FLOPs of the full-channel step: fwd = 2 T P^2 D + depth (24 L D^2 + 4 L^2 D) + 2 D cls
    fwd = 0.0
    for nb, Lc in subs:
        fwd += nb * (2.0 * (Lc - 1) * w["patch"] ** 2 * D + depth * (24.0 * Lc * D * D + 4.0 * Lc * Lc * D) + 2.0 * D * w["classes"])
    line["full_channels"]["model_tflops"] = 3.0 * fwd * world / (ms_full / kf * 1e-3) / 1e12
    print(json.dumps(line), flush=True)
    _leave(world)


def _leave(world: int) -> None:
    """End of a multi-rank run.  The NCCL communicator is NOT torn down: destroy_process_group() blocks forever while
    CUDA graphs that captured collectives of that communicator are alive (seen at N = 2: the JSON line was out, the
    process never exited), and there is nothing to save -- flush and leave."""
    if world > 1:
        import torch

        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def attention_vs_sdpa(B: int, L: int, H: int):
    """dcv_attn_fwd / dcv_attn_bwd (prep + main + finish) against torch.nn.functional.scaled_dot_product_attention with
    the cuDNN backend (the library kernel for this op on Blackwell), forward and backward, same B / L / H, bf16."""
    import torch
    from torch.nn.attention import SDPBackend, sdpa_kernel

    from diverse_channel_vit_b200 import kernels as Kn

    dev = torch.device("cuda")
    D = H * 64
    qkv = torch.randn(B * L, 3 * D, device=dev).bfloat16()
    do = torch.randn(B * L, D, device=dev).bfloat16()

    def bench(fn, n=20, warm=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    o = torch.empty(B * L, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, Kn.lpad(L), device=dev)
    dqkv = torch.empty_like(qkv)
    delta = Kn.delta_ws(B, H, L, dev)
    acc = torch.empty(B, H, L, 64, device=dev)
    ours_f = bench(lambda: Kn.attn_fwd(qkv, B, L, H, o=o, lse2=lse))
    ours_b = bench(lambda: Kn.attn_bwd(qkv, o, do, lse, B, L, H, dqkv=dqkv, delta=delta, dq_acc=acc))
    q, k, v = (t.contiguous().requires_grad_(True) for t in qkv.reshape(B, L, 3, H, 64).permute(2, 0, 3, 1, 4))
    with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
        lib_f = bench(lambda: torch.nn.functional.scaled_dot_product_attention(q.detach(), k.detach(), v.detach()))
        out = torch.nn.functional.scaled_dot_product_attention(q, k, v)
        gout = torch.randn_like(out)
        lib_b = bench(lambda: torch.autograd.grad(out, (q, k, v), gout, retain_graph=True))
    ff, fb = 4.0 * B * H * L * L * 64, 8.0 * B * H * L * L * 64
    return {"shape": {"B": B, "L": L, "H": H, "head_dim": 64},
            "fwd_us": {"ours": ours_f, "sdpa_cudnn": lib_f}, "bwd_us": {"ours_prep_main_finish": ours_b, "sdpa_cudnn": lib_b},
            "fwd_tflops": {"ours": ff / ours_f / 1e6, "sdpa_cudnn": ff / lib_f / 1e6},
            "bwd_tflops": {"ours": fb / ours_b / 1e6, "sdpa_cudnn": fb / lib_b / 1e6}}


def eval_mode(args, wname):
    """`--mode eval`: inference forward of the drop-in module (model.eval(), torch.inference_mode; reference
    trainer.py:385-472), all channels, BASELINE.json configs[4] "eval".  Same JSON contract, metric = eval images/sec."""
    import torch
    import torch.distributed as dist

    from diverse_channel_vit_b200 import _lib
    from diverse_channel_vit_b200.dichavit import dichavit

    w = WORKLOADS[wname]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    set_seeds(2025, True)
    chunks = w.get("chunks")
    if chunks:
        raise SystemExit("--mode eval: use jumpcp / so2sat / vitb")
    model = dichavit(model_cfg(w), mapper={"train": list(range(w["channels"]))}).to(dev).eval()
    B = w["batch"] * 4  # inference keeps no activations: a larger batch fills the GPU
    gen = torch.Generator(device="cpu").manual_seed(2025 + rank)
    x_host = torch.randn(B, w["channels"], w["img"], w["img"], generator=gen).pin_memory()
    x_dev = x_host.to(dev)
    x_buf = [torch.empty_like(x_dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.inference_mode():
            if e2e:
                main = torch.cuda.current_stream()
                ready = [torch.cuda.Event(), torch.cuda.Event()]
                freed = [torch.cuda.Event(), torch.cuda.Event()]
                for sl in range(2):
                    freed[sl].record(main)

                def upload(i):
                    sl = i & 1
                    copy_stream.wait_event(freed[sl])
                    with torch.cuda.stream(copy_stream):
                        x_buf[sl].copy_(x_host, non_blocking=True)
                        ready[sl].record(copy_stream)

                upload(0)
                for i in range(n):
                    sl = i & 1
                    main.wait_event(ready[sl])
                    if i + 1 < n:
                        upload(i + 1)
                    out = model(x_buf[sl], "train")
                    freed[sl].record(main)
                    out.argmax(dim=1).cpu()  # D2H of the predictions
            else:
                for _ in range(n):
                    model(x_dev, "train")
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    timed(W, False)
    clk = ClockSampler(local) if rank == 0 else None
    l0 = _lib.launch_count()
    t0 = time.time()
    ms = timed(K, False)
    t1 = time.time()
    n_launch = _lib.launch_count() - l0
    clocks = clk.stop(t0, t1) if clk else None
    timed(2, True)
    ms_e2e = timed(K, True)
    if rank == 0:
        D = model.dim
        npatch = (w["img"] // w["patch"]) ** 2
        L = 1 + w["channels"] * npatch
        depth = len(model.feature_extractor.blocks)
        fwd = B * (2.0 * (L - 1) * w["patch"] ** 2 * D + depth * (24.0 * L * D * D + 4.0 * L * L * D) + 2.0 * D * w["classes"])
        line = {"metric": "eval images/sec (fwd)", "value": world * B * K / (ms / 1e3), "unit": UNIT, "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": w["desc"] + " -- EVAL (inference forward, all channels)", "per_gpu_batch": B,
                           "global_batch": B * world, "parallelism": f"dp{world}"},
                "e2e": {"value": world * B * K / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
                        "d2h_bytes_per_step": B * 8, "ms_per_step": ms_e2e / K},
                "gpu_launches": int(n_launch), "clocks": clocks,
                "model_tflops": fwd * world / (ms / K * 1e-3) / 1e12}
        print(json.dumps(line), flush=True)
    _leave(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="jumpcp", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eager", action="store_true", help="skip the PyTorch-eager-on-GPU / SDPA comparison legs")
    ap.add_argument("--mode", default="train", choices=["train", "eval"], help="eval: inference forward (configs[4])")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args, args.workload)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.mode == "eval":
        eval_mode(args, args.workload)
    else:
        ours(args, args.workload)


if __name__ == "__main__":
    main()
