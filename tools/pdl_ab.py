"""GPU: CUDA-graph training step with and without programmatic dependent launch, captured into two sets of graphs in
ONE process and timed interleaved (run-to-run and box-to-box variation is ~3 %, the effect being measured too).
Usage: python tools/pdl_ab.py"""
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402
from diverse_channel_vit_b200 import _lib  # noqa: E402
from diverse_channel_vit_b200.dichavit import dichavit  # noqa: E402
from diverse_channel_vit_b200.graphs import GraphedTrainStep  # noqa: E402
from diverse_channel_vit_b200.optim import FusedAdamW  # noqa: E402

lib = _lib.lib()
w = bench.WORKLOADS["jumpcp"]
x = torch.randn(32, 8, 224, 224, device="cuda")
y = torch.randint(0, 161, (32,), device="cuda")
forced = {"c": 8}
CS = (1, 2, 4, 6, 8)


def make(pdl):
    lib.dcv_debug_set_pdl(pdl)
    bench.set_seeds(2025, True)
    m = dichavit(bench.model_cfg(w), mapper={"train": list(range(8))}).cuda().train()
    opt = FusedAdamW(m, lr=4e-4, weight_decay=0.04, device_schedule=True)
    step = GraphedTrainStep(m, opt)
    pe = m.feature_extractor.patch_embed
    orig = pe.draw_host

    def draw(chunk_name, n_in):
        d = orig(chunk_name, n_in)
        d["c_new"] = forced["c"]
        return d

    pe.draw_host = draw
    for cs in CS:  # capture every bucket under this PDL setting
        forced["c"] = cs
        for _ in range(3):
            step(x, y, "train").item()
    torch.cuda.synchronize()
    return step


steps = {0: make(0), 1: make(1)}


def timed(step, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step(x, y, "train")
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"{'C_sel':>5s} {'plain ms/step':>14s} {'PDL ms/step':>12s} {'ratio':>7s}")
for cs in CS:
    forced["c"] = cs
    acc = {0: [], 1: []}
    for rep in range(4):
        for k in (0, 1):
            timed(steps[k], 2)
            acc[k].append(timed(steps[k]))
    a, b = min(acc[0]), min(acc[1])
    print(f"{cs:5d} {a:14.3f} {b:12.3f} {b / a:7.3f}   (all: plain {[round(v, 3) for v in acc[0]]} pdl {[round(v, 3) for v in acc[1]]})", flush=True)
