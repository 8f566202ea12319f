"""GPU: CUDA-graph training step with the block backward's weight-gradient GEMMs / accumulator clears on the caller's
stream (0) and on the library's side stream = a parallel branch of the graph (1), captured into two sets of graphs in
ONE process and timed interleaved, per sampled channel count C' (JUMP-CP shape, B = 32).
Usage: python tools/overlap_ab.py > profiles/r2_bwd_overlap_ab.txt"""
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
CS = (1, 2, 3, 4, 5, 6, 7, 8)


def make(overlap):
    lib.dcv_debug_set_bwd_overlap(overlap)
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
    for cs in CS:  # capture every bucket under this setting
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


lib.dcv_debug_set_bwd_overlap(-1)
print(f"{'C_sel':>5s} {'serial ms/step':>14s} {'branch ms/step':>14s} {'ratio':>7s}")
tot = {0: 0.0, 1: 0.0}
for cs in CS:
    forced["c"] = cs
    acc = {0: [], 1: []}
    for rep in range(3):
        for k in (0, 1):
            timed(steps[k], 2)
            acc[k].append(timed(steps[k]))
    a, b = min(acc[0]), min(acc[1])
    tot[0] += a
    tot[1] += b
    print(f"{cs:5d} {a:14.3f} {b:14.3f} {b / a:7.3f}   (all: serial {[round(v, 3) for v in acc[0]]} branch {[round(v, 3) for v in acc[1]]})", flush=True)
print(f"sum over C' = 1..8 (the uniform DCS mix): serial {tot[0]:.3f} ms, branch {tot[1]:.3f} ms, ratio {tot[1] / tot[0]:.3f}")
