"""GPU: the CUDA-graph training step per sampled channel count C' (JUMP-CP shape, B = 32): device time per step, the
host's wall-clock per step (graph launch + the loss read), and the eager launch sequence next to it -- shows that no
C' bucket is host-bound any more (round 1: 2.7 ms of Python + ~290 launches against 1.5 ms of kernels at C' = 1).
Usage: python tools/per_cprime_graphs.py > profiles/r2_per_cprime.txt"""
import random
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402
from diverse_channel_vit_b200.dichavit import dichavit  # noqa: E402
from diverse_channel_vit_b200.graphs import GraphedTrainStep  # noqa: E402
from diverse_channel_vit_b200.optim import FusedAdamW  # noqa: E402

w = bench.WORKLOADS["jumpcp"]
bench.set_seeds(2025, True)
m = dichavit(bench.model_cfg(w), mapper={"train": list(range(8))}).cuda().train()
opt = FusedAdamW(m, lr=4e-4, weight_decay=0.04, device_schedule=True)
step = GraphedTrainStep(m, opt)
pe = m.feature_extractor.patch_embed
x = torch.randn(32, 8, 224, 224, device="cuda")
y = torch.randint(0, 161, (32,), device="cuda")
orig = pe.draw_host
forced = {"c": 8}


def draw(chunk_name, n_in):
    d = orig(chunk_name, n_in)  # consumes the python RNG exactly like an unforced step
    d["c_new"] = forced["c"]
    return d


pe.draw_host = draw
print(f"{'C_sel':>5s} {'L':>5s} {'graph ms/step (device)':>24s} {'graph host wall ms/step':>24s} {'eager ms/step (device)':>24s} "
      f"{'eager host wall ms/step':>24s} {'img/s (graph)':>14s}")
for cs in range(1, 9):
    forced["c"] = cs
    row = []
    for eager in (False, True):
        for _ in range(3):
            step(x, y, "train", eager=eager).item()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            step(x, y, "train", eager=eager).item()  # the per-step loss read of a training loop
        e1.record()
        torch.cuda.synchronize()
        row += [e0.elapsed_time(e1) / n, (time.perf_counter() - t0) / n * 1e3]
    print(f"{cs:5d} {1 + 196 * cs:5d} {row[0]:24.3f} {row[1]:24.3f} {row[2]:24.3f} {row[3]:24.3f} {32 / row[0] * 1e3:14.0f}", flush=True)
