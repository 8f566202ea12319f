"""GPU: clock64 timeline of one CTA of the fused patch-embedding kernel (csrc/embed_fused.cu) at the JUMP-CP shape
(B = 32, 8 x 224 x 224, all channels): where a tile's ~50 k cycles go.  Prints cycles relative to the CTA's entry."""
import ctypes
import os
import sys

sys.path.insert(0, ".")
import torch

import bench
from diverse_channel_vit_b200 import _lib
from diverse_channel_vit_b200.dichavit import dichavit

w = bench.WORKLOADS["jumpcp"]
bench.set_seeds(2025, True)
m = dichavit(bench.model_cfg(w), mapper={"train": list(range(8))}).cuda().train()
m.feature_extractor.patch_embed.enable_sample = False
x = torch.randn(32, 8, 224, 224, device="cuda")
lib = _lib.lib()
MODE = int(os.environ.get("EF_MODE", "1"))
lib.dcv_debug_set_embed_fused(MODE)
with torch.no_grad():
    for _ in range(2):
        m(x, "train")
    torch.cuda.synchronize()
    buf = torch.zeros(512, dtype=torch.int64, device="cuda")
    _lib.check(lib.dcv_debug_embed_timeline(ctypes.c_void_p(buf.data_ptr())))
    m(x, "train")
    torch.cuda.synchronize()
    _lib.check(lib.dcv_debug_embed_timeline(None))
t = buf.cpu().tolist()
if MODE == 2:
    # persistent kernel, CTA 5: slot = role * 128 + 16 * tile iteration + point
    t0 = t[511]
    r = lambda i: t[i] - t0 if t[i] else None
    for j in range(4):
        print(f"tile {j}: MMA  wait-acc_empty {r(16 * j)} -> {r(16 * j + 1)} | per kc (a_full seen, 48 MMAs issued): "
              f"{[(r(16 * j + 2 + 2 * k), r(16 * j + 3 + 2 * k)) for k in range(4)]}")
        print(f"tile {j}: conv per kc (stage_full seen, a_empty seen, converted): "
              f"{[(r(128 + 4 * (4 * j + k)), r(128 + 4 * (4 * j + k) + 1), r(128 + 4 * (4 * j + k) + 2)) for k in range(4)]}")
        print(f"tile {j}: epi  addend requested {r(256 + 16 * j)} | acc_full seen {r(256 + 16 * j + 1)} | pass 1 done {r(256 + 16 * j + 2)} "
              f"| row norms exchanged {r(256 + 16 * j + 3)} | pass 2 done {r(256 + 16 * j + 4)}")
    sys.exit(0)
t0 = t[70]


def rel(i):
    return t[i] - t0 if t[i] else None


print("entry 0 | setup done", rel(71), "| pdl_wait done", rel(72), "| exit", rel(73))
print("image TMA issue (kc 0-3):", [rel(i) for i in range(4)])
print("weight TMA issue (it 0-7):", [rel(8 + i) for i in range(8)])
print("conv warp 4, per kc (stage_full seen, a_empty seen, converted):", [[rel(32 + 3 * k + j) for j in range(3)] for k in range(4)])
print("MMA: a_full seen (kc 0-3):", [rel(88 + k) for k in range(4)])
print("MMA: operands ready (it 0-7):", [rel(16 + i) for i in range(8)])
print("MMA: issued (it 0-7):       ", [rel(24 + i) for i in range(8)])
for name, tb in (("warp 4", 48), ("warp 8", 96)):
    print(f"epilogue {name}: acc_full seen {rel(tb)}, bias barrier {rel(tb + 1)}, pass-1 chunks {[rel(tb + 2 + i) for i in range(6)]}, "
          f"ss barrier {rel(tb + 8)}, pass-2 chunks {[rel(tb + 9 + i) for i in range(6)]}")
