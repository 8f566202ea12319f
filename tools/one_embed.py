"""Minimal driver for ncu: a few forward passes of the drop-in module at the JUMP-CP shape (B = 32, 8 x 224 x 224), so
that the patch-embedding kernels can be captured (ncu -k regex:embed_fused_kernel ...)."""
import sys
sys.path.insert(0, ".")
import torch
import bench
from diverse_channel_vit_b200.dichavit import dichavit
w = bench.WORKLOADS["jumpcp"]
bench.set_seeds(2025, True)
m = dichavit(bench.model_cfg(w), mapper={"train": list(range(8))}).cuda().train()
m.feature_extractor.patch_embed.enable_sample = False
x = torch.randn(32, 8, 224, 224, device="cuda")
with torch.no_grad():
    for _ in range(3):
        out = m(x, "train")
torch.cuda.synchronize()
print("ok")
