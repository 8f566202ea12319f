import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, ".")
import torch, bench
from diverse_channel_vit_b200.dichavit import dichavit
w = bench.WORKLOADS["so2sat"]
m = dichavit(bench.model_cfg(w), mapper={"train": list(range(18))}).cuda().train()
m.feature_extractor.patch_embed.enable_sample = False
x = torch.randn(8, 18, 32, 32, device="cuda")
for cs in (18,):
    out, extra = m(x, "train")
    torch.cuda.synchronize()
    print("ok", out.shape, float(extra))
