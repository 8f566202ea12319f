import sys, cProfile, pstats, io
sys.path.insert(0, ".")
import torch, torch.nn.functional as F
import bench
from diverse_channel_vit_b200.dichavit import dichavit
from diverse_channel_vit_b200.optim import FusedAdamW
w = bench.WORKLOADS["jumpcp"]
bench.set_seeds(2025, True)
m = dichavit(bench.model_cfg(w), mapper={"train": list(range(8))}).cuda().train()
opt = FusedAdamW(m, lr=4e-4, weight_decay=0.04)
pe = m.feature_extractor.patch_embed
x = torch.randn(32, 8, 224, 224, device="cuda"); y = torch.randint(0, 161, (32,), device="cuda")
chan = pe.chunk_channels("train", x.device)
cs = 1
it = torch.arange(cs, dtype=torch.int32, device="cuda")
pe.select_channels = lambda *_a, **_k: (cs, it, chan[it.long()].to(torch.int32))
def step():
    opt.zero_grad(); out, extra = m(x, "train"); (F.cross_entropy(out, y) + extra).backward(); opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:5000])
