"""Step time vs sum of kernel time for each sampled channel count C' (JUMP-CP shape): shows where the host
(Python + ~290 launches) rather than the GPU bounds the step."""
import sys, ctypes, time
sys.path.insert(0, ".")
import torch, torch.nn.functional as F
import bench
from diverse_channel_vit_b200 import _lib
from diverse_channel_vit_b200.dichavit import dichavit
from diverse_channel_vit_b200.optim import FusedAdamW
w = bench.WORKLOADS["jumpcp"]
bench.set_seeds(2025, True)
m = dichavit(bench.model_cfg(w), mapper={"train": list(range(8))}).cuda().train()
opt = FusedAdamW(m, lr=4e-4, weight_decay=0.04)
m.direct_grad = True
pe = m.feature_extractor.patch_embed
x = torch.randn(32, 8, 224, 224, device="cuda"); y = torch.randint(0, 161, (32,), device="cuda")
chan = pe.chunk_channels("train", x.device)
lib = _lib.lib(); nt = lib.dcv_profile_num_tags(); lib.dcv_profile_tag_name.restype = ctypes.c_char_p
def step():
    opt.zero_grad(); out, extra = m(x, "train"); (F.cross_entropy(out, y) + extra).backward(); opt.step()
for cs in (8, 1, 2, 4):
    it = torch.arange(cs, dtype=torch.int32, device="cuda")
    pe.select_channels = lambda *_a, **_k: (cs, it, chan[it.long()].to(torch.int32))
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10): step()
    e1.record(); t_host = (time.perf_counter() - t0) / 10 * 1e3
    torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 10
    msb = (ctypes.c_double * nt)(); cnt = (ctypes.c_longlong * nt)()
    lib.dcv_profile_start()
    for _ in range(5): step()
    lib.dcv_profile_stop(msb, cnt, nt)
    names = [lib.dcv_profile_tag_name(i) for i in range(nt)]
    if cs <= 2:
        print("   ", {(n.decode() if isinstance(n, bytes) else n): round(msb[i] / 5, 3) for i, n in enumerate(names) if cnt[i]})
    print(f"C'={cs}: step {ms:6.2f} ms (host enqueue {t_host:5.2f} ms), kernel sum {sum(msb)/5:6.2f} ms, img/s {32/ms*1e3:7.0f}", flush=True)
