"""Raw NCCL all-reduce time for the flat gradient (torchrun --nproc-per-node N tools/nccl_probe.py)."""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
for n in (21_600_000, 5_400_000, 1_000_000):
    g = torch.randn(n, device="cuda")
    for _ in range(5): dist.all_reduce(g)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dist.all_reduce(g)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        ms = e0.elapsed_time(e1) / 20
        print(f"world {dist.get_world_size()} all_reduce {n * 4 / 1e6:.1f} MB fp32: {ms * 1e3:.1f} us  ({n * 4 / ms / 1e6:.1f} GB/s algbw)", flush=True)
dist.destroy_process_group()
