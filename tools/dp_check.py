"""torchrun --nproc-per-node 2 tools/dp_check.py : data-parallel gradients (NCCL, overlapped buckets) vs the same
module run on the concatenated batch on one GPU."""
import os, random, sys
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from tests.util import O, build_cuda_model, cases, cuda_step, make_inputs, rel_l2

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
oc, mapper, chunk, has_head, B, wseed, iseed, xlam = cases()["tiny_jumpcp"]
weights = O.make_weights(oc, has_head, wseed)
x, y = make_inputs(oc, 4 * world, 8, oc.num_classes, iseed)
idx = [6, 1, 4, 2]
model = build_cuda_model(oc, mapper, weights, f"cuda:{local}")
model.enable_data_parallel()
sl = slice(rank * 4, rank * 4 + 4)
_, _, _, g_dp = cuda_step(model, x[sl].cuda(), y[sl].cuda(), chunk, has_head, xlam, indices=idx)
torch.cuda.synchronize()
if rank == 0:
    ref = build_cuda_model(oc, mapper, weights, "cuda:0")
    _, _, _, g_ref = cuda_step(ref, x.cuda(), y.cuda(), chunk, has_head, xlam, indices=idx)
    worst = max((rel_l2(g_dp[k], g_ref[k]), k) for k in g_ref if g_ref[k] is not None and g_ref[k].abs().max() > 0)
    print("DP vs single-GPU full batch: worst grad rel", worst)
    assert worst[0] < 2e-2
    print("DP OK")
dist.barrier()
dist.destroy_process_group()
