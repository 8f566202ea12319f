"""torchrun --nproc-per-node 2 tools/dp_check.py : data-parallel training over NCCL against one GPU.
 1. gradients of one step (overlapped all-reduce buckets) == the same module on the concatenated batch on one GPU
 2. the CUDA-graph training step (graphs.GraphedTrainStep, the all-reduce captured inside the graph) over several
    optimiser steps == the eager data-parallel path: same DCS draws, same losses, same parameters on every rank.
Run by tests/test_dp_nccl_gpu.py when two devices are visible."""
import os, random, sys
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from tests.util import O, build_cuda_model, cases, cuda_step, make_inputs, ref_cfg, rel_l2
from diverse_channel_vit_b200.dichavit import dichavit
from diverse_channel_vit_b200.graphs import GraphedTrainStep
from diverse_channel_vit_b200.optim import FusedAdamW
from diverse_channel_vit_b200.trainer_glue import training_loss

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
    print("DP vs single-GPU full batch: worst grad rel", worst, flush=True)
    assert worst[0] < 2e-2
dist.barrier()


# ---- 2. graphed DP steps vs eager DP steps ----
def fresh():
    cfg = ref_cfg(oc)
    cfg.update(dict(enable_sample=True, hcs_sampling="lowest_cosine_prob", hcs_sampling_temp=0.1))
    m = dichavit(cfg, mapper=mapper)
    m.load_state_dict({k: weights[k].clone() for k in m.state_dict() if k in weights}, strict=False)
    m = m.cuda().train()
    m.enable_data_parallel()
    return m, FusedAdamW(m, lr=2e-4, weight_decay=0.04, clip_grad_norm=1.0, device_schedule=True)


def seed(s):
    random.seed(s); torch.manual_seed(s + 2); torch.cuda.manual_seed_all(s + 4)


xs, ys = x[sl].cuda(), y[sl].cuda()
n_steps = 8
me, oe = fresh()
me.direct_grad = True
seed(11)
eager = []
for _ in range(n_steps):
    oe.zero_grad()
    out, extra = me(xs, chunk)
    loss = training_loss(me, out, extra, ys, has_head, xlam)
    loss.backward()
    oe.step()
    eager.append(loss.item())
mg, og = fresh()
step = GraphedTrainStep(mg, og, extra_loss_lambda=xlam)
seed(11)
graphed = [step(xs, ys, chunk).item() for _ in range(n_steps)]
torch.cuda.synchronize()
for i, (a, b) in enumerate(zip(graphed, eager)):
    assert abs(a - b) <= 2e-2 * abs(b) + 1e-3, (rank, i, a, b)
# replicas stay identical: every rank holds the same parameters after the graphed steps
div = []
for k, p_ in mg.named_parameters():
    r0 = p_.detach().clone()
    dist.broadcast(r0, src=0)
    if not torch.equal(r0, p_.detach()):
        div.append((rel_l2(p_, r0), (p_.detach() - r0).abs().max().item(), k))
if div:
    div.sort(reverse=True)
    print(f"rank {rank}: {len(div)} parameters differ from rank 0; worst {div[:5]}", flush=True)
assert not div, f"rank {rank}: replicas are no longer bit-identical"
worst_p = max(rel_l2(pg, pe_) for (k, pg), (_, pe_) in zip(mg.named_parameters(), me.named_parameters()) if k != "proxies")
if rank == 0:
    print(f"graphed DP == eager DP over {n_steps} steps: graphs {step.graph_launches} launches / {len(step.buckets)} buckets, "
          f"worst parameter rel {worst_p:.2e}", flush=True)
assert worst_p < 3e-2
dist.barrier()
if rank == 0:
    print("DP OK", flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)  # communicator not torn down: destroy_process_group() blocks while graphs that captured its collectives live
