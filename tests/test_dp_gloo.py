"""CPU, world_size 2 over gloo: the bucketed gradient all-reduce of the data-parallel path (_GradReducer) and the
property that makes batch sharding exact: rank-averaged gradients of equal shards == gradient of the whole batch."""
import os
import socket
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, overlap, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from diverse_channel_vit_b200.dichavit import _GradReducer

        depth = 12
        sizes = {"embed": 1000, **{f"block{i}": 700 + i for i in range(depth)}, "tail": 300}
        groups, off = [], 0
        for name in ["embed"] + [f"block{i}" for i in range(depth)] + ["tail"]:
            groups.append((name, off, off + sizes[name]))
            off += sizes[name]
        module = types.SimpleNamespace(_groups=groups, _pg=None, _comm_stream=None, _overlap=overlap)
        g = torch.Generator().manual_seed(100 + rank)
        grad = torch.randn(off, generator=g)
        want = sum(torch.randn(off, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
        red = _GradReducer(module, grad)
        # backward order: tail, blocks 11..0, embed
        red.ready("tail", flush=True)
        for i in reversed(range(depth)):
            red.ready(f"block{i}")
        red.ready("embed", flush=True)
        red.finish()
        ok = torch.allclose(grad, want, atol=1e-6)
        covered = sorted(red.ranges)
        contiguous = covered[0][0] == 0 and covered[-1][1] == off and all(
            covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
        q.put((rank, bool(ok), bool(contiguous), len(covered)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_bucketed_allreduce_world2(overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, contiguous, n in res:
        assert ok and contiguous
        assert n == (1 if not overlap else 6)  # tail | 4 x 3 blocks (last merged with embed) -> 1 + 3 + 1 + ... buckets


def test_shard_average_equals_full_batch_gradient():
    """main CE loss and TDL are batch means, CDL is batch independent -> with equal shards the
    rank-averaged gradient equals the single-process gradient of the concatenated batch (oracle, fp64)."""
    from tests.util import O, cases, make_inputs

    oc, mapper, chunk, has_head, B, wseed, iseed, xlam = cases()["tiny_jumpcp"]
    oc.depth = 2
    weights = {k: v.double() for k, v in O.make_weights(oc, has_head, wseed).items()}
    x, y = make_inputs(oc, 4, 8, oc.num_classes, iseed)
    idx = [5, 0, 3]
    _, _, g_full = O.loss_and_grads(x, y, weights, oc, mapper[chunk], has_head, indices=idx, extra_loss_lambda=xlam,
                                    dtype=torch.float64)
    shards = [O.loss_and_grads(x[s], y[s], weights, oc, mapper[chunk], has_head, indices=idx, extra_loss_lambda=xlam,
                               dtype=torch.float64)[2] for s in (slice(0, 2), slice(2, 4))]
    for k, g in g_full.items():
        if g is None:
            continue
        avg = (shards[0][k] + shards[1][k]) / 2
        torch.testing.assert_close(avg, g, rtol=1e-9, atol=1e-12)


def _bcast_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.util import O, cases, ref_cfg
        from diverse_channel_vit_b200.dichavit import dichavit

        oc, mapper, chunk, has_head, *_ = cases()["tiny_jumpcp"]
        torch.manual_seed(1000 + rank)  # the reference seeds every process differently by default (trainer.py:81)
        m = dichavit(ref_cfg(oc), mapper=mapper)
        before = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).clone()
        m.enable_data_parallel()             # flat buffer not built yet: the broadcast happens when it is
        m._ensure_flat(torch.device("cpu"))
        after = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
        gathered = [torch.empty_like(after) for _ in range(world)]
        dist.all_gather(gathered, after)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        changed = not torch.equal(before, after)
        # trainer-owned parameters: their gradient is averaged on request
        m.proxies.grad = torch.full_like(m.proxies, float(rank + 1))
        m.allreduce_external_grads()
        ext_ok = torch.allclose(m.proxies.grad, torch.full_like(m.proxies, (1 + world) / 2))
        q.put((rank, bool(same), bool(changed), bool(ext_ok)))
    finally:
        dist.destroy_process_group()


def test_enable_data_parallel_broadcasts_rank0_parameters():
    """DDP(model) broadcasts rank 0's parameters (trainer.py:1185); enable_data_parallel must too, or replicas that
    were initialised from different seeds would apply the averaged gradient to different weights for ever."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(same for _, same, _, _ in res) and all(ext for _, _, _, ext in res)
    assert res[0][2] is False and res[1][2] is True  # rank 0 keeps its weights, rank 1 received them
