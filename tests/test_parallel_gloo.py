"""N > 1 host logic on CPU: world_size-2 gloo process group (no GPU): seed sharding partitions the batch, the flat
gradient bucket all-reduces to the sum, and local/global weighting reproduces the single-process mean gradient."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pytorch_graphsage_b200.parallel import FlatGradBucket, shard_seeds
    torch.manual_seed(0)                                       # identical replicas
    lin1, lin2 = torch.nn.Linear(7, 5, bias=False), torch.nn.Linear(5, 3)
    params = list(lin1.parameters()) + list(lin2.parameters())
    bucket = FlatGradBucket(params, head=list(lin2.parameters()), device='cpu')
    bucket.attach()
    ids = torch.arange(101)
    mine = shard_seeds(ids, rank, world)
    # per-rank "gradient" = sum over my seeds of a per-seed term; the weighted all-reduce must give the global mean
    x = torch.randn(101, 7, generator=torch.Generator().manual_seed(1))
    loss = lin2(torch.relu(lin1(x[mine]))).pow(2).mean()
    grads = torch.autograd.grad(loss, params)
    for p, g in zip(params, grads):
        bucket.grad_of(p).copy_(g)
    scale = mine.shape[0] / ids.shape[0]                       # 51/101 on rank 0, 50/101 on rank 1: uneven shards
    bucket.all_reduce_head(scale)                              # the bucket weights BEFORE the sum: sum_r scale_r * g_r
    bucket.all_reduce_tail(scale)
    ref = torch.autograd.grad(lin2(torch.relu(lin1(x))).pow(2).mean(), params)
    ok = all(torch.allclose(p.grad, r, rtol=1e-5, atol=1e-6) for p, r in zip(params, ref))
    head_first = bucket.params[0] is lin2.weight or bucket.params[0] is lin2.bias
    out.put((rank, mine.tolist(), ok, head_first, bucket.head_numel))
    dist.destroy_process_group()


def test_shard_and_bucket_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seeds = res[0][1] + res[1][1]
    assert seeds == list(range(101)) and abs(len(res[0][1]) - len(res[1][1])) <= 1          # a partition of the batch
    assert all(r[2] for r in res), 'weighted all-reduce != single-process gradient of the global mean loss'
    assert all(r[3] for r in res) and res[0][4] == 64 + 64                                   # head params lead the bucket


def test_shard_seeds_partition_property():
    from pytorch_graphsage_b200.parallel import shard_seeds
    for n in (0, 1, 7, 512, 1000):
        for world in (1, 2, 3, 8):
            ids = torch.arange(n)
            parts = [shard_seeds(ids, r, world) for r in range(world)]
            assert torch.equal(torch.cat(parts), ids)
            want = [len(a) for a in np.array_split(np.arange(n), world)]
            assert [p.shape[0] for p in parts] == want
