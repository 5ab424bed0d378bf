"""Two GPUs, one process each (NCCL): seed-sharded forward + backward reproduces the single-GPU run --
sampled ids bit-exact (each rank consumes the global batch's draws and keeps its slice), all-reduced gradients
equal the full-batch gradients.  Needs >= 2 GPUs (skipped on the 1-GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch
from torch.nn import functional as F

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, skew=0):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import pytorch_graphsage_b200 as g
    from pytorch_graphsage_b200.parallel import shard_seeds
    from tests import util
    from tests.test_gpu_model import build_model
    fix = util.load('model_mean_identity')
    model = build_model(g, fix, 'mean', 'identity', True)
    feats = torch.from_numpy(fix['feats'])
    ids = torch.from_numpy(fix['ids0'])
    targets = torch.from_numpy(np.random.RandomState(0).randint(0, fix['logits'].shape[1], ids.shape[0]))
    if skew:                                                          # uneven shards: rank 0 takes `skew` seeds more than half
        cut = ids.shape[0] // 2 + skew
        mine, tmine = (ids[:cut], targets[:cut]) if rank == 0 else (ids[cut:], targets[cut:])
        first = 0 if rank == 0 else cut
    else:
        mine, tmine = shard_seeds(ids, rank, world), shard_seeds(targets, rank, world)
        first = int(sum(shard_seeds(ids, r, world).shape[0] for r in range(rank)))
    g.set_seeds(int(fix['seed']))                                  # every rank: the same stream as the single process
    side = torch.cuda.Stream()
    preds = model.train_step(mine, feats, tmine.cuda(), F.cross_entropy, optimizer=False, clip=None,
                                grad_scale=mine.shape[0] / ids.shape[0], overlap_stream=side, shard=(ids.shape[0], first))
    torch.cuda.synchronize()
    st = g.default_rng().get_state()
    q.put((rank, first, model.peek('ids2').cpu().numpy(), preds.cpu().numpy(),
           {n: p.grad.cpu().numpy() for n, p in model.named_parameters()}, st[1], st[2], model._bucket().collective))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('skew', [0, 2])
def test_two_gpu_sharded_step_equals_single_gpu(skew):
    """skew = 2: shards of different sizes (8 and 4 of the fixture's 12 seeds) -- every rank weights its gradient by ITS local/global batch before the sum
    (scaling after the sum lets replicas drift)."""
    import torch.multiprocessing as mp
    from oracle import layers
    from tests import util
    fix = util.load('model_mean_identity')
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, skew)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    S1, S2 = [int(s) for s in fix['fanout']]
    ids2 = np.concatenate([r[2] for r in res])
    assert np.array_equal(ids2, fix['ids2']), 'sharded sampling is not the single-process sampling'
    np.testing.assert_allclose(np.concatenate([r[3] for r in res]), fix['logits'], rtol=1e-4, atol=1e-5)
    for r in res:                                                   # every rank ends at the single-process stream position
        assert np.array_equal(r[5], fix['key_after']) and r[6] == int(fix['pos_after'])
    # all-reduced gradients == full-batch autograd gradients of the oracle
    targets = torch.from_numpy(np.random.RandomState(0).randint(0, fix['logits'].shape[1], fix['ids0'].shape[0]))
    ps = {k: v.clone().requires_grad_(True) for k, v in util.params_of(fix).items()}
    hop = [torch.from_numpy(fix[k]) for k in ('ids0', 'ids1', 'ids2')]
    F.cross_entropy(layers.forward_stack(hop, torch.from_numpy(fix['feats']), ps), targets).backward()
    for name, want in ps.items():
        for r in res:
            np.testing.assert_allclose(r[4][name], want.grad.numpy(), rtol=2e-3, atol=2e-5, err_msg=name)
    print('collective:', res[0][7])
