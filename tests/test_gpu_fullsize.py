"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run 2 M-row gathers
in seconds, so only the sampler is compared element by element; the float path is checked by invariants).
Reddit shape: 232 965 nodes / 11 M edges / d = 602, B = 8192 seeds, fanout [25, 10].  GPU only."""
import numpy as np
import pytest
import torch

from oracle import sampler as osampler
from oracle.mt19937 import MT19937Oracle

pytestmark = pytest.mark.gpu

B, S1, S2 = 8192, 25, 10


@pytest.fixture(scope='module')
def g():
    import pytorch_graphsage_b200 as g
    return g


@pytest.fixture(scope='module')
def world(g):
    from pytorch_graphsage_b200 import synth
    prob = synth.make_problem('reddit', seed=0, with_feats=False)
    adj = prob['adj']
    graph = g.GraphCSR.from_synth(adj)
    ids0 = synth.seed_batch(prob, B, seed=5)
    rng = g.DeviceMT19937(123 ** 2)
    s = g.SparseUniformNeighborSampler(graph, rng=rng)
    ids1 = s(ids=torch.from_numpy(ids0).cuda(), n_samples=S1)
    ids2 = s(ids=ids1, n_samples=S2)
    return dict(prob=prob, adj=adj, graph=graph, ids0=ids0, ids1=ids1, ids2=ids2, rng=rng)


def test_fullsize_sampler_bit_exact_vs_oracle(g, world):
    """2.25 M samples at full size: every index, and the stream position afterwards."""
    adj = world['adj']
    indptr, data, shape = adj['indptr'], adj['data'], adj['shape']
    deg = np.diff(indptr)
    indices = np.arange(data.shape[0], dtype=np.int64) - np.repeat(indptr[:-1], deg)
    o = MT19937Oracle(123 ** 2)
    want1 = osampler.sparse_sample(indptr, indices, data, shape, deg, world['ids0'], S1, o.randint)
    want2 = osampler.sparse_sample(indptr, indices, data, shape, deg, want1, S2, o.randint)
    assert np.array_equal(world['ids1'].cpu().numpy(), want1)
    assert np.array_equal(world['ids2'].cpu().numpy(), want2)
    st = world['rng'].get_state()
    assert np.array_equal(st[1], o.key) and st[2] == o.pos
    world['graph'].check()


def test_fullsize_samples_are_neighbours(g, world):
    """Property: every sampled id is one of its parent's stored neighbours (0 only for an empty row)."""
    adj = world['adj']
    indptr, data = adj['indptr'], adj['data']
    parents = world['ids1'].cpu().numpy()
    kids = world['ids2'].cpu().numpy().reshape(-1, S2)
    rs = np.random.RandomState(0)
    for p in rs.randint(0, parents.shape[0], 3000):
        row = data[indptr[parents[p]]:indptr[parents[p] + 1]]
        assert (np.isin(kids[p], row).all() if row.size else not kids[p].any())


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
def test_fullsize_gather_mean_invariants(g, world, dtype):
    """Linearity, neighbour-order invariance and a checksum of checksums for the fused gather+mean at full size."""
    n_rows, d = world['adj']['shape'][0], 602
    n1 = B * S1
    ids2 = world['ids2']
    gen = torch.Generator(device='cuda').manual_seed(1)
    ta = torch.randint(-8, 9, (n_rows, d), generator=gen, device='cuda').float()       # small integers: sums are exact in fp32
    tb = torch.randint(-8, 9, (n_rows, d), generator=gen, device='cuda').float()
    pad = lambda t: g.ops.pad_table(t, dtype)[0][:, :d]
    A, Bt, AB = pad(ta), pad(tb), pad(ta + tb)
    ma = g.ops.gather_reduce(A, ids2, n1, S2, 'sum', out_dtype=torch.float32)
    mb = g.ops.gather_reduce(Bt, ids2, n1, S2, 'sum', out_dtype=torch.float32)
    mab = g.ops.gather_reduce(AB, ids2, n1, S2, 'sum', out_dtype=torch.float32)
    assert torch.equal(ma + mb, mab)                                                    # linearity, exactly
    perm = ids2.view(n1, S2).flip(1).contiguous().view(-1)                              # reverse each parent's neighbour list
    assert torch.equal(g.ops.gather_reduce(A, perm, n1, S2, 'sum', out_dtype=torch.float32), ma)
    # checksum of checksums: sum over all outputs == sum over table rows weighted by how often each row was sampled
    counts = torch.bincount(ids2, minlength=n_rows).double()
    want = (counts.unsqueeze(1) * ta.double()).sum(dim=0)
    assert torch.equal(ma.double().sum(dim=0), want)
    # the mean is the sum scaled by 1/S (one fp32 multiply)
    mean = g.ops.gather_reduce(A, ids2, n1, S2, 'mean', out_dtype=torch.float32)
    assert torch.equal(mean, ma * (1.0 / S2))
    # max: idempotent under duplication of the neighbour list, and equal to torch's segmented max
    mx = g.ops.gather_reduce(A, ids2, n1, S2, 'max', out_dtype=torch.float32)
    dup = torch.cat([ids2.view(n1, S2), ids2.view(n1, S2)], dim=1).contiguous().view(-1)
    assert torch.equal(g.ops.gather_reduce(A, dup, n1, 2 * S2, 'max', out_dtype=torch.float32), mx)
    probe = torch.arange(0, n1, 97, device='cuda')
    assert torch.equal(mx[probe], ta[ids2.view(n1, S2)[probe]].max(dim=1)[0])


def test_fullsize_projection_linearity(g, world):
    """tcgen05 projection at the layer-1 shape (204 800 x 602 -> 2 x 128): exact on small-integer operands
    (every product and partial sum is representable), so it must equal the fp32 FFMA kernel bit for bit."""
    n, d, O = B * S1, 602, 128
    gen = torch.Generator(device='cuda').manual_seed(2)
    table = g.ops.pad_table(torch.randint(-4, 5, (world['adj']['shape'][0], d), generator=gen, device='cuda').float(), torch.bfloat16)[0][:, :d]
    m = g.ops.pad_table(torch.randint(-4, 5, (n, d), generator=gen, device='cuda').float(), torch.bfloat16)[0][:, :d]
    wx = g.ops.pad_table(torch.randint(-2, 3, (O, d), generator=gen, device='cuda').float(), torch.bfloat16)[0][:, :d]
    wn = g.ops.pad_table(torch.randint(-2, 3, (O, d), generator=gen, device='cuda').float(), torch.bfloat16)[0][:, :d]
    ids = world['ids1']
    segs = [dict(a=table, ids=ids, w=wx, col0=0), dict(a=m, w=wn, col0=O)]
    tc = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=False)
    ff = g.ops.linear(segs, n, act='relu', out_dtype=torch.float32, exact=True)
    assert torch.equal(tc, ff)
