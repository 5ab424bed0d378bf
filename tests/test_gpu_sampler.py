"""Samplers through the C ABI vs golden fixtures (made by the reference) and vs the oracle.  GPU only."""
import numpy as np
import pytest
import torch

from oracle import sampler as osampler
from oracle.mt19937 import MT19937Oracle
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def g():
    import pytorch_graphsage_b200 as g
    return g


def _ids(a):
    return torch.from_numpy(np.asarray(a, dtype=np.int64)).cuda()


@pytest.mark.parametrize('mode', ['device', 'host'])
def test_sparse_canonical_golden(g, mode):
    fix = util.load('sampler_canonical')
    graph = g.GraphCSR.from_triplets(fix['trip'])
    assert graph.shape == tuple(fix['shape']) and graph.canonical
    assert np.array_equal(graph.degrees, fix['degrees'])
    rng = g.DeviceMT19937(int(fix['seed']))
    np.random.seed(int(fix['seed']))
    s = g.SparseUniformNeighborSampler(graph, rng=rng, mode=mode)
    ids1 = s(ids=_ids(fix['ids0']), n_samples=25)
    ids2 = s(ids=ids1, n_samples=10)
    assert ids1.dtype == torch.int64 and ids1.is_cuda and ids1.shape == (fix['ids0'].shape[0] * 25,)
    assert np.array_equal(ids1.cpu().numpy(), fix['ids1']) and np.array_equal(ids2.cpu().numpy(), fix['ids2'])
    st = rng.get_state() if mode == 'device' else np.random.get_state()
    assert np.array_equal(st[1], fix['key_after']) and st[2] == int(fix['pos_after'])
    assert np.array_equal(s(ids=_ids(fix['ids0']), n_samples=3).cpu().numpy(), fix['ids3'])
    graph.check()


def test_sparse_from_scipy_matches_triplets(g):
    from scipy.sparse import csr_matrix
    fix = util.load('sampler_canonical')
    A = csr_matrix((fix['trip'][0], (fix['trip'][1], fix['trip'][2])))
    s = g.SparseUniformNeighborSampler(A, rng=g.DeviceMT19937(int(fix['seed'])))
    assert np.array_equal(s.degrees, fix['degrees'])
    assert np.array_equal(s(ids=_ids(fix['ids0']), n_samples=25).cpu().numpy(), fix['ids1'])
    with pytest.raises(AssertionError):
        g.SparseUniformNeighborSampler(np.zeros((3, 3)))          # nn_modules.py:73


def test_sparse_general_matrix_golden(g):
    fix = util.load('sampler_general')
    graph = g.GraphCSR.from_triplets(fix['trip'])
    assert not graph.canonical and graph.shape == tuple(fix['shape'])
    assert np.array_equal(graph.degrees, fix['degrees'])
    s = g.SparseUniformNeighborSampler(graph, rng=g.DeviceMT19937(int(fix['seed'])))
    assert np.array_equal(s(ids=_ids(fix['ids0']), n_samples=9).cpu().numpy(), fix['out'])


def test_cpu_ids_round_trip(g):
    fix = util.load('sampler_canonical')
    s = g.SparseUniformNeighborSampler(g.GraphCSR.from_triplets(fix['trip']), rng=g.DeviceMT19937(int(fix['seed'])))
    out = s(ids=torch.from_numpy(fix['ids0']), n_samples=25)        # CPU in -> CPU out, like the reference
    assert not out.is_cuda and np.array_equal(out.numpy(), fix['ids1'])


def test_errors(g):
    fix = util.load('sampler_general')
    graph = g.GraphCSR.from_triplets(fix['trip'])
    s = g.SparseUniformNeighborSampler(graph, rng=g.DeviceMT19937(1))
    with pytest.raises(AssertionError):
        s(ids=_ids([1, 2]), n_samples=0)                             # nn_modules.py:81
    s(ids=_ids([int(fix['shape'][0])]), n_samples=2)                 # out of range: scipy raises IndexError
    with pytest.raises(IndexError):
        graph.check()
    graph.check()                                                    # flag is cleared once raised
    assert s(ids=_ids([]), n_samples=4).shape == (0,)                # empty batch


def test_dense_golden(g):
    fix = util.load('sampler_dense')
    s = g.UniformNeighborSampler(torch.from_numpy(fix['adj']))
    torch.manual_seed(int(fix['seed']))
    out1 = s(ids=_ids(fix['ids0']), n_samples=25)
    out2 = s(ids=out1.contiguous().view(-1), n_samples=10)
    assert out1.shape == (fix['ids0'].shape[0], 25)
    assert np.array_equal(out1.cpu().numpy(), fix['out1']) and np.array_equal(out2.cpu().numpy(), fix['out2'])


@pytest.mark.parametrize('trial', range(4))
def test_sparse_vs_oracle_random_graphs(g, trial):
    """Fresh graphs (power-law, isolated nodes, max degree not a power of two) against the CPU oracle."""
    from pytorch_graphsage_b200 import synth
    adj = synth.make_sparse_adjacency(3000 + 501 * trial, 40000, alpha=1.2 + 0.15 * trial, clip=50 + 77 * trial,
                                      seed=trial, isolated_frac=0.1)
    graph = g.GraphCSR.from_synth(adj)
    indptr, data, shape = adj['indptr'], adj['data'], adj['shape']
    indices = np.arange(data.shape[0]) - np.repeat(indptr[:-1], np.diff(indptr))
    deg = osampler.row_degrees(indptr, data)
    assert np.array_equal(graph.degrees, deg)
    ids0 = np.random.RandomState(trial).randint(0, shape[0], 512)
    rng, o = g.DeviceMT19937(1000 + trial), MT19937Oracle(1000 + trial)
    s = g.SparseUniformNeighborSampler(graph, rng=rng)
    got1 = s(ids=_ids(ids0), n_samples=25)
    got2 = s(ids=got1, n_samples=10)
    want1 = osampler.sparse_sample(indptr, indices, data, shape, deg, ids0, 25, o.randint)
    want2 = osampler.sparse_sample(indptr, indices, data, shape, deg, want1, 10, o.randint)
    assert np.array_equal(got1.cpu().numpy(), want1) and np.array_equal(got2.cpu().numpy(), want2)
    st = rng.get_state()
    assert np.array_equal(st[1], o.key) and st[2] == o.pos


def test_epoch_iterator_matches_reference_shuffle(g):
    """problem.py:141-153: the epoch shuffle comes from the same stream, before the sampler draws."""
    from pytorch_graphsage_b200.problem import iterate
    nodes = np.arange(1, 1338)
    targets = np.arange(2000)
    g.set_seeds(99)
    got = [(ids.cpu().numpy(), t.cpu().numpy(), p) for ids, t, p in iterate(nodes, targets, batch_size=512, shuffle=True)]
    rs = np.random.RandomState(99)
    idx = rs.permutation(np.arange(nodes.shape[0]))                  # what NodeProblem.iterate does
    chunks = np.array_split(idx, nodes.shape[0] // 512 + 1)
    assert len(got) == len(chunks)
    for (ids, t, p), (k, c) in zip(got, enumerate(chunks)):
        assert np.array_equal(ids, nodes[c]) and np.array_equal(t, targets[nodes[c]]) and p == k / len(chunks)
    # the stream continues where numpy's would: the next sampler draw matches
    assert np.array_equal(g.ops.u32_to_numpy(g.default_rng().randint(1000, 50)).astype(np.int64), rs.choice(1000, 50))
