"""Pins oracle/sampler.py and oracle/layers.py against fixtures produced by the reference's own classes
(tests/golden/make_golden.py) -- and, inside the build container, against the live reference."""
import numpy as np
import pytest
import torch

from oracle import layers, sampler
from oracle.mt19937 import MT19937Oracle
from tests import util


def _graph(fix):
    indptr, indices, data, shape = sampler.csr_from_triplets(*fix['trip'])
    assert tuple(shape) == tuple(fix['shape'])
    deg = sampler.row_degrees(indptr, data)
    assert np.array_equal(deg, fix['degrees']) if 'degrees' in fix else True
    return indptr, indices, data, shape, deg


def test_sparse_sampler_canonical():
    fix = util.load('sampler_canonical')
    g = _graph(fix)
    rng = MT19937Oracle(int(fix['seed']))
    ids1 = sampler.sparse_sample(*g, fix['ids0'], 25, rng.randint)
    ids2 = sampler.sparse_sample(*g, ids1, 10, rng.randint)
    assert np.array_equal(ids1, fix['ids1']) and np.array_equal(ids2, fix['ids2'])
    assert np.array_equal(rng.key, fix['key_after']) and rng.pos == int(fix['pos_after'])
    assert np.array_equal(sampler.sparse_sample(*g, fix['ids0'], 3, rng.randint), fix['ids3'])
    # zero-degree rows and the dummy id 0 sample the dummy
    deg = g[4]
    assert (deg == 0).sum() > 5 and fix['ids0'][0] == 0
    assert not ids1.reshape(-1, 25)[deg[fix['ids0']] == 0].any()


def test_sparse_sampler_numpy_global_stream():
    """Same thing drawing from numpy's RandomState directly (the reference's actual RNG object)."""
    fix = util.load('sampler_canonical')
    g = _graph(fix)
    rs = np.random.RandomState(int(fix['seed']))
    ids1 = sampler.sparse_sample(*g, fix['ids0'], 25, lambda hi, n: rs.choice(hi, n))
    assert np.array_equal(ids1, fix['ids1'])


def test_sparse_sampler_general_matrix():
    fix = util.load('sampler_general')
    g = _graph(fix)
    rng = MT19937Oracle(int(fix['seed']))
    assert np.array_equal(sampler.sparse_sample(*g, fix['ids0'], 9, rng.randint), fix['out'])


def test_sparse_sampler_rejects_out_of_range():
    fix = util.load('sampler_general')
    g = _graph(fix)
    with pytest.raises(IndexError):
        sampler.sparse_sample(*g, np.array([int(fix['shape'][0])]), 2, MT19937Oracle(1).randint)


def test_dense_sampler():
    fix = util.load('sampler_dense')
    out1 = sampler.dense_sample(fix['adj'], fix['ids0'], 25, fix['perm1'])
    out2 = sampler.dense_sample(fix['adj'], out1.reshape(-1), 10, fix['perm2'])
    assert np.array_equal(out1, fix['out1']) and np.array_equal(out2, fix['out2'])
    torch.manual_seed(int(fix['seed']))           # the permutation is torch's CPU generator
    assert np.array_equal(torch.randperm(128).numpy(), fix['perm1'])


@pytest.mark.parametrize('agg,prep,with_feats', util.MODEL_CASES)
def test_forward_stack_matches_reference(agg, prep, with_feats):
    fix = util.load(util.case_name(agg, prep, with_feats))
    g = _graph(fix)
    rng = MT19937Oracle(int(fix['seed']))
    S1, S2 = [int(s) for s in fix['fanout']]
    ids1 = sampler.sparse_sample(*g, fix['ids0'], S1, rng.randint)
    ids2 = sampler.sparse_sample(*g, ids1, S2, rng.randint)
    assert np.array_equal(ids1, fix['ids1']) and np.array_equal(ids2, fix['ids2'])
    assert np.array_equal(rng.key, fix['key_after']) and rng.pos == int(fix['pos_after'])

    params = util.params_of(fix)
    feats = torch.from_numpy(fix['feats']) if with_feats else None
    hop_ids = [torch.from_numpy(a) for a in (fix['ids0'], ids1, ids2)]
    out, trace = layers.forward_stack(hop_ids, feats, params, aggregator=agg, prep=prep,
                                      n_nodes=int(fix['n_nodes']), return_intermediates=True)
    tol = dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(trace['layer0'][0].numpy(), fix['l1_a'], **tol)
    np.testing.assert_allclose(trace['layer0'][1].numpy(), fix['l1_b'], **tol)
    np.testing.assert_allclose(trace['layer1'][0].numpy(), fix['l2'], **tol)
    np.testing.assert_allclose(out.numpy(), fix['logits'], **tol)

    # fp64 yardstick: the fp32 reference itself sits within ~1e-5 of exact arithmetic
    p64 = util.params_of(fix, torch.float64)
    out64 = layers.forward_stack(hop_ids, feats.double() if with_feats else None, p64, aggregator=agg, prep=prep,
                                 n_nodes=int(fix['n_nodes']))
    np.testing.assert_allclose(out64.numpy(), fix['logits'], rtol=2e-4, atol=2e-5)


def test_forward_stack_with_the_dense_sampler_matches_reference():
    """BASELINE config C1's path: train.py's default `uniform_neighbor_sampler` (dense table, one torch.randperm per hop)."""
    fix = util.load('model_dense_mean_identity')
    torch.manual_seed(int(fix['seed']))
    K = fix['adj'].shape[1]
    ids1 = sampler.dense_sample(fix['adj'], fix['ids0'], int(fix['fanout'][0]), torch.randperm(K).numpy())
    ids2 = sampler.dense_sample(fix['adj'], ids1.reshape(-1), int(fix['fanout'][1]), torch.randperm(K).numpy())
    assert np.array_equal(ids1, fix['ids1']) and np.array_equal(ids2, fix['ids2'])
    hop_ids = [torch.from_numpy(np.ascontiguousarray(a).reshape(-1)) for a in (fix['ids0'], ids1, ids2)]
    out = layers.forward_stack(hop_ids, torch.from_numpy(fix['feats']), util.params_of(fix), n_nodes=int(fix['n_nodes']))
    np.testing.assert_allclose(out.numpy(), fix['logits'], rtol=1e-5, atol=1e-6)


@pytest.mark.skipif(not util.have_reference(), reason='live reference only exists in the build container')
def test_oracle_against_live_reference_random_graphs():
    """Beyond the committed fixtures: fresh random graphs / seeds through the reference's sampler class."""
    from scipy.sparse import csr_matrix
    ref_nn, _ = util.import_reference()
    from pytorch_graphsage_b200 import synth
    for trial in range(4):
        adj = synth.make_sparse_adjacency(500 + 97 * trial, 6000, alpha=1.2 + 0.2 * trial, clip=20 + 31 * trial,
                                          seed=trial, isolated_frac=0.1)
        trip = synth.triplets(adj)
        A = csr_matrix((trip[0], (trip[1], trip[2])))
        ref = ref_nn.SparseUniformNeighborSampler(adj=A)
        g = _graph(dict(trip=trip, shape=np.array(A.shape), degrees=ref.degrees))
        ids = np.random.RandomState(trial).randint(0, A.shape[0], 200)
        np.random.seed(1000 + trial)
        want1 = ref(ids=torch.LongTensor(ids), n_samples=7).numpy()
        want2 = ref(ids=torch.LongTensor(want1), n_samples=4).numpy()
        rng = MT19937Oracle(1000 + trial)
        got1 = sampler.sparse_sample(*g, ids, 7, rng.randint)
        got2 = sampler.sparse_sample(*g, got1, 4, rng.randint)
        assert np.array_equal(want1, got1) and np.array_equal(want2, got2)
        st = np.random.get_state()
        assert np.array_equal(rng.key, st[1]) and rng.pos == st[2]
