"""End-to-end parity of the engine (C ABI gsage_engine_forward) with GSSupervised.forward of the reference,
on the golden fixtures the reference itself produced.  GPU only.

Bars: sampled ids bit-exact (and the RNG stream position after the batch); fp32 activations/logits within
rtol 1e-4 / atol 1e-5 of the reference's fp32 values (the reference itself sits ~1e-5 from fp64, see
tests/test_oracle_golden.py); bf16 compute within rtol 3e-2 / atol 3e-2 of the fp32 reference."""
import numpy as np
import pytest
import torch
from torch.nn import functional as F

from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def g():
    import pytorch_graphsage_b200 as g
    return g


def build_model(g, fix, agg, prep, with_feats, compute_dtype=torch.float32, **kw):
    graph = g.GraphCSR.from_triplets(fix['trip'])
    S1, S2 = [int(s) for s in fix['fanout']]
    O1, O2 = [int(s) for s in fix['out_dims']]
    d = fix['feats'].shape[1] if with_feats else None
    model = g.GSSupervised(
        input_dim=d, n_nodes=int(fix['n_nodes']), n_classes=fix['logits'].shape[1],
        layer_specs=[dict(n_train_samples=S1, n_val_samples=S1, output_dim=O1, activation=F.relu),
                     dict(n_train_samples=S2, n_val_samples=S2, output_dim=O2, activation=lambda x: x)],
        aggregator_class=util.aggregator_class(g, fix, agg), prep_class=g.prep_lookup[prep],
        sampler_class=g.sampler_lookup['sparse_uniform_neighbor_sampler'], adj=graph, train_adj=graph,
        compute_dtype=compute_dtype, **kw)
    missing = model.load_state_dict(util.params_of(fix), strict=True)      # the reference's own state_dict
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.cuda()


@pytest.mark.parametrize('agg,prep,with_feats', util.MODEL_CASES)
def test_engine_matches_reference_fp32(g, agg, prep, with_feats):
    fix = util.load(util.case_name(agg, prep, with_feats))
    model = build_model(g, fix, agg, prep, with_feats)
    feats = torch.from_numpy(fix['feats']) if with_feats else None
    g.set_seeds(int(fix['seed']))                                         # train.py:133
    logits = model(torch.from_numpy(fix['ids0']), feats, train=True)
    # sampled indices: bit-exact, and the stream ends where numpy's ended
    assert np.array_equal(model.peek('ids1').cpu().numpy(), fix['ids1'])
    assert np.array_equal(model.peek('ids2').cpu().numpy(), fix['ids2'])
    st = g.default_rng().get_state()
    assert np.array_equal(st[1], fix['key_after']) and st[2] == int(fix['pos_after'])
    tol = dict(rtol=1e-4, atol=1e-5)
    B = fix['ids0'].shape[0]
    l1 = model.peek('layer1').float().cpu().numpy()
    np.testing.assert_allclose(l1[:B], fix['l1_a'], **tol)
    np.testing.assert_allclose(l1[B:], fix['l1_b'], **tol)
    np.testing.assert_allclose(model.peek('layer2').cpu().numpy(), fix['l2'], **tol)
    np.testing.assert_allclose(logits.cpu().numpy(), fix['logits'], **tol)


@pytest.mark.parametrize('agg,prep,with_feats', util.MODEL_CASES)
def test_narrow_operator_api_matches_reference(g, agg, prep, with_feats):
    """The reference's own call order through sampler(ids) / prep(ids, feats) / agg(x, neibs)."""
    fix = util.load(util.case_name(agg, prep, with_feats))
    model = build_model(g, fix, agg, prep, with_feats)
    feats = torch.from_numpy(fix['feats']) if with_feats else None
    g.set_seeds(int(fix['seed']))
    logits = model.forward_reference_order(torch.from_numpy(fix['ids0']), feats, train=True)
    np.testing.assert_allclose(logits.detach().cpu().numpy(), fix['logits'], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('agg,prep,with_feats', [('mean', 'identity', True), ('max_pool', 'identity', True),
                                                 ('attention', 'identity', True), ('mean', 'node_embedding', False),
                                                 ('lstm', 'identity', True)])
def test_engine_bf16_compute(g, agg, prep, with_feats):
    fix = util.load(util.case_name(agg, prep, with_feats))
    model = build_model(g, fix, agg, prep, with_feats, compute_dtype=torch.bfloat16)
    feats = torch.from_numpy(fix['feats']) if with_feats else None
    g.set_seeds(int(fix['seed']))
    logits = model(torch.from_numpy(fix['ids0']), feats, train=True)
    assert np.array_equal(model.peek('ids2').cpu().numpy(), fix['ids2'])       # sampling is dtype independent
    np.testing.assert_allclose(logits.cpu().numpy(), fix['logits'], rtol=3e-2, atol=3e-2)


def test_lstm_engine_in_blocks_of_parents(g, monkeypatch):
    """The LSTM aggregator projects the input half of the gates for a BLOCK of parents at once (<= 2 GiB of gate rows); with a
    60-row block the 12-seed fixture runs as 6 + 50 + 6 blocks and must give the same logits."""
    monkeypatch.setenv('GSAGE_LSTM_BLOCK_ROWS', '60')
    fix = util.load('model_lstm_identity')
    model = build_model(g, fix, 'lstm', 'identity', True)
    g.set_seeds(int(fix['seed']))
    logits = model(torch.from_numpy(fix['ids0']), torch.from_numpy(fix['feats']), train=True)
    np.testing.assert_allclose(model.peek('layer2').cpu().numpy(), fix['l2'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(logits.detach().cpu().numpy(), fix['logits'], rtol=1e-4, atol=1e-5)


def test_host_buffer_entry_point(g):
    """gsage_engine_forward_host: pinned host ids in, pinned host logits out (the e2e path bench.py times)."""
    fix = util.load('model_mean_identity')
    model = build_model(g, fix, 'mean', 'identity', True)
    ids = torch.from_numpy(fix['ids0']).pin_memory()
    out = torch.empty(fix['logits'].shape, dtype=torch.float32).pin_memory()
    g.set_seeds(int(fix['seed']))
    model.forward_host(ids, torch.from_numpy(fix['feats']), out)
    np.testing.assert_allclose(out.numpy(), fix['logits'], rtol=1e-4, atol=1e-5)


def test_consecutive_batches_continue_the_stream(g):
    """Two batches back to back == the reference's two consecutive forward calls (same global stream)."""
    fix = util.load('model_mean_identity')
    model = build_model(g, fix, 'mean', 'identity', True)
    feats = torch.from_numpy(fix['feats'])
    from oracle import layers, sampler as osampler
    from oracle.mt19937 import MT19937Oracle
    indptr, indices, data, shape = osampler.csr_from_triplets(*fix['trip'])
    deg = osampler.row_degrees(indptr, data)
    o = MT19937Oracle(int(fix['seed']))
    g.set_seeds(int(fix['seed']))
    params = util.params_of(fix)
    for batch in (fix['ids0'], fix['ids0'][::-1].copy(), fix['ids0'][:5]):
        ids1 = osampler.sparse_sample(indptr, indices, data, shape, deg, batch, 25, o.randint)
        ids2 = osampler.sparse_sample(indptr, indices, data, shape, deg, ids1, 10, o.randint)
        want = layers.forward_stack([torch.from_numpy(a) for a in (batch, ids1, ids2)], feats, params)
        got = model(torch.from_numpy(batch), feats)
        assert np.array_equal(model.peek('ids2').cpu().numpy(), ids2)
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)


def test_engine_tf32_projection_mode(g):
    """fp32 tables, projections on the tensor cores as TF32 (allow_tf32=True): ids still bit-exact, logits within 5e-3."""
    fix = util.load('model_mean_node_embedding_nofeats')
    model = build_model(g, fix, 'mean', 'node_embedding', False, allow_tf32=True)
    g.set_seeds(int(fix['seed']))
    logits = model(torch.from_numpy(fix['ids0']), None)
    assert np.array_equal(model.peek('ids2').cpu().numpy(), fix['ids2'])
    np.testing.assert_allclose(logits.cpu().numpy(), fix['logits'], rtol=5e-3, atol=5e-3)


def test_sample_ahead_is_bit_identical(g):
    """gsage_engine_sample_ahead: batch i+1 is drawn on the engine's own stream while batch i aggregates.  Draw order on
    the RNG stream = call order, so ids and logits equal the plain back-to-back run (and the oracle's)."""
    fix = util.load('model_mean_identity')
    feats = torch.from_numpy(fix['feats'])
    batches = [torch.from_numpy(b).cuda() for b in (fix['ids0'], fix['ids0'][::-1].copy(), fix['ids0'][:7].copy(), fix['ids0'])]
    plain = build_model(g, fix, 'mean', 'identity', True)
    g.set_seeds(int(fix['seed']))
    want = []
    for b in batches:
        out = plain(b, feats)
        want.append((plain.peek('ids1').cpu().numpy(), plain.peek('ids2').cpu().numpy(), out.cpu().numpy()))
    state_plain = g.default_rng().get_state()

    model = build_model(g, fix, 'mean', 'identity', True)
    g.set_seeds(int(fix['seed']))
    model.sample_ahead(batches[0], feats)
    for i, b in enumerate(batches):
        if i + 1 < len(batches):
            # queue the forward of batch i right after the sample-ahead of batch i+1 is requested?  No: the engine holds
            # ONE pending batch, so the order is forward(i) [consumes the pending sample] then sample_ahead(i+1)
            pass
        out = model(b, feats)
        ids1, ids2 = model.peek('ids1').cpu().numpy(), model.peek('ids2').cpu().numpy()
        if i + 1 < len(batches):
            model.sample_ahead(batches[i + 1], feats)
        assert np.array_equal(ids1, want[i][0]) and np.array_equal(ids2, want[i][1])
        np.testing.assert_array_equal(out.cpu().numpy(), want[i][2])
    st = g.default_rng().get_state()
    assert np.array_equal(st[1], state_plain[1]) and st[2] == state_plain[2]
    # the first batch also equals the reference's golden trace
    assert np.array_equal(want[0][1], fix['ids2'])


def test_sample_ahead_rejects_a_different_batch(g):
    fix = util.load('model_mean_identity')
    feats = torch.from_numpy(fix['feats'])
    model = build_model(g, fix, 'mean', 'identity', True)
    a = torch.from_numpy(fix['ids0']).cuda()
    b = torch.from_numpy(fix['ids0'][::-1].copy()).cuda()
    g.set_seeds(1)
    model.sample_ahead(a, feats)
    with pytest.raises(ValueError):
        model.sample_ahead(b, feats)            # one pending batch at a time
    with pytest.raises(ValueError):
        model(b, feats)                         # its draws are already consumed: the forward must name the same batch
    model(a, feats)


def test_sample_ahead_host_entry(g):
    fix = util.load('model_mean_identity')
    model = build_model(g, fix, 'mean', 'identity', True)
    ids = torch.from_numpy(fix['ids0']).pin_memory()
    out = torch.empty(fix['logits'].shape, dtype=torch.float32).pin_memory()
    g.set_seeds(int(fix['seed']))
    model.sample_ahead(ids, torch.from_numpy(fix['feats']), host=True)
    model.forward_host(ids, torch.from_numpy(fix['feats']), out)
    assert np.array_equal(model.peek('ids2').cpu().numpy(), fix['ids2'])
    np.testing.assert_allclose(out.numpy(), fix['logits'], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('dtype,tol', [(torch.float32, dict(rtol=1e-4, atol=1e-5)), (torch.bfloat16, dict(rtol=3e-2, atol=3e-2))])
def test_engine_with_the_dense_sampler_matches_reference(g, dtype, tol):
    """train.py's DEFAULT sampler (`uniform_neighbor_sampler`, nn_modules.py:19-49) through the engine: the two
    torch.randperm(K) draws come from the CPU generator like the reference's, so ids are bit-exact under the same seed."""
    fix = util.load('model_dense_mean_identity')
    S1, S2 = [int(s) for s in fix['fanout']]
    O1, O2 = [int(s) for s in fix['out_dims']]
    adj = torch.from_numpy(fix['adj'])
    model = g.GSSupervised(
        input_dim=fix['feats'].shape[1], n_nodes=int(fix['n_nodes']), n_classes=fix['logits'].shape[1],
        layer_specs=[dict(n_train_samples=S1, n_val_samples=S1, output_dim=O1, activation=F.relu),
                     dict(n_train_samples=S2, n_val_samples=S2, output_dim=O2, activation=lambda x: x)],
        aggregator_class=g.aggregator_lookup['mean'], prep_class=g.prep_lookup['identity'],
        sampler_class=g.sampler_lookup['uniform_neighbor_sampler'], adj=adj, train_adj=adj, compute_dtype=dtype)
    model.load_state_dict(util.params_of(fix), strict=True)
    model = model.cuda()
    g.set_seeds(int(fix['seed']))
    logits = model(torch.from_numpy(fix['ids0']), torch.from_numpy(fix['feats']), train=True)
    assert np.array_equal(model.peek('ids1').cpu().numpy(), fix['ids1'].reshape(-1))
    assert np.array_equal(model.peek('ids2').cpu().numpy(), fix['ids2'].reshape(-1))
    np.testing.assert_allclose(logits.cpu().numpy(), fix['logits'], **tol)


def test_pipelined_host_entry_equals_plain_host_entry(g):
    """gsage_engine_forward_host_next with a next batch: that batch's H2D, sampling AND forward are queued before the call
    blocks on its own logits (which travel on a copy stream).  Same draws in the same order: logits bit-identical to the
    unpipelined calls, and the RNG stream ends at the same position."""
    fix = util.load('model_mean_identity')
    feats = torch.from_numpy(fix['feats'])
    batches = [torch.from_numpy(b).pin_memory() for b in (fix['ids0'], fix['ids0'][::-1].copy(), fix['ids0'][:7].copy(), fix['ids0'])]
    n_classes = fix['logits'].shape[1]
    plain = build_model(g, fix, 'mean', 'identity', True)
    g.set_seeds(int(fix['seed']))
    want = []
    for b in batches:
        out = torch.empty((b.shape[0], n_classes), dtype=torch.float32).pin_memory()
        plain.forward_host(b, feats, out)
        want.append(out.numpy().copy())
    state_plain = g.default_rng().get_state()
    np.testing.assert_allclose(want[0], fix['logits'], rtol=1e-4, atol=1e-5)

    model = build_model(g, fix, 'mean', 'identity', True)
    g.set_seeds(int(fix['seed']))
    for i, b in enumerate(batches):
        out = torch.empty((b.shape[0], n_classes), dtype=torch.float32).pin_memory()
        nxt = batches[i + 1] if i + 1 < len(batches) else None
        model.forward_host(b, feats, out, next_ids_host=nxt)
        np.testing.assert_array_equal(out.numpy(), want[i])
    st = g.default_rng().get_state()
    assert np.array_equal(st[1], state_plain[1]) and st[2] == state_plain[2]
    # a queued forward must be collected before anything else runs on the engine
    out = torch.empty((batches[0].shape[0], n_classes), dtype=torch.float32).pin_memory()
    model.forward_host(batches[0], feats, out, next_ids_host=batches[1])
    with pytest.raises(ValueError):
        model(torch.from_numpy(fix['ids0']).cuda(), feats)
    with pytest.raises(ValueError):
        model.forward_host(batches[2], feats, out)
    out1 = torch.empty((batches[1].shape[0], n_classes), dtype=torch.float32).pin_memory()
    model.forward_host(batches[1], feats, out1)
