"""`loss.backward()` THROUGH the plug-in API (`agg(x, neibs)` / `prep(...)` are autograd-differentiable in the
reference, models.py:100-101): the narrow operator calls of operators.py record a torch autograd graph whose backward passes are
library kernels (gsage_wgrad, gsage_linear transposed, narrow_backward.cu).  Compared with torch autograd through the CPU oracle on
the fixtures' own sampled ids.  GPU only.  fp32: rtol 2e-3 / atol 2e-5 (atomics reorder the weight-gradient sums)."""
import numpy as np
import pytest
import torch
from torch.nn import functional as F

from oracle import layers
from tests import util
from tests.test_gpu_model import build_model

pytestmark = pytest.mark.gpu

TOL = dict(rtol=2e-3, atol=2e-5)


@pytest.fixture(scope='module')
def g():
    import pytorch_graphsage_b200 as g
    return g


@pytest.mark.parametrize('agg,prep,with_feats', [
    ('mean', 'identity', True), ('max_pool', 'identity', True), ('mean_pool', 'identity', True), ('attention', 'identity', True),
    ('mean', 'linear', True), ('mean', 'node_embedding', True), ('mean', 'node_embedding', False),
    ('max_pool', 'node_embedding', False), ('attention', 'node_embedding', False), ('lstm', 'identity', True)])
def test_loss_backward_through_the_plugins(g, agg, prep, with_feats):
    """models.py:97-104 as the reference runs it: forward through sampler / prep / aggregator plug-ins, loss, loss.backward(),
    every parameter gradient against the oracle's -- including the registry combinations the fused engine's backward does not
    cover (attention + node_embedding, node_embedding WITH features, pool / attention in fp32)."""
    fix = util.load(util.case_name(agg, prep, with_feats))
    model = build_model(g, fix, agg, prep, with_feats)
    feats = torch.from_numpy(fix['feats']) if with_feats else None
    targets = torch.from_numpy(np.random.RandomState(0).randint(0, fix['logits'].shape[1], fix['ids0'].shape[0]))
    g.set_seeds(int(fix['seed']))
    model.zero_grad()
    preds = model.forward_reference_order(torch.from_numpy(fix['ids0']), feats, train=True)
    np.testing.assert_allclose(preds.detach().cpu().numpy(), fix['logits'], rtol=1e-4, atol=1e-5)
    loss = F.cross_entropy(preds, targets.cuda())
    loss.backward()

    hop_ids = [torch.from_numpy(fix[k]) for k in ('ids0', 'ids1', 'ids2')]
    ps = {k: v.clone().requires_grad_(True) for k, v in util.params_of(fix).items()}
    kw = dict(n_nodes=int(fix['n_nodes'])) if prep == 'node_embedding' else {}
    want = F.cross_entropy(layers.forward_stack(hop_ids, feats, ps, aggregator=agg, prep=prep, **kw), targets)
    want.backward()
    assert abs(loss.item() - want.item()) < 1e-4
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        np.testing.assert_allclose(p.grad.cpu().numpy(), ps[name].grad.numpy(), err_msg=name, **TOL)


@pytest.mark.parametrize('agg', ['mean', 'max_pool', 'mean_pool', 'attention', 'lstm'])
def test_aggregator_call_is_differentiable_wrt_its_inputs(g, agg):
    """agg(x, neibs): gradients w.r.t. x and neibs (what lets layer 2 back-propagate into layer 1), vs the oracle's aggregator."""
    gen = torch.Generator().manual_seed(11)
    n, S, d, O = 37, 7, 20, 16
    torch.manual_seed(3)
    kw = dict(hidden_dim=24) if agg == 'lstm' else {}
    mod = g.aggregator_lookup[agg](input_dim=d, output_dim=O, activation=F.relu, **kw).cuda()
    x = torch.randn((n, d), generator=gen)
    nb = torch.randn((n * S, d), generator=gen)
    xg, nbg = x.cuda().requires_grad_(True), nb.cuda().requires_grad_(True)
    out = mod(xg, nbg)
    probe = torch.randn(out.shape, generator=gen)
    (out * probe.cuda()).sum().backward()

    ps = {'a.' + k: v.detach().cpu().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    xr, nbr = x.clone().requires_grad_(True), nb.clone().requires_grad_(True)
    want = layers.AGGREGATORS[agg](xr, nbr, ps, 'a.', 'relu')
    np.testing.assert_allclose(out.detach().cpu().numpy(), want.detach().numpy(), rtol=1e-4, atol=1e-5)
    (want * probe).sum().backward()
    np.testing.assert_allclose(xg.grad.cpu().numpy(), xr.grad.numpy(), **TOL)
    np.testing.assert_allclose(nbg.grad.cpu().numpy(), nbr.grad.numpy(), **TOL)
    for name, p in mod.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), ps['a.' + name].grad.numpy(), err_msg=name, **TOL)


def test_stock_torch_optimizer_with_zero_grad_set_to_none(g):
    """torch's optimizer.zero_grad() sets p.grad = None (set_to_none=True is the default): the next backward must hand the
    parameters their gradients again, or clip + step silently do nothing."""
    fix = util.load('model_mean_identity')
    model = build_model(g, fix, 'mean', 'identity', True)
    feats = torch.from_numpy(fix['feats'])
    ids = torch.from_numpy(fix['ids0'])
    targets = torch.from_numpy(np.random.RandomState(2).randint(0, fix['logits'].shape[1], ids.shape[0])).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    for step in range(2):
        opt.zero_grad()                                                       # reference order: models.py:98
        assert all(p.grad is None for p in model.parameters())
        g.set_seeds(int(fix['seed']))
        model.train_step(ids, feats, targets, F.cross_entropy, optimizer=opt)
        assert all(p.grad is not None for p in model.parameters())
    moved = [n for n, p in model.named_parameters() if (p.detach() - before[n]).abs().max().item() > 0]
    assert sorted(moved) == sorted(before), 'parameters that never moved: %s' % sorted(set(before) - set(moved))


def test_own_optimizer_steps_by_default_and_follows_set_progress(g):
    """models.py:64-69, 93-104: the model owns its Adam; train_step(ids, feats, targets, loss_fn) -- the reference's exact call --
    updates the weights and returns preds; set_progress changes the learning rate the next step uses."""
    fix = util.load('model_mean_identity')
    graph_model = build_model(g, fix, 'mean', 'identity', True, lr_init=0.02, lr_schedule='linear')
    feats = torch.from_numpy(fix['feats'])
    ids = torch.from_numpy(fix['ids0'])
    targets = torch.from_numpy(np.random.RandomState(2).randint(0, fix['logits'].shape[1], ids.shape[0])).cuda()
    assert graph_model.lr == 0.02 and graph_model.optimizer.param_groups[0]['lr'] == 0.02
    graph_model.set_progress(0.5)
    assert abs(graph_model.lr - 0.01) < 1e-12 and abs(graph_model.optimizer.param_groups[0]['lr'] - 0.01) < 1e-12
    before = graph_model.fc.weight.detach().clone()
    g.set_seeds(int(fix['seed']))
    preds = graph_model.train_step(ids, feats, targets, F.cross_entropy)
    assert torch.is_tensor(preds) and tuple(preds.shape) == (ids.shape[0], fix['logits'].shape[1])
    step1 = (graph_model.fc.weight.detach() - before).abs().max().item()
    assert 0 < step1 <= 0.01 * 1.001                                          # Adam's first step moves every weight by ~lr


def test_sample_ahead_waits_for_the_producer_of_its_ids(g):
    """the sampler stream must be ordered after whatever produced the next batch's ids on the caller's stream.
    Here the ids are produced by a kernel queued right before the call, behind a long-running kernel: without the event the
    sampler stream would copy the buffer before it is written."""
    fix = util.load('model_mean_identity')
    feats = torch.from_numpy(fix['feats'])
    ids = torch.from_numpy(fix['ids0']).cuda()
    B = ids.shape[0]
    want = {}
    for mode in ('plain', 'ahead'):
        model = build_model(g, fix, 'mean', 'identity', True)
        g.set_seeds(int(fix['seed']))
        model(ids, feats)                                                         # batch 0
        if mode == 'plain':
            out = model(ids.flip(0).contiguous(), feats)
        else:
            big = torch.randn((4096, 4096), device='cuda')
            stage = torch.zeros((B,), dtype=torch.int64, device='cuda')           # stale content: zeros (the dummy node)
            for _ in range(20):
                big = big @ big * 1e-3                                            # keeps the stream busy for a few ms
            stage.copy_(ids.flip(0))                                              # the producer, queued behind the matmuls
            model.sample_ahead(stage, feats)                                      # standalone call: ordered after everything queued so far
            out = model(stage, feats)
        want[mode] = (out.cpu().numpy(), model.peek('ids2').cpu().numpy())
    assert np.array_equal(want['plain'][1], want['ahead'][1]), 'sample-ahead read its ids before they were written'
    np.testing.assert_allclose(want['plain'][0], want['ahead'][0], rtol=1e-5, atol=1e-6)


def test_forward_next_ids_is_bit_identical_and_ordered(g):
    """forward(..., next_ids=): the ids are marked ready BEFORE the forward is queued (gsage_engine_inputs_ready), so the sampling
    overlaps it; ids and logits equal the unpipelined run."""
    fix = util.load('model_mean_identity')
    feats = torch.from_numpy(fix['feats'])
    base = torch.from_numpy(fix['ids0'])
    batches = [base, base.flip(0).contiguous(), base.roll(5).contiguous()]
    model = build_model(g, fix, 'mean', 'identity', True)
    g.set_seeds(int(fix['seed']))
    plain = [(model(b, feats).cpu().numpy(), model.peek('ids2').cpu().numpy()) for b in batches]
    model = build_model(g, fix, 'mean', 'identity', True)
    g.set_seeds(int(fix['seed']))
    dev = [b.cuda() for b in batches]
    model.sample_ahead(dev[0], feats)
    for i, b in enumerate(dev):
        out = model(b, feats, next_ids=dev[i + 1] if i + 1 < len(dev) else None)
        assert np.array_equal(model.peek('ids2').cpu().numpy(), plain[i][1])
        np.testing.assert_allclose(out.cpu().numpy(), plain[i][0], rtol=1e-5, atol=1e-6)


def test_out_of_range_ids_raise_index_error_one_call_late(g):
    """`feats[ids]` (models.py:76) raises IndexError for an id outside the table; the device path flags it (sticky, mapped host
    memory) and the NEXT forward -- or model.check() -- raises, without a synchronisation in between."""
    fix = util.load('model_mean_identity')
    model = build_model(g, fix, 'mean', 'identity', True)
    feats = torch.from_numpy(fix['feats'])
    ids = torch.from_numpy(fix['ids0']).clone()
    g.set_seeds(1)
    model(ids, feats)
    model.check()                                                               # clean
    bad = ids.clone()
    bad[3] = 10 ** 7
    model(bad, feats)
    with pytest.raises(IndexError):
        model.check()
    model(bad, feats)
    torch.cuda.synchronize()
    with pytest.raises(IndexError):
        model(ids, feats)                                                       # polled at the start of the next forward
    model(ids, feats)                                                           # reported once, cleared
    model.check()


def test_feature_table_cache_follows_in_place_updates(g):
    """the padded device copy of `feats` is keyed by tensor object + version, not by address."""
    fix = util.load('model_mean_identity')
    model = build_model(g, fix, 'mean', 'identity', True)
    feats = torch.from_numpy(fix['feats']).clone()
    ids = torch.from_numpy(fix['ids0'])
    g.set_seeds(int(fix['seed']))
    a = model(ids, feats).cpu().numpy()
    feats[:, ::2].mul_(-1.0)                                                    # in place: same data_ptr, new _version (a plain rescale
                                                                                # would be undone by the final F.normalize)
    g.set_seeds(int(fix['seed']))
    b = model(ids, feats).cpu().numpy()
    assert np.abs(a - b).max() > 1e-3
    g.set_seeds(int(fix['seed']))
    fresh = torch.from_numpy(fix['feats']).clone()
    fresh[:, ::2] *= -1.0
    c = model(ids, fresh).cpu().numpy()
    np.testing.assert_allclose(b, c, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('task', ['classification', 'multilabel_classification', 'regression_mae'])
def test_device_metrics_match_sklearn(g, task):
    """problem.py:44-64 on the device (gsage_metric_f1 / gsage_metric_mae) vs the sklearn calls the reference makes."""
    from oracle import metrics as ometrics
    from pytorch_graphsage_b200.problem import ProblemMetrics
    rs = np.random.RandomState(5)
    for n, C in ((1, 3), (77, 5), (4096, 41), (513, 121)):
        if task == 'classification':
            preds = rs.randn(n, C).astype(np.float32)
            preds[:, C - 1] -= 10                                               # a class that is never predicted nor (below) true
            y = rs.randint(0, max(1, C - 1), size=(n, 1))
            want = ometrics.classification(y, preds)
        elif task == 'multilabel_classification':
            preds = rs.randn(n, C).astype(np.float32)
            y = (rs.rand(n, C) < 0.2).astype(np.float32)
            preds[:, 0] = -1
            y[:, 0] = 0                                                         # an empty label: F1 0, counted in macro
            want = ometrics.multilabel_classification(y, preds)
        else:
            preds = (rs.rand(n, 1) * 50).astype(np.float32)
            y = rs.randint(15, 61, size=(n, 1)).astype(np.float32)
            want = ometrics.regression_mae(y, preds)
        got_host = getattr(ProblemMetrics, task)(y, preds)                      # numpy in (uploaded)
        got_dev = getattr(ProblemMetrics, task)(torch.from_numpy(y).cuda(), torch.from_numpy(preds).cuda())
        assert got_host == got_dev
        if task == 'regression_mae':
            assert abs(got_dev - want) < 1e-4 * max(1.0, abs(want))
        else:
            assert abs(got_dev['micro'] - want['micro']) < 1e-9 and abs(got_dev['macro'] - want['macro']) < 1e-9


def test_three_layer_stack_runs_over_the_plugin_operators(g):
    """models.py:50-62,85-86 loop over however many `layer_specs` there are (train.py happens to pass two, which is what the fused
    engine implements).  Any other depth takes the reference's dataflow over the plug-in operators: same draws, same logits, and it
    trains.  Three layers, fanouts 4 / 3 / 2, against the oracle on ids re-sampled from the same MT19937 stream."""
    from oracle import sampler as osampler
    from oracle.mt19937 import MT19937Oracle
    from pytorch_graphsage_b200 import synth
    prob = synth.make_problem('tiny', seed=2)
    adj = prob['adj']
    graph = g.GraphCSR.from_synth(adj)
    torch.manual_seed(9)
    specs = [dict(n_train_samples=4, n_val_samples=4, output_dim=16, activation=F.relu),
             dict(n_train_samples=3, n_val_samples=3, output_dim=16, activation=F.relu),
             dict(n_train_samples=2, n_val_samples=2, output_dim=8, activation=lambda x: x)]
    model = g.GSSupervised(input_dim=prob['feats_dim'], n_nodes=prob['n_nodes'], n_classes=prob['n_classes'], layer_specs=specs,
                           aggregator_class=g.aggregator_lookup['mean'], prep_class=g.prep_lookup['identity'],
                           sampler_class=g.sampler_lookup['sparse_uniform_neighbor_sampler'], adj=graph, train_adj=graph).cuda()
    assert not model.fused and not model.has_fused_backward()
    ids0 = synth.seed_batch(prob, 40, seed=5)
    feats = torch.from_numpy(prob['feats'])
    targets = torch.from_numpy(np.random.RandomState(3).randint(0, prob['n_classes'], ids0.shape[0]))
    g.set_seeds(31)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    preds = model.train_step(torch.from_numpy(ids0), feats, targets.cuda(), F.cross_entropy, optimizer=False, clip=None)

    indptr, data, shape = adj['indptr'], adj['data'], adj['shape']
    indices = np.arange(data.shape[0]) - np.repeat(indptr[:-1], np.diff(indptr))
    deg = osampler.row_degrees(indptr, data)
    rs = MT19937Oracle(31)
    hops = [ids0]
    for S in (4, 3, 2):
        hops.append(osampler.sparse_sample(indptr, indices, data, shape, deg, hops[-1], S, rs.randint))
    ps = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in before.items()}
    want = layers.forward_stack([torch.from_numpy(h) for h in hops], feats, ps, acts=('relu', 'relu', 'identity'))
    np.testing.assert_allclose(preds.cpu().numpy(), want.detach().numpy(), rtol=1e-4, atol=1e-5)
    F.cross_entropy(want, targets).backward()
    for name, p in model.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), ps[name].grad.numpy(), err_msg=name, **TOL)
