"""gather_mean_project_umma.cu: the neighbour half of the mean aggregator as ONE kernel (gather + mean + projection; the reduced rows
never reach HBM).  Opt-in (GSAGE_FUSED_LAYER=1): measured against the two kernels it replaces it wins on rows <= 512 bytes and loses on the 1216-byte
reddit rows (profiles/README.md); training always uses the two-kernel path because the backward reads the reduced rows.  GPU only.

Bars: the engine's bf16 logits with the fused kernel equal the two-kernel bf16 logits (same arithmetic order: expected
bit-identical, asserted to 1e-3) and stay inside the bf16 bar (3e-2) against the fp32 reference fixture."""
import numpy as np
import pytest
import torch

from tests import util
from tests.test_gpu_model import build_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def g():
    import pytorch_graphsage_b200 as g
    return g


@pytest.mark.parametrize('case', ['model_mean_identity', 'model_mean_node_embedding_nofeats'])
def test_fused_gather_mean_project_equals_unfused(g, case, monkeypatch):
    fix = util.load(case)
    with_feats = 'feats' in fix
    prep = 'identity' if with_feats else 'node_embedding'
    feats = torch.from_numpy(fix['feats']) if with_feats else None
    ids = torch.from_numpy(fix['ids0'])
    monkeypatch.setenv('GSAGE_FUSED_LAYER', '0')
    base = build_model(g, fix, 'mean', prep, with_feats, compute_dtype=torch.bfloat16)
    g.set_seeds(int(fix['seed']))
    want = base(ids, feats, train=True).cpu().numpy()
    monkeypatch.setenv('GSAGE_FUSED_LAYER', '1')
    fused = build_model(g, fix, 'mean', prep, with_feats, compute_dtype=torch.bfloat16)
    g.set_seeds(int(fix['seed']))
    before = g.launch_count()
    got = fused(ids, feats, train=True).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(got, fix['logits'], rtol=3e-2, atol=3e-2)
    assert g.launch_count() > before


@pytest.mark.parametrize('S,d', [(10, 602), (25, 602), (10, 256), (25, 64), (7, 100)])
def test_fused_gather_mean_project_at_scale(g, S, d, monkeypatch):
    monkeypatch.setenv('GSAGE_FUSED_LAYER', '1')
    """Many tiles per CTA (the A tile is single-buffered: tile i+1's stores wait for tile i's MMAs), ragged last tile, the
    compile-time fanouts (10, 25) and the run-time one, every row-width class (1-3 sixteen-byte units per lane)."""
    from pytorch_graphsage_b200 import ops
    gen = torch.Generator().manual_seed(3)
    rows, O, n = 5000, 128, 148 * 48 * 3 + 17
    table = ops.pad_table(torch.randn((rows, d), generator=gen), torch.bfloat16)[0][:, :d]
    w = ops.pad_table(torch.randn((O, d), generator=gen) / 25, torch.bfloat16)[0][:, :d]
    ids = torch.randint(0, rows, (n * S,), generator=gen).cuda()
    m = ops.gather_reduce(table, ids, n, S, 'mean', d=d, out_dtype=torch.bfloat16)
    want = ops.linear([dict(a=m, w=w)], n, act='relu', out_dtype=torch.bfloat16, exact=False)
    got = ops.gather_mean_project(table, ids, n, S, w, act='relu')
    np.testing.assert_allclose(got.float().cpu().numpy(), want.float().cpu().numpy(), rtol=1e-2, atol=1e-2)


def test_fused_layer_is_off_when_activations_are_kept(g):
    """Training keeps the reduced rows (the backward reads them): the engine must take the two-kernel path there."""
    fix = util.load('model_mean_identity')
    from torch.nn import functional as F
    model = build_model(g, fix, 'mean', 'identity', True, compute_dtype=torch.bfloat16)
    targets = torch.from_numpy(np.random.RandomState(0).randint(0, fix['logits'].shape[1], fix['ids0'].shape[0])).cuda()
    g.set_seeds(int(fix['seed']))
    model.train_step(torch.from_numpy(fix['ids0']), torch.from_numpy(fix['feats']), targets, F.cross_entropy, optimizer=False, clip=None)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
