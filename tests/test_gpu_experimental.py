"""Kernels that were written without GPU time left in the round (no run behind them yet).  Opt-in only --
GSAGE_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -- because an untested tcgen05 pipeline that
deadlocks traps after its bounded mbarrier wait and takes the CUDA context of the whole pytest process with it.

gather_mean_project_umma.cu (GSAGE_FUSED_LAYER=1): the neighbour half of the mean aggregator as ONE kernel (gather + mean +
projection).  Bar: the engine's bf16 logits with the fused kernel equal the unfused bf16 logits (same arithmetic order: expected
bit-identical, asserted to 1e-3) and stay inside the bf16 bar against the fp32 reference fixture."""
import os

import numpy as np
import pytest
import torch

from tests import util
from tests.test_gpu_model import build_model

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get('GSAGE_TEST_EXPERIMENTAL') != '1', reason='opt-in: GSAGE_TEST_EXPERIMENTAL=1')]


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
    base = build_model(g, fix, 'mean', prep, with_feats, compute_dtype=torch.bfloat16)
    g.set_seeds(int(fix['seed']))
    want = base(ids, feats, train=True).cpu().numpy()
    monkeypatch.setenv('GSAGE_FUSED_LAYER', '1')
    fused = build_model(g, fix, 'mean', prep, with_feats, compute_dtype=torch.bfloat16)
    g.set_seeds(int(fix['seed']))
    got = fused(ids, feats, train=True).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(got, fix['logits'], rtol=3e-2, atol=3e-2)


def test_fused_gather_mean_project_at_scale(g, monkeypatch):
    """Many tiles per CTA (the A tile is single-buffered: tile i+1's stores wait for tile i's MMAs), ragged last tile."""
    from pytorch_graphsage_b200 import ops
    gen = torch.Generator().manual_seed(3)
    rows, d, O, S, n = 5000, 602, 128, 10, 148 * 48 * 3 + 17
    table = ops.pad_table(torch.randn((rows, d), generator=gen), torch.bfloat16)[0][:, :d]
    w = ops.pad_table(torch.randn((O, d), generator=gen) / 25, torch.bfloat16)[0][:, :d]
    ids = torch.randint(0, rows, (n * S,), generator=gen).cuda()
    m = ops.gather_reduce(table, ids, n, S, 'mean', d=d, out_dtype=torch.bfloat16)
    want = ops.linear([dict(a=m, w=w)], n, act='relu', out_dtype=torch.bfloat16, exact=False)
    monkeypatch.setenv('GSAGE_FUSED_LAYER', '1')
    got = ops.gather_mean_project(table, ids, n, S, w, act='relu')
    np.testing.assert_allclose(got.float().cpu().numpy(), want.float().cpu().numpy(), rtol=1e-2, atol=1e-2)
