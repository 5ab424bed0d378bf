"""Shared helpers for the test-suite (fixtures loader; reference importer for in-container checks)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
REFERENCE = '/root/reference'

MODEL_CASES = [
    ('mean', 'identity', True), ('max_pool', 'identity', True), ('mean_pool', 'identity', True),
    ('attention', 'identity', True), ('mean', 'linear', True), ('mean', 'node_embedding', True),
    ('mean', 'node_embedding', False), ('max_pool', 'node_embedding', False),
    ('attention', 'node_embedding', False),
    ('lstm', 'identity', True), ('lstm', 'node_embedding', False),          # hidden_dim = 64 (fixture size)
]


def case_name(agg, prep, with_feats):
    return 'model_%s_%s%s' % (agg, prep, '' if with_feats else '_nofeats')


def aggregator_class(g, fix, agg):
    """The registry class for `agg`; the LSTM fixtures were made with a 64-wide state (hidden_dim is a constructor keyword of the
    reference's class too), so the width is read back from the stored weights."""
    cls = g.aggregator_lookup[agg]
    if agg == 'lstm':
        from functools import partial
        return partial(cls, hidden_dim=int(fix['w:agg_layers.0.lstm.weight_hh_l0'].shape[1]))
    return cls


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def params_of(fix, dtype=torch.float32):
    """The reference `state_dict` stored in a model fixture -> {name: tensor}."""
    return {k[2:]: torch.from_numpy(v).to(dtype) if v.dtype.kind == 'f' else torch.from_numpy(v)
            for k, v in fix.items() if k.startswith('w:')}


def have_reference():
    return os.path.isdir(REFERENCE)


def import_reference():
    """The live reference, for in-container cross-checks only (never on the GPU box)."""
    sys.dont_write_bytecode = True
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import nn_modules as ref_nn
    import models as ref_models
    ref_nn.to_numpy = lambda t: t.detach().cpu().numpy()
    return ref_nn, ref_models
