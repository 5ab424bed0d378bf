"""CPU-side checks: the C-ABI library loads and exports every symbol include/gsage_b200.h declares; host-side
logic that needs no GPU (synthetic problem conventions, activation mapping, registries).  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
from torch.nn import functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    from pytorch_graphsage_b200 import build
    return build.build()


def header_functions():
    text = open(os.path.join(ROOT, 'include', 'gsage_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gsage_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    names = header_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), 'include/gsage_b200.h declares %s but the library does not export it' % name


def test_python_binding_covers_the_header(built):
    from pytorch_graphsage_b200 import _lib
    assert sorted(_lib.EXPORTS) == header_functions()
    assert _lib.lib().gsage_abi_version() == 3


def test_struct_layouts_match_the_header(built):
    """ctypes mirrors vs sizeof() as the C compiler sees the header."""
    import subprocess, tempfile
    from pytorch_graphsage_b200 import _lib
    src = '#include <stdio.h>\n#include "gsage_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(gsage_linear_seg), ' \
          'sizeof(gsage_engine_config), sizeof(gsage_layer_weights), sizeof(gsage_weights));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, 't.c'), 'w').write(src)
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), os.path.join(d, 't.c'), '-o', os.path.join(d, 't')])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, 't')]).split()]
    assert sizes == [ctypes.sizeof(_lib.LinearSeg), ctypes.sizeof(_lib.EngineConfig), ctypes.sizeof(_lib.LayerWeights),
                     ctypes.sizeof(_lib.Weights)]


def test_no_cpu_fallback_without_gpu(built):
    import pytorch_graphsage_b200 as g
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(g.GsageError):
        g.DeviceMT19937(1)
    with pytest.raises(g.GsageError):
        g.GraphCSR.from_triplets(np.array([[1], [1], [0]]))


def test_registries_mirror_the_reference():
    import pytorch_graphsage_b200 as g
    assert set(g.sampler_lookup) == {'uniform_neighbor_sampler', 'sparse_uniform_neighbor_sampler'}
    assert set(g.prep_lookup) == {'identity', 'node_embedding', 'linear'}
    assert set(g.aggregator_lookup) == {'mean', 'max_pool', 'mean_pool', 'lstm', 'attention'}          # nn_modules.py:324-330
    agg = g.aggregator_lookup['max_pool'](input_dim=12, output_dim=7, activation=F.relu)
    assert agg.output_dim == 14                                       # AggregatorMixin.output_dim, nn_modules.py:178-182
    assert sorted(agg.state_dict()) == ['fc_neib.weight', 'fc_x.weight', 'mlp.0.bias', 'mlp.0.weight']
    att = g.aggregator_lookup['attention'](input_dim=12, output_dim=7, activation=None)
    assert sorted(att.state_dict()) == ['att.0.weight', 'att.2.weight', 'fc_neib.weight', 'fc_x.weight']
    lstm = g.aggregator_lookup['lstm'](input_dim=12, output_dim=7, activation=None, hidden_dim=64)
    assert sorted(lstm.state_dict()) == ['fc_neib.weight', 'fc_x.weight', 'lstm.bias_hh_l0', 'lstm.bias_ih_l0', 'lstm.weight_hh_l0',
                                         'lstm.weight_ih_l0'] and lstm.output_dim == 14 and lstm.fc_neib.in_features == 64
    with pytest.raises(NotImplementedError):
        g.aggregator_lookup['lstm'](input_dim=12, output_dim=7, activation=None, bidirectional=True)
    prep = g.prep_lookup['node_embedding'](input_dim=5, n_nodes=10)
    assert prep.output_dim == 69 and prep.embedding.weight.shape == (11, 64)
    assert g.prep_lookup['node_embedding'](input_dim=None, n_nodes=10).output_dim == 64


def test_state_dict_interchanges_with_reference_fixture():
    import pytorch_graphsage_b200 as g
    from tests import util
    fix = util.load('model_max_pool_node_embedding_nofeats')
    params = util.params_of(fix)
    agg = g.aggregator_lookup['max_pool'](input_dim=64, output_dim=int(fix['out_dims'][0]), activation=F.relu)
    agg.load_state_dict({k[len('agg_layers.0.'):]: v for k, v in params.items() if k.startswith('agg_layers.0.')})


def test_activation_mapping():
    from pytorch_graphsage_b200.operators import _act_name
    assert _act_name(F.relu) == 'relu' and _act_name(lambda x: x) is None and _act_name(None) is None
    assert _act_name(torch.tanh) == 'tanh'
    with pytest.raises(ValueError):
        _act_name(torch.sigmoid)


def test_synth_follows_the_file_convention():
    from pytorch_graphsage_b200 import synth
    from oracle import sampler as osampler
    adj = synth.make_sparse_adjacency(500, 6000, alpha=1.3, clip=41, seed=2, isolated_frac=0.1)
    trip = synth.triplets(adj)
    indptr, indices, data, shape = osampler.csr_from_triplets(*trip)
    assert shape == adj['shape'] == (501, int(np.diff(adj['indptr']).max()))
    assert np.array_equal(indptr, adj['indptr']) and np.array_equal(data, adj['data'])
    assert indptr[1] == 0                                             # row 0 = dummy, empty
    assert data.min() >= 1 and data.max() <= 500                     # +1 id space
    assert np.array_equal(indices, np.arange(data.shape[0]) - np.repeat(indptr[:-1], np.diff(indptr)))
    feats = synth.make_features(500, 12)
    assert feats.shape == (501, 12) and not feats[0].any()
    prob = synth.make_problem('tiny')
    assert prob['n_nodes'] == prob['adj']['shape'][0] == 1001
