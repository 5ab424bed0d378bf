#!/usr/bin/env python
"""
tests/golden/make_golden.py -- generates the committed golden fixtures by RUNNING THE REFERENCE.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference's hot path (`nn_modules.py`, `models.py`) imports cleanly on torch 2.11 / py3.12; the
only shim is `nn_modules.to_numpy` (the reference's `helpers.to_numpy`, helpers.py:21-25, recurses
forever on torch >= 0.4 because every Tensor is a Variable).  Nothing else is patched: the sampler
classes, prep classes, aggregator classes and `GSSupervised.forward` below are the reference's own.

Fixtures (np.savez_compressed, all small):
  sampler_canonical.npz   SparseUniformNeighborSampler on a reference-convention graph (zero-degree
                          nodes, dummy id 0 in the batch, maxdeg not a power of two)
  sampler_general.npz     same class on an arbitrary scipy matrix (gaps in columns, duplicate
                          entries, explicit zeros)
  sampler_dense.npz       UniformNeighborSampler (dense table, shared randperm)
  model_<agg>_<prep>.npz  GSSupervised.forward with the sparse sampler: weights, sampled ids per
                          hop, every aggregator-layer output, logits
"""

import os
import sys
import warnings

import numpy as np

REF = '/root/reference'
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import torch                                    # noqa: E402
from scipy.sparse import csr_matrix              # noqa: E402
import nn_modules as ref_nn                      # noqa: E402  (the reference)
import models as ref_models                      # noqa: E402  (the reference)
from torch.nn import functional as F             # noqa: E402
from functools import partial                    # noqa: E402

ref_nn.to_numpy = lambda t: t.detach().cpu().numpy()
warnings.filterwarnings('ignore')

from pytorch_graphsage_b200 import synth         # noqa: E402


def ref_parse_csr(x):
    v, r, c = x                                   # /root/reference/problem.py:70-72
    return csr_matrix((v, (r, c)))


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('%-28s %7.1f KB' % (name + '.npz', os.path.getsize(path) / 1024.0))


def rng_fingerprint():
    st = np.random.get_state()
    return np.asarray(st[1], dtype=np.uint32), np.int64(st[2])


def gen_sampler_canonical():
    adj = synth.make_sparse_adjacency(300, 2400, alpha=1.3, clip=37, seed=3, isolated_frac=0.15)
    trip = synth.triplets(adj)
    A = ref_parse_csr(trip)
    sampler = ref_nn.SparseUniformNeighborSampler(adj=A)
    ids0 = np.concatenate([[0, 1, A.shape[0] - 1], np.random.RandomState(1).randint(0, A.shape[0], 61)]).astype(np.int64)
    np.random.seed(123 ** 2)                      # /root/reference/train.py:133
    ids1 = sampler(ids=torch.LongTensor(ids0), n_samples=25).numpy()
    ids2 = sampler(ids=torch.LongTensor(ids1), n_samples=10).numpy()
    key, pos = rng_fingerprint()
    ids3 = sampler(ids=torch.LongTensor(ids0), n_samples=3).numpy()     # stream continues across calls
    save('sampler_canonical', trip=trip, shape=np.array(A.shape), degrees=sampler.degrees,
         ids0=ids0, ids1=ids1, ids2=ids2, ids3=ids3, seed=np.int64(123 ** 2), key_after=key, pos_after=pos)


def gen_sampler_general():
    rs = np.random.RandomState(11)
    R, C, nnz = 60, 23, 500
    r = rs.randint(0, R, nnz)
    c = rs.randint(0, C, nnz)
    v = rs.randint(0, 50, nnz)                    # includes explicit zeros and duplicate (r, c) pairs
    r[-1], c[-1], v[-1] = R - 1, C - 1, 7         # pin the inferred shape
    trip = np.vstack([v, r, c]).astype(np.int64)
    A = ref_parse_csr(trip)
    sampler = ref_nn.SparseUniformNeighborSampler(adj=A)
    ids0 = rs.randint(0, R, 40).astype(np.int64)
    np.random.seed(7)
    out = sampler(ids=torch.LongTensor(ids0), n_samples=9).numpy()
    save('sampler_general', trip=trip, shape=np.array(A.shape), degrees=sampler.degrees, ids0=ids0, out=out,
         seed=np.int64(7))


def gen_sampler_dense():
    rs = np.random.RandomState(5)
    n, K = 120, 128
    adj = rs.randint(0, n, (n + 1, K)).astype(np.int64)
    adj[n] = n                                    # dummy row LAST in the dense convention
    sampler = ref_nn.UniformNeighborSampler(adj=torch.LongTensor(adj))
    ids0 = rs.randint(0, n + 1, 33).astype(np.int64)
    torch.manual_seed(123)
    state = torch.get_rng_state()
    out1 = sampler(ids=torch.LongTensor(ids0), n_samples=25).numpy()
    out2 = sampler(ids=torch.LongTensor(out1.reshape(-1)), n_samples=10).numpy()
    torch.set_rng_state(state)
    perm1 = torch.randperm(K).numpy()
    perm2 = torch.randperm(K).numpy()
    save('sampler_dense', adj=adj, ids0=ids0, out1=out1, out2=out2, perm1=perm1, perm2=perm2, seed=np.int64(123))


def gen_model(agg, prep, fanout=(25, 10), batch=12, d=20, n_nodes=400, with_feats=True, out_dims=(16, 12), agg_kwargs=None):
    adj = synth.make_sparse_adjacency(n_nodes, n_nodes * 9, alpha=1.4, clip=45, seed=17, isolated_frac=0.05)
    trip = synth.triplets(adj)
    A = ref_parse_csr(trip)
    feats = synth.make_features(n_nodes, d, seed=2) if with_feats else None
    n_classes = 5
    torch.manual_seed(123)                        # /root/reference/train.py:81 -> weight init
    model = ref_models.GSSupervised(
        input_dim=(d if with_feats else None), n_nodes=A.shape[0], n_classes=n_classes,
        layer_specs=[
            dict(n_train_samples=fanout[0], n_val_samples=fanout[0], output_dim=out_dims[0], activation=F.relu),
            dict(n_train_samples=fanout[1], n_val_samples=fanout[1], output_dim=out_dims[1], activation=lambda x: x),
        ],
        aggregator_class=(partial(ref_nn.aggregator_lookup[agg], **agg_kwargs) if agg_kwargs else ref_nn.aggregator_lookup[agg]),
        prep_class=ref_nn.prep_lookup[prep],
        sampler_class=ref_nn.sampler_lookup['sparse_uniform_neighbor_sampler'], adj=A, train_adj=A)
    model.eval()
    ids0 = np.concatenate([[0], synth.seed_batch(dict(n_nodes=A.shape[0]), batch - 1, seed=4)]).astype(np.int64)

    hops, layer_outs = [], []
    real_sampler = model.train_sampler

    class Spy(object):                            # records what the reference's sampler returned
        degrees = real_sampler.degrees
        adj = real_sampler.adj

        def __call__(self, ids, n_samples):
            out = real_sampler(ids=ids, n_samples=n_samples)
            hops.append(out.numpy().copy())
            return out
    spy = Spy()
    model.train_sample_fns = [partial(spy, n_samples=s) for s in fanout]
    hooks = [m.register_forward_hook(lambda mod, inp, out: layer_outs.append(out.detach().numpy().copy()))
             for m in model.agg_layers.children()]
    np.random.seed(123 ** 2)
    with torch.no_grad():
        logits = model(torch.LongTensor(ids0), torch.FloatTensor(feats) if with_feats else None, train=True)
    for h in hooks:
        h.remove()
    key, pos = rng_fingerprint()
    arrays = dict(trip=trip, shape=np.array(A.shape), ids0=ids0, ids1=hops[0], ids2=hops[1],
                  fanout=np.array(fanout), out_dims=np.array(out_dims), logits=logits.numpy(),
                  l1_a=layer_outs[0], l1_b=layer_outs[1], l2=layer_outs[2], seed=np.int64(123 ** 2),
                  key_after=key, pos_after=pos, n_nodes=np.int64(A.shape[0]))
    if with_feats:
        arrays['feats'] = feats
    for k, v in model.state_dict().items():
        arrays['w:' + k] = v.numpy()
    save('model_%s_%s%s' % (agg, prep, '' if with_feats else '_nofeats'), **arrays)


def gen_model_dense(agg='mean', prep='identity', fanout=(25, 10), batch=12, d=20, n=300, K=128, out_dims=(16, 12)):
    """GSSupervised.forward with the DENSE sampler -- train.py:55's default (`uniform_neighbor_sampler`), dummy row last."""
    rs = np.random.RandomState(9)
    adj = rs.randint(0, n, (n + 1, K)).astype(np.int64)
    adj[rs.rand(n + 1, K) < 0.2] = n                 # some slots point at the dummy node, like convert.py:77 leaves them
    adj[n] = n
    feats = np.vstack([synth.make_features(n, d, seed=2)[1:n + 1], np.zeros((1, d), dtype=np.float32)]).astype(np.float32)
    n_classes = 5
    torch.manual_seed(123)
    model = ref_models.GSSupervised(
        input_dim=d, n_nodes=n + 1, n_classes=n_classes,
        layer_specs=[
            dict(n_train_samples=fanout[0], n_val_samples=fanout[0], output_dim=out_dims[0], activation=F.relu),
            dict(n_train_samples=fanout[1], n_val_samples=fanout[1], output_dim=out_dims[1], activation=lambda x: x),
        ],
        aggregator_class=ref_nn.aggregator_lookup[agg], prep_class=ref_nn.prep_lookup[prep],
        sampler_class=ref_nn.sampler_lookup['uniform_neighbor_sampler'], adj=torch.LongTensor(adj), train_adj=torch.LongTensor(adj))
    model.eval()
    ids0 = np.concatenate([[n], rs.randint(0, n, batch - 1)]).astype(np.int64)
    hops, layer_outs = [], []
    real_sampler = model.train_sampler

    def spy(ids, n_samples):
        out = real_sampler(ids=ids, n_samples=n_samples)
        hops.append(out.numpy().copy())
        return out
    model.train_sample_fns = [partial(spy, n_samples=s) for s in fanout]
    hooks = [m.register_forward_hook(lambda mod, inp, out: layer_outs.append(out.detach().numpy().copy()))
             for m in model.agg_layers.children()]
    torch.manual_seed(123 ** 2)                      # helpers.set_seeds at train.py:133 seeds torch as well
    with torch.no_grad():
        logits = model(torch.LongTensor(ids0), torch.FloatTensor(feats), train=True)
    for h in hooks:
        h.remove()
    arrays = dict(adj=adj, feats=feats, ids0=ids0, ids1=hops[0], ids2=hops[1], fanout=np.array(fanout), out_dims=np.array(out_dims),
                  logits=logits.numpy(), l1_a=layer_outs[0], l1_b=layer_outs[1], l2=layer_outs[2], seed=np.int64(123 ** 2),
                  n_nodes=np.int64(n + 1))
    for k, v in model.state_dict().items():
        arrays['w:' + k] = v.numpy()
    save('model_dense_%s_%s' % (agg, prep), **arrays)


if __name__ == '__main__':
    gen_sampler_canonical()
    gen_sampler_general()
    gen_sampler_dense()
    for agg in ('mean', 'max_pool', 'mean_pool', 'attention'):
        gen_model(agg, 'identity')
    gen_model('mean', 'linear')
    gen_model('mean', 'node_embedding')
    gen_model('mean', 'node_embedding', with_feats=False)       # the Pokec recipe (run.sh:30-34)
    gen_model('max_pool', 'node_embedding', with_feats=False)   # BASELINE config C3
    gen_model('attention', 'node_embedding', with_feats=False)
    gen_model_dense('mean', 'identity')                          # BASELINE config C1: train.py's default sampler
    # LSTMAggregator (nn_modules.py:259-286) with a 64-wide hidden state (the default 512 would make a 9 MB fixture)
    gen_model('lstm', 'identity', agg_kwargs=dict(hidden_dim=64))
    gen_model('lstm', 'node_embedding', with_feats=False, agg_kwargs=dict(hidden_dim=64))
