"""
The reference's plug-in operators, re-implemented over libgsage_b200 -- same class names, constructor
arguments, call signatures, parameter names and registries as /root/reference/nn_modules.py, so that

    from pytorch_graphsage_b200.operators import sampler_lookup, prep_lookup, aggregator_lookup

is a drop-in for the three dicts train.py resolves its CLI strings through (train.py:95,99,100), and a
reference `state_dict` loads into these modules unchanged.

Two entry points per aggregator:
  * `forward(x, neibs)`           -- the reference's narrow contract (rows already gathered): reduce + both
                                     projections + concat + activation run in the library, and the call is
                                     autograd-differentiable like the reference's (models.py:100-101 back-propagates
                                     through nn_modules.py:196-204, 223-232, 305-321): every library call is wrapped in a
                                     torch.autograd.Function whose backward pass is again library kernels
                                     (gsage_wgrad, gsage_linear with a transposed weight, narrow_backward.cu);
  * `forward_ids(table, ids_self, ids_neib, S)` -- the wide contract: the gather is fused into the reduce
                                     and projection kernels, neighbour rows are never materialised (forward only; the
                                     engine owns the fused backward, model.GSSupervised.backward).
"""

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import ops
from .graph import GraphCSR
from .rng import default_rng
from ._lib import check, lib


def _act_name(fn):
    """Map the callables train.py passes (F.relu, `lambda x: x`) onto the library's activation enum."""
    if fn is None:
        return None
    if fn in (F.relu, torch.relu):
        return 'relu'
    if fn in (torch.tanh, F.tanh):
        return 'tanh'
    if isinstance(fn, str):
        return fn
    probe = torch.tensor([-2.0, 0.5])
    res = fn(probe)
    if torch.equal(res, probe):
        return None
    if torch.equal(res, torch.relu(probe)):
        return 'relu'
    if torch.allclose(res, torch.tanh(probe)):
        return 'tanh'
    raise ValueError('gsage: unsupported activation %r (identity / relu / tanh)' % (fn,))


def _cuda_ids(ids):
    ids = torch.as_tensor(ids)
    was_cuda = ids.is_cuda
    ids = ids.to(device='cuda', dtype=torch.int64).contiguous().view(-1)
    return ids, was_cuda


# --
# autograd plumbing of the narrow API: forward and backward are both library calls, torch only records the graph

def _f32(t):
    t = t if torch.is_tensor(t) else torch.as_tensor(t)
    return t.to(device='cuda', dtype=torch.float32)


class _Linear(torch.autograd.Function):
    """out = act([a0 . w0^T + b0 | a1 . w1^T]) -- one launch, one or two column ranges (nn.Linear / the concat-with-self)."""

    @staticmethod
    def forward(ctx, act, a0, w0, b0, a1, w1):
        a0 = a0.contiguous()
        segs = [dict(a=a0, w=w0, bias=b0, col0=0)]
        if a1 is not None:
            a1 = a1.contiguous()
            segs.append(dict(a=a1, w=w1, col0=w0.shape[0]))
        out = ops.linear(segs, a0.shape[0], act=act)
        ctx.act, ctx.has_bias = act, b0 is not None
        ctx.save_for_backward(a0, w0, a1, w1, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        a0, w0, a1, w1, out = ctx.saved_tensors
        n, O0 = out.shape[0], w0.shape[0]
        dpre = ops.act_backward(dout.contiguous(), out, ctx.act)
        need = ctx.needs_input_grad
        g0 = dpre[:, :O0]
        da0 = ops.linear([dict(a=g0, w=w0, w_trans=True)], n) if need[1] else None
        dw0 = ops.wgrad(g0, a0) if need[2] else None
        db0 = ops.colsum(g0) if (ctx.has_bias and need[3]) else None
        da1 = dw1 = None
        if a1 is not None:
            g1 = dpre[:, O0:]
            da1 = ops.linear([dict(a=g1, w=w1, w_trans=True)], n) if need[4] else None
            dw1 = ops.wgrad(g1, a1) if need[5] else None
        return None, da0, dw0, db0, da1, dw1


class _SegmentReduce(torch.autograd.Function):
    """rows (n*S, d) grouped by parent -> (n, d): mean or max over the S rows (nn_modules.py:197-198, 225-226, 240, 252)."""

    @staticmethod
    def forward(ctx, rows, S, mode):
        rows = rows.contiguous()
        n = rows.shape[0] // S
        out = ops.gather_reduce(rows, None, n, S, mode, out_dtype=torch.float32)
        ctx.S, ctx.mode = S, mode
        ctx.save_for_backward(rows)
        return out

    @staticmethod
    def backward(ctx, dout):
        rows, = ctx.saved_tensors
        dout = dout.contiguous()
        if ctx.mode == 'mean':
            return ops.segment_broadcast(dout, ctx.S, 1.0 / ctx.S), None, None
        return ops.segment_max_backward(rows, dout, ctx.S), None, None


class _AttentionSum(torch.autograd.Function):
    """m_p = sum_j softmax_j(<na_pj, xa_p>) n_pj (nn_modules.py:309-315).  The gradient reaches `neibs` twice: directly (here)
    and through na = att(neibs), which autograd routes through the _Linear calls that produced it."""

    @staticmethod
    def forward(ctx, neibs, na, xa, S):
        neibs, na, xa = neibs.contiguous(), na.contiguous(), xa.contiguous()
        n = xa.shape[0]
        w = ops.attention_weights(na, xa, n, S)
        out = ops.gather_reduce(neibs, None, n, S, 'sum', weights=w, out_dtype=torch.float32)
        ctx.S = S
        ctx.save_for_backward(neibs, na, xa, w)
        return out

    @staticmethod
    def backward(ctx, dm):
        neibs, na, xa, w = ctx.saved_tensors
        dn, dna, dxa = ops.attention_sum_backward(neibs, dm.contiguous(), w, na, xa, ctx.S)
        return dn, dna, dxa, None


class _Embedding(torch.autograd.Function):
    """weight[ids] (nn.Embedding, nn_modules.py:146,149) with its dense gradient."""

    @staticmethod
    def forward(ctx, weight, ids):
        ctx.rows = weight.shape[0]
        ctx.save_for_backward(ids)
        return ops.gather_rows(weight, ids, out_dtype=torch.float32)

    @staticmethod
    def backward(ctx, drows):
        ids, = ctx.saved_tensors
        return ops.embedding_backward(drows.contiguous(), ids, ctx.rows), None


class _L2Normalize(torch.autograd.Function):
    """F.normalize(z, dim=1) (models.py:90)."""

    @staticmethod
    def forward(ctx, z):
        z = z.contiguous()
        ctx.save_for_backward(z)
        return ops.l2_normalize(z)

    @staticmethod
    def backward(ctx, dzn):
        z, = ctx.saved_tensors
        return ops.l2_normalize_backward(z, dzn)


class _LSTMLast(torch.autograd.Function):
    """The S neighbour rows of every parent through a one-layer LSTM in their sampled order; the last hidden state is the
    aggregate (nn_modules.py:276-279).  Forward = one projection x . W_ih^T over all steps + per step h . W_hh^T and
    gsage_lstm_cell; backward = back-propagation through time over the saved states (gsage_lstm_cell_backward per step, then
    ONE weight-gradient reduction per matrix over all steps).  Everything fp32; states are kept time-major."""

    @staticmethod
    def forward(ctx, neibs, w_ih, w_hh, b_ih, b_hh, S):
        neibs = neibs.contiguous()
        n, H, dev = neibs.shape[0] // S, w_hh.shape[1], neibs.device
        gx = ops.linear([dict(a=neibs, w=w_ih)], n * S).view(n, S, 4 * H)         # step t of parent p is row p*S + t
        hs = torch.zeros((S + 1, n, H), dtype=torch.float32, device=dev)          # hs[t + 1] = h_t, hs[0] = the zero initial state
        cs = torch.zeros((S + 1, n, H), dtype=torch.float32, device=dev)
        gh = torch.empty((n, 4 * H), dtype=torch.float32, device=dev)
        c = torch.empty((n, H), dtype=torch.float32, device=dev)
        b_ih, b_hh = b_ih.contiguous(), b_hh.contiguous()
        for t in range(S):
            if t > 0:
                ops.linear([dict(a=hs[t], w=w_hh)], n, out=gh)
            ops.lstm_cell(gx[:, t], gh if t > 0 else None, b_ih, b_hh, c, hs[t + 1], first=(t == 0))
            cs[t + 1].copy_(c)
        ctx.S = S
        ctx.save_for_backward(neibs, w_ih, w_hh, b_ih, b_hh, gx, hs, cs)
        return hs[S].clone()

    @staticmethod
    def backward(ctx, dh_last):
        neibs, w_ih, w_hh, b_ih, b_hh, gx, hs, cs = ctx.saved_tensors
        S = ctx.S
        n, H, dev = hs.shape[1], hs.shape[2], hs.device
        dg = torch.empty((S, n, 4 * H), dtype=torch.float32, device=dev)            # time-major: pairs with hs[:S] for dW_hh
        gh = torch.empty((n, 4 * H), dtype=torch.float32, device=dev)
        dh = dh_last.contiguous().clone()
        dc = torch.zeros((n, H), dtype=torch.float32, device=dev)
        for t in range(S - 1, -1, -1):
            if t > 0:
                ops.linear([dict(a=hs[t], w=w_hh)], n, out=gh)                      # the recurrent gates of step t, recomputed
            ops.lstm_cell_backward(gx[:, t], gh if t > 0 else None, b_ih, b_hh, cs[t] if t > 0 else None, dh, dc, dg[t], first=(t == 0))
            if t > 0:
                dh = ops.linear([dict(a=dg[t], w=w_hh, w_trans=True)], n)           # d loss / d h_{t-1}
        dg_tm = dg.view(S * n, 4 * H)
        dw_hh = ops.wgrad(dg_tm, hs[:S].reshape(S * n, H))
        dg_pm = dg.transpose(0, 1).contiguous().view(n * S, 4 * H)                  # parent-major: row p*S + t, like `neibs`
        dw_ih = ops.wgrad(dg_pm, neibs)
        db = ops.colsum(dg_pm)
        dn = ops.linear([dict(a=dg_pm, w=w_ih, w_trans=True)], n * S) if ctx.needs_input_grad[0] else None
        return dn, dw_ih, dw_hh, db, db.clone(), None


def linear_fn(a, weight, bias=None, act=None):
    """act(a . weight^T + bias) through the library, differentiable."""
    return _Linear.apply(act, a, weight, bias, None, None)


# --
# Samplers

class UniformNeighborSampler(object):
    """Dense 2-D edgelist sampler (nn_modules.py:19-49): `adj[ids][:, randperm(K)][:, :n_samples]`,
    ONE permutation per call shared by all rows, drawn from torch's CPU generator like the reference."""

    def __init__(self, adj):
        self.adj = torch.as_tensor(adj).to(device='cuda', dtype=torch.int64).contiguous()

    def __call__(self, ids, n_samples=-1):
        ids, was_cuda = _cuda_ids(ids)
        K = self.adj.size(1)
        perm = torch.randperm(K).cuda()                      # CPU generator, as nn_modules.py:44
        S = n_samples if n_samples >= 0 else max(0, K + n_samples)
        S = min(S, K)
        out = torch.empty((ids.shape[0], S), dtype=torch.int64, device='cuda')
        check(lib().gsage_sample_dense(ops.ptr(self.adj), self.adj.size(0), K, ops.ptr(ids), ids.shape[0], ops.ptr(perm),
                                       n_samples, ops.ptr(out), ops.stream()))
        return out if was_cuda else out.cpu()


class SparseUniformNeighborSampler(object):
    """Sparse CSR sampler (nn_modules.py:52-101), bit-exact with the reference under the same seed.

    mode='device' (default): draws come from the device-resident MT19937 stream (`rng`, default the global
        one seeded by `set_seeds`) -- no host round trip.
    mode='host': draws are made on the host with the very call the reference makes
        (`np.random.choice(adj.shape[1], (n, S))` on the global numpy stream) and shipped to the kernel."""

    def __init__(self, adj, rng=None, mode='device'):
        if isinstance(adj, GraphCSR):
            self.graph = adj
        else:
            self.graph = GraphCSR.from_scipy(adj)            # asserts sparse.issparse(adj), like :73
        self.adj = adj
        self.rng = rng
        self.mode = mode

    @property
    def degrees(self):
        return self.graph.degrees

    def __call__(self, ids, n_samples=128):
        assert n_samples > 0, 'SparseUniformNeighborSampler: n_samples must be set explicitly'
        ids, was_cuda = _cuda_ids(ids)
        n = ids.shape[0]
        out = torch.empty((n * n_samples,), dtype=torch.int64, device='cuda')
        if self.mode == 'host':
            sel = np.random.choice(self.graph.shape[1], (n, n_samples))          # nn_modules.py:88, global stream
            sel = torch.from_numpy(sel.astype(np.uint32).view(np.int32).reshape(-1)).cuda()
            check(lib().gsage_sample_sparse(self.graph._h, ops.ptr(ids), n, n_samples, ops.ptr(sel), ops.ptr(out), ops.stream()))
        else:
            rng = self.rng or default_rng()
            check(lib().gsage_sample_sparse_rng(self.graph._h, rng._h, ops.ptr(ids), n, n_samples, ops.ptr(out), ops.stream()))
        return out if was_cuda else out.cpu()


sampler_lookup = {
    "uniform_neighbor_sampler": UniformNeighborSampler,
    "sparse_uniform_neighbor_sampler": SparseUniformNeighborSampler,
}


# --
# Preprocessers

class IdentityPrep(nn.Module):
    def __init__(self, input_dim, n_nodes=None):
        super(IdentityPrep, self).__init__()
        self.input_dim = input_dim

    @property
    def output_dim(self):
        return self.input_dim

    def forward(self, ids, feats, layer_idx=0):
        return feats


class NodeEmbeddingPrep(nn.Module):
    """nn_modules.py:126-155: learned (n_nodes+1, 64) table + 64x64 affine; seeds (layer_idx == 0) all look up
    the masked row `n_nodes`."""

    def __init__(self, input_dim, n_nodes, embedding_dim=64):
        super(NodeEmbeddingPrep, self).__init__()
        self.n_nodes = n_nodes
        self.input_dim = input_dim
        self.embedding_dim = embedding_dim
        self.embedding = nn.Embedding(num_embeddings=n_nodes + 1, embedding_dim=embedding_dim)
        self.fc = nn.Linear(embedding_dim, embedding_dim)

    @property
    def output_dim(self):
        return (self.input_dim + self.embedding_dim) if self.input_dim else self.embedding_dim

    def forward(self, ids, feats, layer_idx=0):
        ids, _ = _cuda_ids(ids)
        look = ids if layer_idx > 0 else torch.full_like(ids, self.n_nodes)
        embs = linear_fn(_Embedding.apply(self.embedding.weight, look), self.fc.weight, self.fc.bias)
        if self.input_dim:
            return torch.cat([_f32(feats), embs], dim=1)
        return embs


class LinearPrep(nn.Module):
    def __init__(self, input_dim, n_nodes, output_dim=32):
        super(LinearPrep, self).__init__()
        self.fc = nn.Linear(input_dim, output_dim, bias=False)
        self.output_dim = output_dim

    def forward(self, ids, feats, layer_idx=0):
        return linear_fn(_f32(feats), self.fc.weight)


prep_lookup = {
    "identity": IdentityPrep,
    "node_embedding": NodeEmbeddingPrep,
    "linear": LinearPrep,
}


# --
# Aggregators

_cat = lambda x: torch.cat(x, dim=1)


class AggregatorMixin(object):
    @property
    def output_dim(self):
        tmp = torch.zeros((1, self.output_dim_))
        return self.combine_fn([tmp, tmp]).size(1)

    def _combine(self, x, x_ids, agg, n):
        """act([fc_x(x) | fc_neib(agg)]) -- one launch writing both halves of the concat buffer."""
        if self.combine_fn is not _cat:
            raise NotImplementedError('gsage: only the default concat combine_fn is fused')
        return ops.linear([dict(a=x, ids=x_ids, w=self.fc_x.weight.data, col0=0),
                           dict(a=agg, w=self.fc_neib.weight.data, col0=self.output_dim_)],
                          n, act=_act_name(self.activation))

    def _forward_rows(self, x, neibs):
        """The narrow contract `agg(x, neibs)` (models.py:86), differentiable w.r.t. x, neibs and the parameters."""
        if self.combine_fn is not _cat:
            raise NotImplementedError('gsage: only the default concat combine_fn is fused')
        x, neibs = _f32(x), _f32(neibs)
        S = neibs.size(0) // x.size(0)
        agg = self._reduce_rows(x, neibs, S)
        return _Linear.apply(_act_name(self.activation), x, self.fc_x.weight, None, agg, self.fc_neib.weight)

    def forward_ids(self, table, ids_self, ids_neib, S):
        return self._run(table, ids_self, table, ids_neib, ids_self.shape[0], S)


class MeanAggregator(nn.Module, AggregatorMixin):
    def forward(self, x, neibs):
        """x (N, d), neibs (N*S, d) row-major grouped by parent -> (N, 2*output_dim)."""
        return self._forward_rows(x, neibs)

    def __init__(self, input_dim, output_dim, activation, combine_fn=_cat):
        super(MeanAggregator, self).__init__()
        self.fc_x = nn.Linear(input_dim, output_dim, bias=False)
        self.fc_neib = nn.Linear(input_dim, output_dim, bias=False)
        self.output_dim_ = output_dim
        self.activation = activation
        self.combine_fn = combine_fn

    def _reduce_rows(self, x, neibs, S):
        return _SegmentReduce.apply(neibs, S, 'mean')

    def _run(self, x, x_ids, nb, nb_ids, n, S):
        agg = ops.gather_reduce(nb, nb_ids, n, S, 'mean', d=self.fc_neib.in_features, out_dtype=torch.float32)
        return self._combine(x, x_ids, agg, n)


class PoolAggregator(nn.Module, AggregatorMixin):
    def forward(self, x, neibs):
        """x (N, d), neibs (N*S, d) row-major grouped by parent -> (N, 2*output_dim)."""
        return self._forward_rows(x, neibs)

    def __init__(self, input_dim, output_dim, pool_fn, activation, hidden_dim=512, combine_fn=_cat):
        super(PoolAggregator, self).__init__()
        self.mlp = nn.Sequential(*[nn.Linear(input_dim, hidden_dim, bias=True), nn.ReLU()])
        self.fc_x = nn.Linear(input_dim, output_dim, bias=False)
        self.fc_neib = nn.Linear(hidden_dim, output_dim, bias=False)
        self.output_dim_ = output_dim
        self.activation = activation
        self.pool_fn = pool_fn                          # 'max' | 'mean' (the reference passes lambdas)
        self.combine_fn = combine_fn

    def _reduce_rows(self, x, neibs, S):
        h = linear_fn(neibs, self.mlp[0].weight, self.mlp[0].bias, act='relu')
        return _SegmentReduce.apply(h, S, self.pool_fn)

    def _run(self, x, x_ids, nb, nb_ids, n, S):
        h = ops.linear([dict(a=nb, ids=nb_ids, w=self.mlp[0].weight.data, bias=self.mlp[0].bias.data)], n * S, act='relu')
        agg = ops.gather_reduce(h, None, n, S, self.pool_fn)
        return self._combine(x, x_ids, agg, n)


class MaxPoolAggregator(PoolAggregator):
    def __init__(self, input_dim, output_dim, activation, hidden_dim=512, combine_fn=_cat):
        super(MaxPoolAggregator, self).__init__(input_dim=input_dim, output_dim=output_dim, pool_fn='max',
                                                activation=activation, hidden_dim=hidden_dim, combine_fn=combine_fn)


class MeanPoolAggregator(PoolAggregator):
    def __init__(self, input_dim, output_dim, activation, hidden_dim=512, combine_fn=_cat):
        super(MeanPoolAggregator, self).__init__(input_dim=input_dim, output_dim=output_dim, pool_fn='mean',
                                                 activation=activation, hidden_dim=hidden_dim, combine_fn=combine_fn)


class AttentionAggregator(nn.Module, AggregatorMixin):
    def forward(self, x, neibs):
        """x (N, d), neibs (N*S, d) row-major grouped by parent -> (N, 2*output_dim)."""
        return self._forward_rows(x, neibs)

    def __init__(self, input_dim, output_dim, activation, hidden_dim=32, combine_fn=_cat):
        super(AttentionAggregator, self).__init__()
        self.att = nn.Sequential(*[
            nn.Linear(input_dim, hidden_dim, bias=False),
            nn.Tanh(),
            nn.Linear(hidden_dim, hidden_dim, bias=False),
        ])
        self.fc_x = nn.Linear(input_dim, output_dim, bias=False)
        self.fc_neib = nn.Linear(input_dim, output_dim, bias=False)
        self.output_dim_ = output_dim
        self.activation = activation
        self.combine_fn = combine_fn

    def _att(self, a, ids, n):
        t = ops.linear([dict(a=a, ids=ids, w=self.att[0].weight.data)], n, act='tanh')
        return ops.linear([dict(a=t, w=self.att[2].weight.data)], n)

    def _reduce_rows(self, x, neibs, S):
        assert S > 1, 'AttentionAggregator: S must be > 1'
        att = lambda rows: linear_fn(linear_fn(rows, self.att[0].weight, act='tanh'), self.att[2].weight)
        return _AttentionSum.apply(neibs, att(neibs), att(x), S)

    def _run(self, x, x_ids, nb, nb_ids, n, S):
        assert S > 1, 'AttentionAggregator: S must be > 1'
        w = ops.attention_weights(self._att(nb, nb_ids, n * S), self._att(x, x_ids, n), n, S)
        agg = ops.gather_reduce(nb, nb_ids, n, S, 'sum', weights=w, d=self.fc_neib.in_features, out_dtype=torch.float32)
        return self._combine(x, x_ids, agg, n)


class LSTMAggregator(nn.Module, AggregatorMixin):
    """nn_modules.py:259-286: the S neighbour rows of every parent through a one-layer nn.LSTM in their sampled order; the LAST
    hidden state is the aggregate.  `self.lstm` is a stock nn.LSTM used as the parameter holder (same state_dict keys, same
    init order as the reference); the recurrence runs on the library's kernels: per step two projections + gsage_lstm_cell."""

    def forward(self, x, neibs):
        """x (N, d), neibs (N*S, d) row-major grouped by parent -> (N, 2*output_dim)."""
        return self._forward_rows(x, neibs)

    def __init__(self, input_dim, output_dim, activation, hidden_dim=512, bidirectional=False, combine_fn=_cat):
        super(LSTMAggregator, self).__init__()
        assert not hidden_dim % 2, "LSTMAggregator: hiddem_dim % 2 != 0"
        if bidirectional:
            raise NotImplementedError('gsage: LSTMAggregator(bidirectional=True) is not built (the registry default is False)')
        self.lstm = nn.LSTM(input_dim, hidden_dim, bidirectional=False, batch_first=True)
        self.fc_x = nn.Linear(input_dim, output_dim, bias=False)
        self.fc_neib = nn.Linear(hidden_dim, output_dim, bias=False)
        self.output_dim_ = output_dim
        self.activation = activation
        self.combine_fn = combine_fn

    def _reduce_rows(self, x, neibs, S):
        return _LSTMLast.apply(neibs, self.lstm.weight_ih_l0, self.lstm.weight_hh_l0, self.lstm.bias_ih_l0, self.lstm.bias_hh_l0, S)

    def _run(self, x, x_ids, nb, nb_ids, n, S):
        H, dev = self.lstm.hidden_size, x.device
        w_ih, w_hh = self.lstm.weight_ih_l0.data, self.lstm.weight_hh_l0.data
        b_ih, b_hh = self.lstm.bias_ih_l0.data.contiguous(), self.lstm.bias_hh_l0.data.contiguous()
        # x . W_ih^T for all S steps of every parent in one projection; step t of parent p is row p*S + t
        gx = ops.linear([dict(a=nb, ids=nb_ids, w=w_ih)], n * S).view(n, S, 4 * H)
        gh = torch.empty((n, 4 * H), dtype=torch.float32, device=dev)
        c = torch.empty((n, H), dtype=torch.float32, device=dev)
        h = torch.empty_like(c)
        for t in range(S):
            if t > 0:
                ops.linear([dict(a=h, w=w_hh)], n, out=gh)
            ops.lstm_cell(gx[:, t], gh if t > 0 else None, b_ih, b_hh, c, h, first=(t == 0))
        return self._combine(x, x_ids, h, n)


aggregator_lookup = {
    "mean": MeanAggregator,
    "max_pool": MaxPoolAggregator,
    "mean_pool": MeanPoolAggregator,
    "attention": AttentionAggregator,
    "lstm": LSTMAggregator,
}
