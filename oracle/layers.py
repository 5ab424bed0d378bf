"""
oracle/layers.py -- TEST INFRASTRUCTURE ONLY (not a product path).

CPU restatement (torch CPU tensors, any float dtype -- fp32 to mirror the reference, fp64 as the
"exact" yardstick tolerances are calibrated against) of the reference's prep / aggregator maths
and of the `GSSupervised.forward` layer loop.  Weights are passed explicitly, keyed by the
reference's own `state_dict` names, so a reference `state_dict` drives this file unchanged.

  * `prep_identity`        /root/reference/nn_modules.py:122-123
  * `prep_linear`          /root/reference/nn_modules.py:165-166
  * `prep_node_embedding`  /root/reference/nn_modules.py:144-155
  * `agg_mean`             /root/reference/nn_modules.py:196-204
  * `agg_pool`             /root/reference/nn_modules.py:223-232 (+ :240 max, :252 mean)
  * `agg_attention`        /root/reference/nn_modules.py:305-321
  * `forward_stack`        /root/reference/models.py:71-91

Pinned by tests/test_oracle_golden.py against fixtures made by the reference's own modules
(tests/golden/make_golden.py).
"""

import torch


def _act(name):
    if name in (None, 'identity', 'none'):
        return lambda t: t
    if name == 'relu':
        return torch.relu
    raise ValueError(name)


# -- preps ------------------------------------------------------------------------------------

def prep_identity(ids, feats, params=None, layer_idx=0, prefix='prep.'):
    return feats


def prep_linear(ids, feats, params, layer_idx=0, prefix='prep.'):
    return feats @ params[prefix + 'fc.weight'].t()


def prep_node_embedding(ids, feats, params, layer_idx=0, prefix='prep.', n_nodes=None):
    """layer_idx == 0: every row looks up the SAME masked row `n_nodes` (the seed never sees its own
    embedding); deeper hops look up their own id.  Then a 64x64 affine; concatenated after feats."""
    table = params[prefix + 'embedding.weight']
    if n_nodes is None:
        n_nodes = table.shape[0] - 1
    look = ids if layer_idx > 0 else torch.full_like(ids, n_nodes)
    emb = table[look] @ params[prefix + 'fc.weight'].t() + params[prefix + 'fc.bias']
    return emb if feats is None else torch.cat([feats, emb], dim=1)


PREPS = {'identity': prep_identity, 'linear': prep_linear, 'node_embedding': prep_node_embedding}


# -- aggregators --------------------------------------------------------------------------------
# contract: x (N, d); neibs (N*S, d) row-major grouped by parent; returns (N, 2*O)

def _combine(x, agg, params, prefix, act):
    out = torch.cat([x @ params[prefix + 'fc_x.weight'].t(), agg @ params[prefix + 'fc_neib.weight'].t()], dim=1)
    return _act(act)(out)


def agg_mean(x, neibs, params, prefix, act):
    m = neibs.reshape(x.shape[0], -1, neibs.shape[1]).mean(dim=1)   # dummy (zero) rows count in S
    return _combine(x, m, params, prefix, act)


def agg_pool(x, neibs, params, prefix, act, reducer='max'):
    h = torch.relu(neibs @ params[prefix + 'mlp.0.weight'].t() + params[prefix + 'mlp.0.bias'])
    h = h.reshape(x.shape[0], -1, h.shape[1])
    p = h.max(dim=1)[0] if reducer == 'max' else h.mean(dim=1)
    return _combine(x, p, params, prefix, act)


def agg_attention(x, neibs, params, prefix, act):
    """a(v) = W2 tanh(W1 v); s_ij = <a(n_ij), a(x_i)>; softmax over j; weighted neighbour sum.
    Dummy rows score exactly 0 (not -inf).  S == 1 is out of contract (the reference's squeeze
    collapses the axis and softmaxes over the batch)."""
    w1, w2 = params[prefix + 'att.0.weight'], params[prefix + 'att.2.weight']
    att = lambda v: torch.tanh(v @ w1.t()) @ w2.t()
    n = x.shape[0]
    na = att(neibs).reshape(n, -1, w2.shape[0])
    xa = att(x)
    assert na.shape[1] > 1, "attention aggregator: S must be > 1"
    score = torch.einsum('nsh,nh->ns', na, xa)
    w = torch.softmax(score, dim=1)
    m = (neibs.reshape(n, -1, neibs.shape[1]) * w.unsqueeze(-1)).sum(dim=1)
    return _combine(x, m, params, prefix, act)


def agg_lstm(x, neibs, params, prefix, act):
    """nn_modules.py:259-286: a one-layer unidirectional nn.LSTM (batch_first) over the S neighbour rows of every parent in
    their sampled order, zero initial state, LAST hidden state -> fc_neib.  torch's cell (gate order i, f, g, o in the 4H rows
    of weight_ih / weight_hh): c' = sigmoid(f) c + sigmoid(i) tanh(g), h' = sigmoid(o) tanh(c')."""
    w_ih, w_hh = params[prefix + 'lstm.weight_ih_l0'], params[prefix + 'lstm.weight_hh_l0']
    b = params[prefix + 'lstm.bias_ih_l0'] + params[prefix + 'lstm.bias_hh_l0']
    n, H = x.shape[0], w_hh.shape[1]
    seq = neibs.reshape(n, -1, neibs.shape[1])
    h = torch.zeros((n, H), dtype=x.dtype)
    c = torch.zeros((n, H), dtype=x.dtype)
    for t in range(seq.shape[1]):
        gates = seq[:, t] @ w_ih.t() + h @ w_hh.t() + b
        i, f, g, o = gates[:, :H], gates[:, H:2 * H], gates[:, 2 * H:3 * H], gates[:, 3 * H:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
    return _combine(x, h, params, prefix, act)


AGGREGATORS = {
    'lstm': agg_lstm,
    'mean': agg_mean,
    'max_pool': lambda *a, **k: agg_pool(*a, reducer='max', **k),
    'mean_pool': lambda *a, **k: agg_pool(*a, reducer='mean', **k),
    'attention': agg_attention,
}


# -- layer loop -----------------------------------------------------------------------------------

def forward_stack(hop_ids, feats, params, aggregator='mean', prep='identity', acts=('relu', 'identity'),
                  n_nodes=None, return_intermediates=False):
    """`hop_ids` = [ids0 (B,), ids1 (B*S1,), ids2 (B*S1*S2,)] already sampled (torch int64).
    Gathers + preps every hop, applies layer k to every adjacent pair, L2-normalises, final fc."""
    prep_fn = PREPS[prep]
    agg_fn = AGGREGATORS[aggregator]
    kw = {'n_nodes': n_nodes} if prep == 'node_embedding' else {}
    levels = []
    for k, ids in enumerate(hop_ids):
        rows = feats[ids] if feats is not None else None
        levels.append(prep_fn(ids, rows, params, layer_idx=k, **kw))
    trace = {'prep': list(levels)}
    for layer, act in enumerate(acts):
        prefix = 'agg_layers.%d.' % layer
        levels = [agg_fn(levels[k], levels[k + 1], params, prefix, act) for k in range(len(levels) - 1)]
        trace['layer%d' % layer] = list(levels)
    assert len(levels) == 1
    z = levels[0]
    z = z / z.norm(dim=1, keepdim=True).clamp_min(1e-12)      # F.normalize(dim=1), eps 1e-12
    out = z @ params['fc.weight'].t() + params['fc.bias']
    if return_intermediates:
        trace['normalized'] = z
        return out, trace
    return out
