"""
GSSupervised -- the reference's model class (/root/reference/models.py:21-104) over the fused engine.

Same constructor signature, same submodule / parameter names (`prep.*`, `agg_layers.k.*`, `fc.*`), so a
reference `state_dict` loads with `load_state_dict` and vice versa.  `forward(ids, feats, train)` returns the
same logits; what differs is *who calls whom*: the reference's Python loop gathers rows and hands them to
the aggregators, here the ids go to the library's engine (gsage_engine_forward) which samples, gathers,
aggregates and projects on the device.  `forward_reference_order` keeps the reference's call order through
the narrow operator API for parity debugging.
"""

import ctypes as C
import weakref
from functools import partial

import torch
from torch import nn
from torch.nn import functional as F

from . import _lib, ops
from ._lib import check, lib
from .graph import GraphCSR
from .operators import (AttentionAggregator, LSTMAggregator, MeanAggregator, MeanPoolAggregator, MaxPoolAggregator, NodeEmbeddingPrep,
                        IdentityPrep, LinearPrep, SparseUniformNeighborSampler, UniformNeighborSampler, _act_name)
from .rng import default_rng
from .lr import LRSchedule


def _agg_name(cls):
    cls = getattr(cls, 'func', cls)               # functools.partial(LSTMAggregator, hidden_dim=...) and the like
    for name, k in (('mean', MeanAggregator), ('max_pool', MaxPoolAggregator), ('mean_pool', MeanPoolAggregator),
                    ('attention', AttentionAggregator), ('lstm', LSTMAggregator)):
        if cls is k:
            return name
    raise ValueError('gsage engine: unsupported aggregator class %r' % (cls,))


def _prep_name(cls):
    for name, k in (('identity', IdentityPrep), ('node_embedding', NodeEmbeddingPrep), ('linear', LinearPrep)):
        if cls is k:
            return name
    raise ValueError('gsage engine: unsupported prep class %r' % (cls,))


class FeatureTable(object):
    """`problem.feats` on the device in the layout the kernels want: 32-byte-aligned zero-padded rows,
    fp32 or bf16.  `.view` is the logical (rows, d) tensor."""

    def __init__(self, feats, dtype=torch.float32):
        self.store, self.d = ops.pad_table(feats, dtype)
        self.view = self.store[:, :self.d]
        self.ld = self.store.shape[1]
        self.rows = self.store.shape[0]
        self.dtype = dtype

    @classmethod
    def from_store(cls, store, d):
        """Wrap a device tensor that already has the layout (rows, ld >= d, zero padded, 32-byte rows)."""
        assert store.is_cuda and store.dim() == 2 and store.stride(1) == 1 and (store.stride(0) * store.element_size()) % 32 == 0
        self = cls.__new__(cls)
        self.store, self.d = store, d
        self.view = store[:, :d]
        self.ld, self.rows, self.dtype = store.stride(0), store.shape[0], store.dtype
        return self


class GSSupervised(nn.Module):
    def __init__(self, input_dim, n_nodes, n_classes, layer_specs, aggregator_class, prep_class, sampler_class,
                 adj, train_adj, lr_init=0.01, weight_decay=0.0, lr_schedule='constant', epochs=10,
                 compute_dtype=torch.float32, max_batch=512, rng=None, allow_tf32=False):
        super(GSSupervised, self).__init__()
        # the fused engine implements the two-layer stack train.py:105-118 hard-codes; any other depth (models.py:50-62,85-86 loop over
        # whatever `layer_specs` holds) runs the reference's dataflow over the plug-in operators (forward_reference_order)
        self.fused = len(layer_specs) == 2
        self.layer_specs = layer_specs
        self.n_nodes, self.n_classes = n_nodes, n_classes
        self.compute_dtype = compute_dtype
        self.max_batch = max_batch
        self.allow_tf32 = allow_tf32          # fp32 tables, projections as TF32 on the tensor cores (~1e-3 relative)
        self.rng = rng

        # Sampler (models.py:41-44) -- one device graph per adjacency, shared when they are the same object
        self.train_sampler = sampler_class(adj=train_adj)
        self.val_sampler = self.train_sampler if adj is train_adj else sampler_class(adj=adj)
        self.train_sample_fns = [partial(self.train_sampler, n_samples=s['n_train_samples']) for s in layer_specs]
        self.val_sample_fns = [partial(self.val_sampler, n_samples=s['n_val_samples']) for s in layer_specs]

        # Prep (models.py:47-48)
        self.prep = prep_class(input_dim=input_dim, n_nodes=n_nodes)
        self.input_dim = input_dim
        input_dim = self.prep.output_dim

        # Network (models.py:50-62)
        agg_layers = []
        for spec in layer_specs:
            agg = aggregator_class(input_dim=input_dim, output_dim=spec['output_dim'], activation=spec['activation'])
            agg_layers.append(agg)
            input_dim = agg.output_dim
        self.agg_layers = nn.Sequential(*agg_layers)
        self.fc = nn.Linear(input_dim, n_classes, bias=True)

        self._agg_name, self._prep_name = _agg_name(aggregator_class), _prep_name(prep_class)
        self._engines = {}
        self._tables = {}
        self._last = None
        self.last_loss = None

        # Optimiser (models.py:64-69): the schedule named by `lr_schedule` with `lr_init` bound, its value at progress 0, and an
        # Adam over all parameters.  The Adam is parallel.FusedAdam -- clip_grad_norm 5 + the Adam update as one native call
        # over flat buffers (models.py:102-103) -- built lazily because the reference constructs it before `model.cuda()`
        from .parallel import FusedAdam
        self.lr_scheduler = partial(getattr(LRSchedule, lr_schedule), lr_init=lr_init)
        self.lr = self.lr_scheduler(0.0)
        self.optimizer = FusedAdam(self, lr=self.lr, weight_decay=weight_decay, lazy=True)

    def set_progress(self, progress):
        """models.py:93-95: evaluate the schedule at `progress` (epochs) and hand the rate to the optimiser."""
        self.lr = self.lr_scheduler(progress)
        LRSchedule.set_lr(self.optimizer, self.lr)

    # -- engine plumbing --------------------------------------------------------------------------
    def _table(self, feats):
        if feats is None:
            return None
        if isinstance(feats, FeatureTable):
            return feats
        # the padded device copy is cached per tensor OBJECT and content version: an in-place update of `feats` (torch bumps
        # `_version`) or a new tensor at a recycled address gets a fresh copy, a dead tensor drops its entry
        key = (id(feats), self.compute_dtype)
        version = feats._version if torch.is_tensor(feats) else None
        hit = self._tables.get(key)
        if hit is not None and hit[0]() is feats and hit[1] == version:
            return hit[2]
        table = FeatureTable(feats, self.compute_dtype)
        try:
            ref = weakref.ref(feats, lambda _r, k=key, t=self._tables: t.pop(k, None))
        except TypeError:                                      # numpy arrays of some subclasses: keep the object alive instead
            ref = (lambda o: (lambda: o))(feats)
        self._tables[key] = (ref, version, table)
        return table

    def _engine(self, table, fanout, batch):
        key = (None if table is None else table.store.data_ptr(), tuple(fanout))
        eng = self._engines.get(key)
        if eng is not None and eng['max_batch'] >= batch:
            return eng
        if eng is not None:
            lib().gsage_engine_destroy(eng['h'])
        ops._bind_device()
        cfg = _lib.EngineConfig()
        cfg.aggregator, cfg.prep, cfg.n_layers = _lib.AGGREGATOR[self._agg_name], _lib.PREP[self._prep_name], 2
        for k, spec in enumerate(self.layer_specs):
            cfg.fanout[k] = int(fanout[k])
            cfg.out_dim[k] = int(spec['output_dim'])
            cfg.act[k] = _lib.ACT[_act_name(spec['activation'])]
        cfg.n_classes = self.n_classes
        cfg.compute_dtype = _lib.BF16 if self.compute_dtype == torch.bfloat16 else _lib.F32
        if table is not None:
            cfg.feats_dev, cfg.feats_dtype = table.store.data_ptr(), ops.dt(table.store)
            cfg.feats_ld, cfg.feats_dim, cfg.feats_rows = table.ld, table.d, table.rows
        emb_keep = None
        if self._prep_name == 'node_embedding':
            emb = self.prep.embedding.weight.data
            if self.compute_dtype == torch.bfloat16:
                emb_keep = FeatureTable(emb, torch.bfloat16)
                cfg.emb_dev, cfg.emb_dtype, cfg.emb_ld = emb_keep.store.data_ptr(), _lib.BF16, emb_keep.ld
            else:
                cfg.emb_dev, cfg.emb_dtype, cfg.emb_ld = emb.data_ptr(), _lib.F32, emb.stride(0)
            cfg.emb_dim, cfg.n_nodes = self.prep.embedding_dim, self.n_nodes
        agg0 = self.agg_layers[0]
        cfg.hidden_dim = (agg0.mlp[0].out_features if hasattr(agg0, 'mlp') else agg0.att[0].out_features if hasattr(agg0, 'att') else
                          agg0.lstm.hidden_size if hasattr(agg0, 'lstm') else 0)
        cfg.max_batch = max(batch, self.max_batch)
        cfg.allow_tf32 = 1 if self.allow_tf32 else 0
        h = C.c_void_p()
        check(lib().gsage_engine_create(C.byref(cfg), C.byref(h)))
        eng = dict(h=h, max_batch=cfg.max_batch, emb_keep=emb_keep, cfg=cfg)
        self._engines[key] = eng
        return eng

    def _reset_engines(self):
        """Drop every engine (they hold raw pointers into parameter / table storage that is about to move)."""
        for eng in self._engines.values():
            lib().gsage_engine_destroy(eng['h'])
        self._engines = {}
        self._last = None

    def _weights(self):
        w = _lib.Weights()
        p = lambda t: t.data.data_ptr()
        for k, agg in enumerate(self.agg_layers.children()):
            for t in agg.parameters():
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), 'GSSupervised: move the model to CUDA (fp32)'
            w.layer[k].fc_x, w.layer[k].fc_neib = p(agg.fc_x.weight), p(agg.fc_neib.weight)
            if hasattr(agg, 'mlp'):
                w.layer[k].mlp_w, w.layer[k].mlp_b = p(agg.mlp[0].weight), p(agg.mlp[0].bias)
            if hasattr(agg, 'att'):
                w.layer[k].att_w1, w.layer[k].att_w2 = p(agg.att[0].weight), p(agg.att[2].weight)
            if hasattr(agg, 'lstm'):
                w.layer[k].lstm_w_ih, w.layer[k].lstm_w_hh = p(agg.lstm.weight_ih_l0), p(agg.lstm.weight_hh_l0)
                w.layer[k].lstm_b_ih, w.layer[k].lstm_b_hh = p(agg.lstm.bias_ih_l0), p(agg.lstm.bias_hh_l0)
        w.fc_w, w.fc_b = p(self.fc.weight), p(self.fc.bias)
        if self._prep_name == 'node_embedding':
            w.prep_fc_w, w.prep_fc_b = p(self.prep.fc.weight), p(self.prep.fc.bias)
        if self._prep_name == 'linear':
            w.prep_fc_w, w.prep_out_dim = p(self.prep.fc.weight), self.prep.output_dim
        return w

    def _push_weights(self, eng):
        """Hand the current parameters to the engine -- only when a parameter moved or was updated in place
        (torch bumps `_version` on every in-place write, e.g. an optimiser step)."""
        stamp = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (getattr(self, '_weights_epoch', 0),)
        if eng.get('stamp') != stamp:
            if eng.get('emb_keep') is not None:                     # bf16 mode: the engine reads a bf16 copy of the learned table
                keep = eng['emb_keep']
                keep.store[:, :keep.d].copy_(self.prep.embedding.weight.data)
            w = self._weights()
            check(lib().gsage_engine_set_weights(eng['h'], C.byref(w), ops.stream()))
            eng['stamp'] = stamp

    # -- the reference's public surface ---------------------------------------------------------------
    def forward(self, ids, feats, train=True, shard=None, keep_activations=False, next_ids=None, next_shard=None):
        """models.py:71-91.  `ids`: int64 tensor of seed ids (CUDA, or CPU -> copied); `feats`: the node feature
        table (tensor / FeatureTable) or None.  Returns fp32 logits (B, n_classes) on the GPU.

        `next_ids` (keyword extra): the batch after this one.  Its draws and CSR lookups are queued on the engine's sampler
        stream right behind this forward (sample-ahead) and ordered after whatever produced `next_ids` on the current stream
        BEFORE this call -- the next forward must then be given that same batch.  Out-of-range ids and rng window faults of
        EARLIER forwards are raised here (IndexError / GsageError) from sticky flags polled without a synchronisation."""
        if not self.fused:
            assert shard is None and next_ids is None, 'GSSupervised: seed sharding / sample-ahead need the two-layer engine'
            with torch.no_grad():
                return self.forward_reference_order(ids, feats, train=train)
        sampler = self.train_sampler if train else self.val_sampler
        fanout = [s['n_train_samples' if train else 'n_val_samples'] for s in self.layer_specs]
        ids = torch.as_tensor(ids).to(device='cuda', dtype=torch.int64).contiguous().view(-1)
        table = self._table(feats)
        eng = self._engine(table, fanout, max(ids.shape[0], 0 if next_ids is None else int(next_ids.shape[0])))
        check(lib().gsage_engine_poll_errors(eng['h']))
        self._push_weights(eng)
        check(lib().gsage_engine_keep_activations(eng['h'], 1 if keep_activations else 0))
        if next_ids is not None:
            next_ids = torch.as_tensor(next_ids).to(device='cuda', dtype=torch.int64).contiguous().view(-1)
            check(lib().gsage_engine_inputs_ready(eng['h'], ops.stream()))     # next_ids is complete on this stream HERE
        if isinstance(sampler, UniformNeighborSampler):
            # the dense 2-D edgelist sampler (train.py:55's default): one torch.randperm(K) per hop from the CPU generator,
            # hop 0 first -- exactly the draws nn_modules.py:44 makes -- then the same engine
            assert shard is None and next_ids is None, 'GSSupervised: seed sharding / sample-ahead are implemented for the sparse sampler'
            K = sampler.adj.size(1)
            perms = [torch.randperm(K).cuda() for _ in fanout]
            out = torch.empty((ids.shape[0], self.n_classes), dtype=torch.float32, device='cuda')
            check(lib().gsage_engine_forward_dense(eng['h'], ops.ptr(sampler.adj), sampler.adj.size(0), K, ops.ptr(perms[0]), ops.ptr(perms[1]),
                                                   ops.ptr(ids), ids.shape[0], ops.ptr(out), ops.stream()))
            self._last = eng
            return out
        assert isinstance(sampler, SparseUniformNeighborSampler), 'GSSupervised: unknown sampler class %r' % (type(sampler),)
        rng = sampler.rng or self.rng or default_rng()
        out = torch.empty((ids.shape[0], self.n_classes), dtype=torch.float32, device='cuda')
        if shard is None:
            check(lib().gsage_engine_forward(eng['h'], sampler.graph._h, rng._h, ops.ptr(ids), ids.shape[0], ops.ptr(out), ops.stream()))
        else:                                  # (global_batch, first): this rank holds seeds [first, first + len(ids))
            check(lib().gsage_engine_forward_sharded(eng['h'], sampler.graph._h, rng._h, ops.ptr(ids), ids.shape[0], int(shard[0]),
                                                     int(shard[1]), ops.ptr(out), ops.stream()))
        self._last = eng
        if next_ids is not None:
            self.sample_ahead(next_ids, feats, train=train, shard=next_shard)
        return out

    def check(self):
        """Synchronise and raise what the sticky device flags hold: IndexError for an id outside the adjacency (where the
        reference's `feats[ids]` / scipy indexing raise, models.py:76), GsageError for an rng look-ahead fault."""
        torch.cuda.synchronize()
        if self._last is not None:
            check(lib().gsage_engine_poll_errors(self._last['h']))

    def sample_ahead(self, ids, feats, train=True, shard=None, host=False):
        """Draw and sample both hops of the NEXT batch now, on the engine's own stream (gsage_engine_sample_ahead).  The next
        `forward` / `forward_host` must be given the same `ids` tensor.  Draw order on the RNG stream = call order, so
        the sampled ids are bit-identical with and without it.  Called on its own, the sampler stream is ordered after
        everything queued on the current stream so far (always correct; it then starts behind the forward in flight);
        `forward(..., next_ids=)` / `train_step(..., next_ids=)` mark the ids as ready BEFORE they queue their forward, which
        is what lets the sampling overlap it."""
        sampler = self.train_sampler if train else self.val_sampler
        fanout = [s['n_train_samples' if train else 'n_val_samples'] for s in self.layer_specs]
        if not host:
            ids = torch.as_tensor(ids).to(device='cuda', dtype=torch.int64).contiguous().view(-1)
        table = self._table(feats)
        eng = self._engine(table, fanout, ids.shape[0])
        rng = sampler.rng or self.rng or default_rng()
        B = ids.shape[0]
        if host:
            check(lib().gsage_engine_sample_ahead_host(eng['h'], sampler.graph._h, rng._h, C.c_void_p(ids.data_ptr()), B, ops.stream()))
        else:
            gB, first = (B, 0) if shard is None else (int(shard[0]), int(shard[1]))
            check(lib().gsage_engine_sample_ahead(eng['h'], sampler.graph._h, rng._h, ops.ptr(ids), B, gB, first, ops.stream()))
        self._ahead_ids = ids                 # the copy runs later on the sampler stream: keep the source alive
        return ids

    def forward_host(self, ids_host, feats, logits_host, train=True, next_ids_host=None):
        """End-to-end entry for host buffers (pinned numpy / torch CPU tensors): H2D ids, forward, D2H logits.
        `next_ids_host`: the batch after this one -- its H2D copy and sampling are queued before this call blocks on
        its own result (sample-ahead); the next call must then be given that same tensor."""
        sampler = self.train_sampler if train else self.val_sampler
        fanout = [s['n_train_samples' if train else 'n_val_samples'] for s in self.layer_specs]
        table = self._table(feats)
        B = ids_host.shape[0]
        eng = self._engine(table, fanout, max(B, next_ids_host.shape[0] if next_ids_host is not None else 0))
        self._push_weights(eng)
        rng = sampler.rng or self.rng or default_rng()
        nxt = C.c_void_p(next_ids_host.data_ptr()) if next_ids_host is not None else None
        check(lib().gsage_engine_forward_host_next(eng['h'], sampler.graph._h, rng._h, C.c_void_p(ids_host.data_ptr()), B,
                                                   nxt, next_ids_host.shape[0] if next_ids_host is not None else 0,
                                                   C.c_void_p(logits_host.data_ptr()), ops.stream()))
        self._ahead_ids = next_ids_host
        self._last = eng
        return logits_host

    # -- training (models.py:97-104) ---------------------------------------------------------------------
    def _bucket(self):
        from .parallel import FlatGradBucket
        if getattr(self, '_grad_bucket', None) is None:
            head = list(self.fc.parameters()) + list(self.agg_layers[1].parameters())
            self._grad_bucket = FlatGradBucket(self.parameters(), head=head)
            self._grad_bucket.attach()
        return self._grad_bucket

    def has_fused_backward(self):
        """Whether model.backward (the engine's hand-written gradient pass) covers this aggregator x prep x dtype; everything else
        trains through the narrow plug-in API's autograd Functions (train_step picks automatically), the LSTM aggregator included."""
        bf16 = self.compute_dtype == torch.bfloat16
        with_feats = self.input_dim is not None
        if not self.fused:
            return False
        if self._agg_name == 'mean':
            return self._prep_name in ('identity', 'linear') or (self._prep_name == 'node_embedding' and not with_feats and not bf16
                                                                  and not self.allow_tf32)
        if self._agg_name in ('max_pool', 'mean_pool'):
            return bf16 and (self._prep_name == 'identity' or (self._prep_name == 'node_embedding' and not with_feats))
        if self._agg_name == 'attention':
            return bf16 and self._prep_name == 'identity'
        return False

    def backward(self, dlogits, grad_scale=1.0, overlap_stream=None):
        """Parameter gradients of the last forward into the flat bucket (p.grad are views of it), summed over the
        ranks of the default process group.  `dlogits` = d loss / d logits (B, n_classes) fp32 on the GPU.
        `grad_scale` weights this rank's contribution (local_batch / global_batch for a mean loss).
        The head (fc + layer 2) is all-reduced on `overlap_stream` while layer 1's weight gradients are computed."""
        if not self.has_fused_backward():
            raise NotImplementedError('gsage: no fused backward for %s + %s in this dtype; train_step back-propagates through the narrow '
                                      'plug-in API instead (forward_reference_order + loss.backward())' % (self._agg_name, self._prep_name))
        bucket = self._bucket()
        bucket.attach()              # optimizer.zero_grad(set_to_none=True) (torch's default) drops p.grad: point it at the bucket again
        g = _lib.Grads()
        aggs = list(self.agg_layers.children())
        if self._agg_name == 'attention':
            ag = _lib.AttentionGrads()
            for k in range(2):
                g.fc_x[k] = bucket.grad_of(aggs[k].fc_x.weight).data_ptr()
                g.fc_neib[k] = bucket.grad_of(aggs[k].fc_neib.weight).data_ptr()
                ag.att_w1[k] = bucket.grad_of(aggs[k].att[0].weight).data_ptr()
                ag.att_w2[k] = bucket.grad_of(aggs[k].att[2].weight).data_ptr()
            g.fc_w, g.fc_b = bucket.grad_of(self.fc.weight).data_ptr(), bucket.grad_of(self.fc.bias).data_ptr()
            dlogits = dlogits.contiguous().float()
            check(lib().gsage_engine_backward_attention(self._last['h'], ops.ptr(dlogits), C.byref(g), C.byref(ag), ops.stream()))
            bucket.all_reduce(grad_scale)
            return bucket
        if self._agg_name in ('max_pool', 'mean_pool'):
            # one call computes every gradient (gsage_engine_backward_pool); no head / layer-1 split to overlap with
            pg = _lib.PoolGrads()
            for k in range(2):
                g.fc_x[k] = bucket.grad_of(aggs[k].fc_x.weight).data_ptr()
                g.fc_neib[k] = bucket.grad_of(aggs[k].fc_neib.weight).data_ptr()
                pg.mlp_w[k] = bucket.grad_of(aggs[k].mlp[0].weight).data_ptr()
                pg.mlp_b[k] = bucket.grad_of(aggs[k].mlp[0].bias).data_ptr()
            g.fc_w, g.fc_b = bucket.grad_of(self.fc.weight).data_ptr(), bucket.grad_of(self.fc.bias).data_ptr()
            dlogits = dlogits.contiguous().float()
            if self._prep_name == 'node_embedding':
                # BASELINE config C3 (Pokec): the prep's affine is folded into layer 1, so the library returns the raw
                # reductions against the embedding rows and the (O x 64)(64 x 64) unfolding products are done here
                O = aggs[0].fc_x.weight.shape[0]
                csum_x = torch.empty((O,), dtype=torch.float32, device='cuda')
                eg = _lib.PoolEmbeddingGrads()
                eg.csum_x, eg.d_table = csum_x.data_ptr(), bucket.grad_of(self.prep.embedding.weight).data_ptr()
                check(lib().gsage_engine_backward_pool_embedding(self._last['h'], ops.ptr(dlogits), C.byref(g), C.byref(pg), C.byref(eg),
                                                                 ops.stream()))
                Wp, bp = self.prep.fc.weight.data, self.prep.fc.bias.data
                Wx, W1 = aggs[0].fc_x.weight.data, aggs[0].mlp[0].weight.data
                gx_raw = bucket.grad_of(aggs[0].fc_x.weight).clone()
                g1_raw = bucket.grad_of(aggs[0].mlp[0].weight).clone()
                c1 = bucket.grad_of(aggs[0].mlp[0].bias)
                bucket.grad_of(aggs[0].fc_x.weight).copy_(gx_raw @ Wp.t() + torch.outer(csum_x, bp))
                bucket.grad_of(aggs[0].mlp[0].weight).copy_(g1_raw @ Wp.t() + torch.outer(c1, bp))
                bucket.grad_of(self.prep.fc.weight).copy_(Wx.t() @ gx_raw + W1.t() @ g1_raw)
                bucket.grad_of(self.prep.fc.bias).copy_(Wx.t() @ csum_x + W1.t() @ c1)
            else:
                check(lib().gsage_engine_backward_pool(self._last['h'], ops.ptr(dlogits), C.byref(g), C.byref(pg), ops.stream()))
            bucket.all_reduce(grad_scale)
            return bucket
        for k in range(2):
            g.fc_x[k] = bucket.grad_of(aggs[k].fc_x.weight).data_ptr()
            g.fc_neib[k] = bucket.grad_of(aggs[k].fc_neib.weight).data_ptr()
        g.fc_w, g.fc_b = bucket.grad_of(self.fc.weight).data_ptr(), bucket.grad_of(self.fc.bias).data_ptr()
        dlogits = dlogits.contiguous().float()
        check(lib().gsage_engine_backward_head(self._last['h'], ops.ptr(dlogits), C.byref(g), ops.stream()))
        main = torch.cuda.current_stream()
        if bucket.one_shot:
            overlap_stream = None            # one launch for the whole bucket after the last gradient kernel: nothing to split
        if overlap_stream is not None:
            overlap_stream.wait_stream(main)
            with torch.cuda.stream(overlap_stream):
                bucket.all_reduce_head(grad_scale)
        if self._prep_name == 'node_embedding':
            self._backward_layer1_embedding(bucket, aggs[0])
        elif self._prep_name == 'linear':
            self._backward_layer1_linear(bucket, aggs[0])
        else:
            check(lib().gsage_engine_backward_layer1(self._last['h'], C.byref(g), ops.stream()))
        if overlap_stream is None:
            bucket.all_reduce(grad_scale)
        else:
            bucket.all_reduce_tail(grad_scale)
            main.wait_stream(overlap_stream)
        return bucket

    def _backward_layer1_linear(self, bucket, agg0):
        """mean + LinearPrep (nn_modules.py:158-166): the library reduces G against the raw feature rows (self rows and neighbour
        means); the three small products that turn the two (O x d) reductions into parameter gradients are done here."""
        O, d = agg0.fc_x.weight.shape[0], self.prep.fc.weight.shape[1]
        raw = torch.empty((2, O, d), dtype=torch.float32, device='cuda')
        lg = _lib.LinearPrepGrads()
        lg.gx_raw, lg.gn_raw = raw[0].data_ptr(), raw[1].data_ptr()
        check(lib().gsage_engine_backward_layer1_linear(self._last['h'], C.byref(lg), ops.stream()))
        Wp, Wx, Wn = self.prep.fc.weight.data, agg0.fc_x.weight.data, agg0.fc_neib.weight.data
        bucket.grad_of(agg0.fc_x.weight).copy_(raw[0] @ Wp.t())
        bucket.grad_of(agg0.fc_neib.weight).copy_(raw[1] @ Wp.t())
        bucket.grad_of(self.prep.fc.weight).copy_(Wx.t() @ raw[0] + Wn.t() @ raw[1])

    def _backward_layer1_embedding(self, bucket, agg0):
        """The Pokec recipe (mean + NodeEmbeddingPrep without features): the library reduces over the 26*B parent rows and
        scatters the table gradient; the four (O x 64)(64 x 64) products that turn those reductions into parameter
        gradients are done here (gsage_engine_backward_layer1_embedding documents the formulas)."""
        O, de = agg0.fc_x.weight.shape[0], self.prep.embedding_dim
        raw = torch.empty((2 * O * de + 2 * O,), dtype=torch.float32, device='cuda')
        gx_raw, gn_raw, csum = raw[:O * de].view(O, de), raw[O * de:2 * O * de].view(O, de), raw[2 * O * de:]
        eg = _lib.EmbeddingGrads()
        eg.gx_raw, eg.gn_raw, eg.csum = gx_raw.data_ptr(), gn_raw.data_ptr(), csum.data_ptr()
        eg.d_table = bucket.grad_of(self.prep.embedding.weight).data_ptr()
        check(lib().gsage_engine_backward_layer1_embedding(self._last['h'], C.byref(eg), ops.stream()))
        Wp, bp = self.prep.fc.weight.data, self.prep.fc.bias.data
        Wx, Wn = agg0.fc_x.weight.data, agg0.fc_neib.weight.data
        cx, cn = csum[:O], csum[O:]
        bucket.grad_of(agg0.fc_x.weight).copy_(gx_raw @ Wp.t() + torch.outer(cx, bp))
        bucket.grad_of(agg0.fc_neib.weight).copy_(gn_raw @ Wp.t() + torch.outer(cn, bp))
        bucket.grad_of(self.prep.fc.weight).copy_(Wx.t() @ gx_raw + Wn.t() @ gn_raw)
        bucket.grad_of(self.prep.fc.bias).copy_(Wx.t() @ cx + Wn.t() @ cn)

    def train_step(self, ids, feats, targets, loss_fn, *, optimizer=None, clip=5.0, grad_scale=1.0, overlap_stream=None, shard=None,
                   next_ids=None, next_shard=None):
        """models.py:97-104, same positional signature, same return value (`preds`): zero_grad, forward, loss, backward,
        clip_grad_norm 5, optimiser step.  The loss is the caller's torch function (problem.py:26-41) evaluated on the
        logits; its gradient enters the library's backward pass; clip + Adam are one native call on the model's own
        optimiser (`self.optimizer`, models.py:69).  The scalar loss of the step is kept in `self.last_loss` (device tensor).

        Keyword-only extras: `optimizer` (a torch optimiser to step instead of the model's own, or False for gradients only),
        `clip`, `grad_scale`
        (this rank's local/global batch weight under seed sharding), `overlap_stream` (all-reduce of the head gradients
        under the layer-1 backward), `shard`, `next_ids` / `next_shard` (sample the next batch ahead, see `forward`)."""
        opt = self.optimizer if optimizer is None else optimizer        # optimizer=False: gradients only, no update
        if opt and getattr(opt, 'flat', 0) is None:
            opt._materialize()                 # before the forward: building the flat buffers re-creates the engines
        if opt:
            opt.zero_grad()
        if not self.has_fused_backward():
            # registry combinations the engine has no fused backward for (attention + node_embedding, node_embedding WITH
            # features, pool / attention in fp32 or behind LinearPrep): the reference's own dataflow over the narrow plug-in
            # API, whose calls are autograd Functions over library kernels (operators.py) -- slower (rows are materialised),
            # same draws, same result
            assert shard is None and next_ids is None, 'GSSupervised: seed sharding / sample-ahead need the fused backward'
            bucket = self._bucket()
            bucket.attach()
            bucket.flat.zero_()
            preds = self.forward_reference_order(ids, feats, train=True)
            loss = loss_fn(preds, targets.squeeze())
            loss.backward()                    # p.grad are views of the flat bucket: autograd accumulates into them in place
            bucket.attach()
            if bucket.sym is not None:         # peers read this rank's LOCAL gradient from its symmetric buffer
                bucket.sym[:bucket.flat.numel()].copy_(bucket.flat)
            bucket.all_reduce(grad_scale)
            preds = preds.detach()
        else:
            preds = self(ids, feats, train=True, shard=shard, keep_activations=True, next_ids=next_ids, next_shard=next_shard)
            leaf = preds.detach().requires_grad_(True)
            loss = loss_fn(leaf, targets.squeeze())
            dlogits, = torch.autograd.grad(loss, leaf)
            self.backward(dlogits, grad_scale=grad_scale, overlap_stream=overlap_stream)
        if opt and getattr(opt, 'fused_clip', False):
            opt.step(clip=clip)                                    # clip + Adam in one native call (parallel.FusedAdam)
        else:
            if clip:
                torch.nn.utils.clip_grad_norm_(self.parameters(), clip)
            if opt:
                opt.step()
        self.last_loss = loss.detach()
        return preds

    def profile(self, enable=True):
        """Switch the engine's CUDA-event stopwatch on/off (all engines of this model)."""
        for eng in self._engines.values():
            check(lib().gsage_engine_profile(eng['h'], 1 if enable else 0))

    def profile_read(self):
        """{category: (ms, launches, algorithmic bytes, flops)} accumulated since the last read, for the last-used engine
        (categories: _lib.PROF_CATS); synchronises."""
        n = len(_lib.PROF_CATS)
        ms, cnt, byt, flo = (C.c_double * n)(), (C.c_int64 * n)(), (C.c_double * n)(), (C.c_double * n)()
        check(lib().gsage_engine_profile_read(self._last['h'], ms, cnt, byt, flo, ops.stream()))
        return {name: (ms[i], cnt[i], byt[i], flo[i]) for i, name in enumerate(_lib.PROF_CATS)}

    def peek(self, what):
        """Device view of an intermediate of the last forward: 'ids0' 'ids1' 'ids2' 'layer1' 'layer2'."""
        code = {'ids0': 0, 'ids1': 1, 'ids2': 2, 'layer1': 10, 'layer2': 11}[what]
        ptr, rows, cols, ld, dtype = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        check(lib().gsage_engine_peek(self._last['h'], code, C.byref(ptr), C.byref(rows), C.byref(cols), C.byref(ld), C.byref(dtype)))
        view = ops.device_view(ptr.value, rows.value * ld.value, dtype.value).clone().view(rows.value, ld.value)
        return view[:, :cols.value] if cols.value > 1 else view.view(-1)

    def forward_reference_order(self, ids, feats, train=True):
        """The reference's own loop (models.py:71-91) over the narrow operator API: sample -> feats[ids] -> prep ->
        aggregators on pre-gathered rows.  Slower (rows are materialised), but every step is a plug-in call exactly where the
        reference makes one, and the result is autograd-differentiable like the reference's: `loss.backward()` reaches every
        parameter through the operators' autograd Functions (operators.py)."""
        sample_fns = self.train_sample_fns if train else self.val_sample_fns
        ids = torch.as_tensor(ids).to(device='cuda', dtype=torch.int64).contiguous().view(-1)
        table = self._table(feats)
        rows = lambda i: ops.gather_rows(table.view, i, out_dtype=torch.float32) if table is not None else None
        all_feats = [self.prep(ids, rows(ids), layer_idx=0)]
        for layer_idx, sampler_fn in enumerate(sample_fns):
            ids = sampler_fn(ids=ids).contiguous().view(-1)
            all_feats.append(self.prep(ids, rows(ids), layer_idx=layer_idx + 1))
        for agg_layer in self.agg_layers.children():
            all_feats = [agg_layer(all_feats[k], all_feats[k + 1]) for k in range(len(all_feats) - 1)]
        assert len(all_feats) == 1, "len(all_feats) != 1"
        from .operators import _L2Normalize, linear_fn
        out = _L2Normalize.apply(all_feats[0])
        return linear_fn(out, self.fc.weight, self.fc.bias)

    def __del__(self):
        try:
            for eng in self._engines.values():
                lib().gsage_engine_destroy(eng['h'])
        except Exception:
            pass
