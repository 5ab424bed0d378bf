"""
The problem side of the reference's training script (/root/reference/problem.py), device-resident:

  * `iterate`        -- the batch iterator `NodeProblem.iterate` (problem.py:141-153) with the per-epoch shuffle drawn from the
                        DEVICE MT19937 stream (`gsage_rng_permutation`), i.e. from the same global stream, at the same position,
                        the reference's `np.random.permutation` would use -- an epoch of seed batches followed by the sampler's
                        draws stays bit-exact with the reference without a host round trip;
  * `ProblemLosses`  -- problem.py:26-41 (stock torch losses on the logits);
  * `ProblemMetrics` -- problem.py:44-64 computed ON THE DEVICE (gsage_metric_f1 / gsage_metric_mae, csrc/metrics.cu);
  * `NodeProblem`    -- the fields and methods train.py touches (problem.py:74-153) over in-memory arrays.
"""

import numpy as np
import torch
from torch.nn import functional as _F

from . import ops as _ops
from ._lib import check as _check, lib as _lib_handle
from .graph import GraphCSR as _GraphCSR
from .rng import default_rng


def iterate(nodes, targets, batch_size=512, shuffle=False, rng=None):
    """Yields (ids, targets, progress) like NodeProblem.iterate: `nodes` = the fold's node ids (int64, numpy or torch),
    chunks follow np.array_split(idx, n // batch_size + 1) exactly; ids/targets are CUDA tensors."""
    nodes = torch.as_tensor(nodes).to(device='cuda', dtype=torch.int64)
    targets = torch.as_tensor(targets).cuda()
    n = nodes.shape[0]
    if shuffle:
        idx = (rng or default_rng()).permutation(n)                  # np.random.permutation(np.arange(n))
    else:
        idx = torch.arange(n, device='cuda')
    n_chunks = n // batch_size + 1
    bounds = np.cumsum([0] + [len(c) for c in np.array_split(np.empty(n, dtype=np.int8), n_chunks)])
    for chunk_id in range(n_chunks):
        mids = nodes[idx[bounds[chunk_id]:bounds[chunk_id + 1]]]
        yield mids, targets[mids], chunk_id / n_chunks


# --------------------------------------------------------------------------------------------------------------------
# NodeProblem: the fields and methods of /root/reference/problem.py:74-153 that train.py touches, over in-memory arrays
# (the h5 container itself is out of scope: h5py is not part of this stack; `problem.h5`'s keys map 1:1 onto the
# arguments of `NodeProblem.from_arrays`, and the sparse adjacency travels as the file's 3 x nnz [v; r; c] array).
# --------------------------------------------------------------------------------------------------------------------
class ProblemLosses(object):
    """problem.py:26-41 -- stock torch losses on the logits (the loss itself is out of scope; its gradient w.r.t. the logits
    is what enters the library's backward pass)."""

    @staticmethod
    def multilabel_classification(preds, targets):
        return _F.multilabel_soft_margin_loss(preds, targets)

    @staticmethod
    def classification(preds, targets):
        return _F.cross_entropy(preds, targets)

    @staticmethod
    def regression_mae(preds, targets):
        return _F.l1_loss(preds, targets.view_as(preds))


def _dev(x, dtype):
    return torch.as_tensor(x).to(device='cuda', dtype=dtype).contiguous()


class ProblemMetrics(object):
    """problem.py:44-64 on the device (gsage_metric_f1 / gsage_metric_mae): same arguments (y_true, y_pred) -- CUDA tensors
    stay where they are, numpy arrays are uploaded -- same return values (a {"micro", "macro"} dict of floats, or a float).
    Only the two scalars cross PCIe, not the (B, n_classes) predictions train.py:150 ships to sklearn every batch."""

    @staticmethod
    def _f1(y_true, y_pred, multilabel):
        preds = _dev(y_pred, torch.float32)
        preds = preds.view(preds.shape[0], -1)
        n, C = preds.shape
        if multilabel:
            tgt = _dev(y_true, torch.float32).view(n, C)
            ld_t = C
        else:
            tgt = _dev(y_true, torch.int64).view(-1)
            assert tgt.shape[0] == n, 'ProblemMetrics: one target per prediction row'
            ld_t = 1
        scratch = torch.empty((3 * C,), dtype=torch.int64, device='cuda')
        out = torch.empty((2,), dtype=torch.float64, device='cuda')
        _ops._bind_device(preds)
        _check(_lib_handle().gsage_metric_f1(_ops.ptr(preds), preds.stride(0), _ops.ptr(tgt), ld_t, n, C, 1 if multilabel else 0,
                                             _ops.ptr(scratch), _ops.ptr(out), _ops.stream()))
        micro, macro = out.tolist()
        return {"micro": float(micro), "macro": float(macro)}

    @staticmethod
    def multilabel_classification(y_true, y_pred):
        return ProblemMetrics._f1(y_true, y_pred, True)

    @staticmethod
    def classification(y_true, y_pred):
        return ProblemMetrics._f1(y_true, y_pred, False)

    @staticmethod
    def regression_mae(y_true, y_pred):
        a, b = _dev(y_pred, torch.float32).view(-1), _dev(y_true, torch.float32).view(-1)
        assert a.shape == b.shape, 'ProblemMetrics.regression_mae: shapes differ'
        out = torch.empty((1,), dtype=torch.float64, device='cuda')
        _ops._bind_device(a)
        _check(_lib_handle().gsage_metric_mae(_ops.ptr(a), _ops.ptr(b), a.shape[0], _ops.ptr(out), _ops.stream()))
        return float(out.item())


class NodeProblem(object):
    """The object train.py drives (problem.py:74-153): `.feats .adj .train_adj .targets .folds .nodes .n_nodes .feats_dim
    .n_classes .task .loss_fn .metric_fn` and `iterate(mode, batch_size, shuffle)`.  Differences, all on the device side of
    the boundary: a sparse adjacency is a device `GraphCSR` (built from the file's [v; r; c] triplets or a scipy matrix)
    instead of a host scipy matrix; `feats` / `targets` live on the GPU; the shuffle of `iterate` is drawn from the device
    MT19937 stream at the position `np.random.permutation` (problem.py:146) would use."""

    def __init__(self, task, n_classes, feats, folds, targets, adj, train_adj, sparse=True, feats_dtype=torch.float32):
        from .model import FeatureTable
        self.task = task.decode() if isinstance(task, bytes) else str(task)
        self.n_classes = int(n_classes) if n_classes is not None else 1
        self.folds = np.asarray(folds)
        if self.folds.dtype.kind == 'S':
            self.folds = self.folds.astype(str)

        def graph(a):
            if isinstance(a, _GraphCSR):
                return a
            a_np = np.asarray(a) if not hasattr(a, 'tocsr') else None
            if a_np is not None and a_np.ndim == 2 and a_np.shape[0] == 3 and sparse:
                return _GraphCSR.from_triplets(a_np)                               # parse_csr_matrix, problem.py:70-72
            return _GraphCSR.from_scipy(a)

        if sparse:
            self.adj = graph(adj)
            self.train_adj = self.adj if train_adj is adj else graph(train_adj)
            self.n_nodes = self.adj.shape[0]
        else:                                                       # dense (N + 1, K) edge-list tables, problem.py:110-116
            self.adj = torch.as_tensor(np.asarray(adj)).to(device='cuda', dtype=torch.int64)
            self.train_adj = self.adj if train_adj is adj else torch.as_tensor(np.asarray(train_adj)).to(device='cuda', dtype=torch.int64)
            self.n_nodes = self.adj.shape[0]
        self.feats = FeatureTable(feats, feats_dtype) if feats is not None else None
        self.feats_dim = self.feats.d if self.feats is not None else None
        tdtype = torch.int64 if self.task == 'classification' else torch.float32
        self.targets = torch.as_tensor(np.asarray(targets)).to(device='cuda', dtype=tdtype)
        self.cuda = True
        self.nodes = {mode: np.where(self.folds == mode)[0] for mode in ('train', 'val', 'test')}
        self.loss_fn = getattr(ProblemLosses, self.task)
        self.metric_fn = getattr(ProblemMetrics, self.task)

    from_arrays = classmethod(lambda cls, **kw: cls(**kw))

    def iterate(self, mode, batch_size=512, shuffle=False):
        return iterate(self.nodes[mode], self.targets, batch_size=batch_size, shuffle=shuffle)
