"""
Seed-node data parallelism: one process per GPU, graph + feature table replicated, every rank
samples / gathers / aggregates its own slice of the seed batch, and the ONLY collective is a sum all-reduce of the
parameter gradients, carried in one flat bucket so it is a single NCCL call (two when the head is overlapped with
the layer-1 weight gradients).  Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""

import torch
import torch.distributed as dist

from . import ops
from ._lib import check, lib


def shard_seeds(ids, rank, world):
    """This rank's contiguous slice of the (already shuffled) global seed batch; slices partition the batch and
    differ in length by at most one (np.array_split semantics, like problem.py:148)."""
    n = ids.shape[0]
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return ids[lo:hi]


class FlatGradBucket(object):
    """All parameter gradients of a model as views into ONE contiguous fp32 buffer.

    `head` parameters (the small late layers whose gradients exist first) sit at the front so that
    `all_reduce_head()` can start while the big layer-1 weight gradients are still being computed."""

    def __init__(self, params, head=(), device=None):
        params = list(params)
        head_ids = set(id(p) for p in head)
        order = [p for p in params if id(p) in head_ids] + [p for p in params if id(p) not in head_ids]
        device = device or order[0].device
        # every slice starts on a 256-byte boundary (the kernels want 16-byte aligned rows, TMA wants more); the padding
        # stays zero: it adds nothing to the gradient norm and an Adam update of (param 0, grad 0) is 0
        pad = lambda n: (n + 63) // 64 * 64
        self.flat = torch.zeros(sum(pad(p.numel()) for p in order), dtype=torch.float32, device=device)
        self.views, self.offsets, off = {}, {}, 0
        for p in order:
            self.views[id(p)] = self.flat[off:off + p.numel()].view_as(p)
            self.offsets[id(p)] = off
            off += pad(p.numel())
            if id(p) in head_ids:
                self.head_numel = off
        if not head_ids:
            self.head_numel = 0
        self.params = order

    def grad_of(self, p):
        return self.views[id(p)]

    def attach(self):
        """Point every parameter's .grad at its slice (the optimiser then reads the reduced values in place)."""
        for p in self.params:
            p.grad = self.views[id(p)]

    def _reduce(self, t, scale, async_op=False):
        # weight FIRST, then sum: shards may differ by one seed (shard_seeds follows np.array_split), so every rank has its
        # own local/global factor and the result must be sum_r scale_r * g_r -- scaling the sum would let replicas drift
        if scale != 1.0:
            t.mul_(scale)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            work = dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op)
            if async_op:
                return work
        return None

    def all_reduce(self, scale=1.0):
        """`scale` this rank's gradients (local/global batch weighting), then sum over ranks: the result equals the
        single-process gradient of the mean loss over the GLOBAL batch (problem.py:33)."""
        self._reduce(self.flat, scale)

    def all_reduce_head(self, scale=1.0):
        self._reduce(self.flat[:self.head_numel], scale)

    def all_reduce_tail(self, scale=1.0):
        self._reduce(self.flat[self.head_numel:], scale)


class FusedAdam(object):
    """`clip_grad_norm(params, 5)` + `torch.optim.Adam.step()` (models.py:102-103) as ONE native call (gsage_adam_step: two
    launches) over flat buffers: the model's parameters are re-pointed at slices of one contiguous fp32 buffer, laid out
    like its FlatGradBucket, with flat first / second moment buffers beside it.  Same update as
    torch.optim.Adam(lr, betas, eps, weight_decay) after clip_grad_norm_(max_norm=clip).

    It quacks like a torch optimiser where the reference touches one: `param_groups` (a single group; `LRSchedule.set_lr`
    writes its 'lr', models.py:93-95), `zero_grad()`, `step()`.  The flat buffers are built at the first `step()` -- the
    reference constructs its optimiser inside `GSSupervised.__init__`, i.e. BEFORE `model.cuda()` (models.py:69,
    train.py:125-126), when the parameters are not on the device yet."""

    fused_clip = True

    def __init__(self, model, lr=0.01, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, lazy=False):
        self.model = model
        self.param_groups = [dict(params=list(model.parameters()), lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)]
        self.steps = 0
        self.flat = None
        if not lazy:
            self._materialize()

    # the hyper-parameters live in the param group, like torch's optimisers (LRSchedule.set_lr writes 'lr' there)
    lr = property(lambda self: self.param_groups[0]['lr'], lambda self, v: self.param_groups[0].__setitem__('lr', v))
    betas = property(lambda self: self.param_groups[0]['betas'])
    eps = property(lambda self: self.param_groups[0]['eps'])
    weight_decay = property(lambda self: self.param_groups[0]['weight_decay'])

    def _materialize(self):
        model = self.model
        self.bucket = model._bucket()
        order = self.bucket.params
        dev = self.bucket.flat.device
        self.flat = torch.zeros(self.bucket.flat.numel(), dtype=torch.float32, device=dev)
        for p in order:
            assert p.dtype == torch.float32 and p.is_cuda, 'FusedAdam: fp32 CUDA parameters (call model.cuda() before the first step)'
            n, off = p.numel(), self.bucket.offsets[id(p)]
            self.flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + n].view_as(p)            # the parameter now lives inside the flat buffer
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.scratch = torch.zeros(1, dtype=torch.float32, device=dev)
        model._reset_engines()                                    # engines hold raw pointers of the old parameter storage

    def step(self, clip=5.0):
        if self.flat is None:
            self._materialize()
        self.steps += 1
        check(lib().gsage_adam_step(ops.ptr(self.flat), ops.ptr(self.bucket.flat), ops.ptr(self.m), ops.ptr(self.v), self.flat.numel(),
                                    float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                                    self.steps, float(clip or 0.0), ops.ptr(self.scratch), ops.stream()))
        self.model._weights_epoch = getattr(self.model, '_weights_epoch', 0) + 1     # in-place update torch did not see

    def zero_grad(self, set_to_none=False):
        pass                                                      # every backward overwrites the whole bucket
