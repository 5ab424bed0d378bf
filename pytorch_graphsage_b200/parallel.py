"""
Seed-node data parallelism: one process per GPU, graph + feature table replicated, every rank
samples / gathers / aggregates its own slice of the seed batch, and the ONLY collective is a sum all-reduce of the
parameter gradients, carried in one flat bucket so it is a single NCCL call (two when the head is overlapped with
the layer-1 weight gradients).  Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""

import ctypes as C
import os

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
    `all_reduce_head()` can start while the big layer-1 weight gradients are still being computed.

    Under NCCL with more than one rank the bucket is TWO buffers: `sym`, the rank's slice of a symmetric-memory allocation
    (mapped into every process of the box) that the backward kernels write and the PEERS read, and `flat`, the private buffer
    the reduced gradient lands in (`p.grad` are views of it; the optimiser reads it).  `all_reduce` is then one launch of
    gsage_peer_allreduce per rank: a one-shot all-reduce over NVLink peer memory fused with the gradient norm (peer_allreduce.cu),
    no NCCL call.  Buckets too large for a one-shot (a learned embedding table's dense gradient), other backends (gloo in the
    CPU tests) and `GSAGE_SYMM_ALLREDUCE=0` use torch.distributed's all_reduce on `flat` itself."""

    ONE_SHOT_MAX_BYTES = 8 << 20

    def __init__(self, params, head=(), device=None):
        params = list(params)
        head_ids = set(id(p) for p in head)
        order = [p for p in params if id(p) in head_ids] + [p for p in params if id(p) not in head_ids]
        device = device or order[0].device
        # every slice starts on a 256-byte boundary (the kernels want 16-byte aligned rows, TMA wants more); the padding
        # stays zero: it adds nothing to the gradient norm and an Adam update of (param 0, grad 0) is 0
        pad = lambda n: (n + 63) // 64 * 64
        total = sum(pad(p.numel()) for p in order)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.sym, self.peer_ptrs, self.epoch, self.collective = None, None, 0, 'none (1 rank)'
        self.sumsq = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.collective = 'torch.distributed all_reduce (%s)' % dist.get_backend()
            self._try_symmetric(total, device)
        src = self.sym if self.sym is not None else self.flat
        self.views, self.pre, self.offsets, off = {}, {}, {}, 0
        for p in order:
            self.views[id(p)] = self.flat[off:off + p.numel()].view_as(p)      # reduced gradient: what p.grad points at
            self.pre[id(p)] = src[off:off + p.numel()].view_as(p)             # where the backward kernels write
            self.offsets[id(p)] = off
            off += pad(p.numel())
            if id(p) in head_ids:
                self.head_numel = off
        if not head_ids:
            self.head_numel = 0
        self.params = order

    def _try_symmetric(self, total, device):
        if os.environ.get('GSAGE_SYMM_ALLREDUCE', '1') == '0' or torch.device(device).type != 'cuda' or dist.get_backend() != 'nccl':
            return
        if total * 4 > self.ONE_SHOT_MAX_BYTES or dist.get_world_size() > 16:
            self.collective += ' -- bucket of %.1f MB is too large for the one-shot peer all-reduce' % (total * 4 / 1e6)
            return
        try:
            import torch.distributed._symmetric_memory as symm_mem
            words = int(lib().gsage_peer_allreduce_words(total))
            sym = symm_mem.empty(words, dtype=torch.float32, device=torch.device(device))
            sym.zero_()
            handle = symm_mem.rendezvous(sym, dist.group.WORLD)
            ptrs = [int(x) for x in handle.buffer_ptrs]
            assert len(ptrs) == dist.get_world_size() and ptrs[dist.get_rank()] == sym.data_ptr()
            torch.cuda.synchronize()
            dist.barrier()                                               # every rank's flags are zero before anybody's first call
        except Exception as exc:                                         # no symmetric memory on this box: say so, use NCCL
            self.collective += ' -- symmetric memory unavailable (%s: %s)' % (type(exc).__name__, str(exc)[:80])
            return
        self.sym, self._handle = sym, handle
        self.peer_ptrs = (C.c_uint64 * len(ptrs))(*ptrs)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=device)
        self.collective = 'gsage_peer_allreduce: one-shot all-reduce over NVLink peer memory (symmetric memory), fused with the gradient norm'

    def grad_of(self, p):
        """Where the backward pass writes this parameter's LOCAL gradient (before the all-reduce)."""
        return self.pre[id(p)]

    def attach(self):
        """Point every parameter's .grad at its slice (the optimiser then reads the reduced values in place)."""
        for p in self.params:
            p.grad = self.views[id(p)]

    def _reduce(self, lo, hi, scale):
        # weight FIRST, then sum: shards may differ by one seed (shard_seeds follows np.array_split), so every rank has its
        # own local/global factor and the result must be sum_r scale_r * g_r -- scaling the sum would let replicas drift
        if self.sym is not None:
            assert lo == 0 and hi == self.flat.numel(), 'the one-shot peer all-reduce takes the whole bucket'
            self.epoch += 1
            check(lib().gsage_peer_allreduce(self.peer_ptrs, dist.get_world_size(), dist.get_rank(), self.flat.numel(), self.epoch,
                                             float(scale), ops.ptr(self.flat), ops.ptr(self.sumsq), ops.stream()))
            return
        t = self.flat[lo:hi]
        if scale != 1.0:
            t.mul_(scale)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    @property
    def one_shot(self):
        return self.sym is not None

    def all_reduce(self, scale=1.0):
        """`scale` this rank's gradients (local/global batch weighting), then sum over ranks: the result equals the
        single-process gradient of the mean loss over the GLOBAL batch (problem.py:33)."""
        self._reduce(0, self.flat.numel(), scale)

    def all_reduce_head(self, scale=1.0):
        self._reduce(0, self.head_numel, scale)

    def all_reduce_tail(self, scale=1.0):
        self._reduce(self.head_numel, self.flat.numel(), scale)


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
