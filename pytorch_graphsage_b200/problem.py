"""
Batch iterator of the reference (`NodeProblem.iterate`, /root/reference/problem.py:141-153) with the per-epoch
shuffle drawn from the DEVICE MT19937 stream (`gsage_rng_permutation`), i.e. from the same global stream, at the
same position, the reference's `np.random.permutation` would use -- so an epoch of seed batches followed by the
sampler's draws stays bit-exact with the reference without a host round trip.  SURVEY.md 8(f) row 1.
"""

import numpy as np
import torch

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
