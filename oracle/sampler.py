"""
oracle/sampler.py -- TEST INFRASTRUCTURE ONLY (not a product path).

CPU (numpy) restatement of the reference's two neighbour samplers and of the CSR loader.

  * `csr_from_triplets`  follows /root/reference/problem.py:70-72  (`parse_csr_matrix`)
  * `row_degrees`        follows /root/reference/nn_modules.py:76-78 (degree table)
  * `sparse_sample`      follows /root/reference/nn_modules.py:80-101 (`SparseUniformNeighborSampler.__call__`)
  * `dense_sample`       follows /root/reference/nn_modules.py:42-49  (`UniformNeighborSampler.__call__`)

Pinned by tests/test_oracle_golden.py against fixtures produced by running the reference's own
classes (tests/golden/make_golden.py) and, when /root/reference is importable, against the live
reference.
"""

import numpy as np


def csr_from_triplets(v, r, c, shape=None):
    """`csr_matrix((v, (r, c)))` without scipy: duplicates summed, column indices sorted per row,
    shape inferred as (max r + 1, max c + 1) like scipy does.  Returns (indptr, indices, data, shape)."""
    v = np.asarray(v, dtype=np.int64)
    r = np.asarray(r, dtype=np.int64)
    c = np.asarray(c, dtype=np.int64)
    if shape is None:
        shape = (int(r.max()) + 1 if r.size else 0, int(c.max()) + 1 if c.size else 0)
    order = np.lexsort((c, r))
    r, c, v = r[order], c[order], v[order]
    if r.size:
        new = np.ones(r.size, dtype=bool)
        new[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
        grp = np.cumsum(new) - 1
        data = np.zeros(int(grp[-1]) + 1, dtype=np.int64)
        np.add.at(data, grp, v)
        r, c = r[new], c[new]
    else:
        data = v
    indptr = np.zeros(shape[0] + 1, dtype=np.int64)
    np.add.at(indptr, r + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr, c, data, shape


def row_degrees(indptr, data):
    """degrees[row] = number of *non-zero stored values* in the row (the reference counts
    `adj.nonzero()[0]`, which drops explicitly stored zeros)."""
    nz = (data != 0).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(nz)])
    return csum[indptr[1:]] - csum[indptr[:-1]]


def csr_lookup(indptr, indices, data, rows, cols):
    """A[rows, cols] point lookup; absent entries read as 0 (scipy `csr_sample_values`)."""
    out = np.zeros(rows.shape[0], dtype=np.int64)
    for p in range(rows.shape[0]):
        lo, hi = indptr[rows[p]], indptr[rows[p] + 1]
        k = lo + np.searchsorted(indices[lo:hi], cols[p])
        if k < hi and indices[k] == cols[p]:
            out[p] = data[k]
    return out


def csr_lookup_fast(indptr, indices, data, rows, cols):
    """Vectorised `csr_lookup` for the reference's on-disk convention (/root/reference/utils/convert.py:100-126:
    columns of row r are exactly 0..deg-1), falling back to the loop otherwise."""
    deg = indptr[rows + 1] - indptr[rows]
    k = indptr[rows] + cols
    inb = cols < deg
    ksafe = np.where(inb, k, 0)
    canonical = np.all(indices[ksafe][inb] == cols[inb]) if indices.size else True
    if not canonical:
        return csr_lookup(indptr, indices, data, rows, cols)
    out = np.zeros(rows.shape[0], dtype=np.int64)
    if indices.size:
        out[inb] = data[ksafe][inb]
    return out


def sparse_sample(indptr, indices, data, shape, degrees, ids, n_samples, draw):
    """out[i*S + j] = A[ids[i], u % degrees[ids[i]]] with u = the (i*S+j)-th bounded draw in [0, shape[1]).

    `draw(hi, count)` must return the legacy `np.random.choice(hi, count)` stream
    (numpy's RandomState.randint, or oracle.mt19937.MT19937Oracle.randint).
    numpy integer semantics: x % 0 == 0 (the reference only gets a RuntimeWarning), so a
    zero-degree row reads column 0, which is absent, which is 0 -- the dummy node."""
    assert n_samples > 0
    ids = np.asarray(ids, dtype=np.int64)
    if ids.size and (ids.min() < 0 or ids.max() >= shape[0]):
        raise IndexError("sparse_sample: id out of range of the adjacency (scipy raises too)")
    sel = np.asarray(draw(shape[1], ids.shape[0] * n_samples), dtype=np.int64).reshape(ids.shape[0], n_samples)
    deg = degrees[ids].reshape(-1, 1)
    with np.errstate(divide='ignore', invalid='ignore'):
        sel = np.where(deg == 0, 0, sel % np.where(deg == 0, 1, deg))
    rows = np.repeat(ids, n_samples)
    return csr_lookup_fast(indptr, indices, data, rows, sel.reshape(-1))


def dense_sample(adj, ids, n_samples, perm):
    """adj[ids][:, perm][:, :S] -- ONE column permutation shared by every row of the batch.
    `perm` is what `torch.randperm(K)` returned on the CPU generator."""
    tmp = adj[np.asarray(ids)]
    tmp = tmp[:, np.asarray(perm)]
    return tmp[:, :n_samples]
