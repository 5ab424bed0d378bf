"""
GraphCSR -- the device-resident adjacency both samplers read.

Host-side mirror of what the reference keeps as a scipy matrix on the CPU
(/root/reference/problem.py:70-72,85-87 and /root/reference/nn_modules.py:72-78).
"""

import ctypes as C

import numpy as np

from ._lib import check, lib
from . import ops


def _i64(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.int64)


class GraphCSR(object):
    def __init__(self, handle):
        self._h = handle
        n_rows, n_cols, nnz, dev_bytes = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        canon = C.c_int()
        check(lib().gsage_graph_info(self._h, C.byref(n_rows), C.byref(n_cols), C.byref(nnz), C.byref(canon), C.byref(dev_bytes)))
        self.shape = (n_rows.value, n_cols.value)
        self.nnz = nnz.value
        self.canonical = bool(canon.value)
        self.device_bytes = dev_bytes.value
        self._degrees = None

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, indptr, data, shape, indices=None):
        """scipy-canonical CSR arrays (sorted, duplicate-free).  `indices=None`: the reference's file
        convention, columns of every row are 0..deg-1."""
        ops._bind_device()
        indptr, data = _i64(indptr), _i64(data)
        indices = None if indices is None else _i64(indices)
        h = C.c_void_p()
        check(lib().gsage_graph_from_csr(indptr.ctypes.data, None if indices is None else indices.ctypes.data,
                                         data.ctypes.data, int(shape[0]), int(shape[1]), C.byref(h)))
        return cls(h)

    @classmethod
    def from_scipy(cls, adj):
        """What `SparseUniformNeighborSampler.__init__` is handed (any scipy.sparse matrix)."""
        from scipy import sparse
        assert sparse.issparse(adj), "SparseUniformNeighborSampler: not sparse.issparse(adj)"
        csr = adj.tocsr(copy=True)
        csr.sum_duplicates()
        csr.sort_indices()
        return cls.from_arrays(csr.indptr, csr.data, csr.shape, indices=csr.indices)

    @classmethod
    def from_triplets(cls, trip):
        """The 3 x nnz `[v; r; c]` array of a sparse problem file == parse_csr_matrix (problem.py:70-72)."""
        ops._bind_device()
        trip = np.asarray(trip)
        v, r, c = _i64(trip[0]), _i64(trip[1]), _i64(trip[2])
        h = C.c_void_p()
        check(lib().gsage_graph_from_triplets(v.ctypes.data, r.ctypes.data, c.ctypes.data, v.shape[0], C.byref(h)))
        return cls(h)

    @classmethod
    def from_synth(cls, adj):
        """pytorch_graphsage_b200.synth.make_sparse_adjacency output."""
        return cls.from_arrays(adj['indptr'], adj['data'], adj['shape'])

    # -- reference-visible attributes -------------------------------------------------------------
    @property
    def degrees(self):
        """`sampler.degrees` of the reference (non-zero entries per row)."""
        if self._degrees is None:
            out = np.empty(self.shape[0], dtype=np.int64)
            check(lib().gsage_graph_degrees_host(self._h, out.ctypes.data))
            self._degrees = out
        return self._degrees

    def check(self):
        """Raise IndexError if a sampler call saw an id outside the adjacency (scipy raises at once; we
        raise at the next sync point)."""
        check(lib().gsage_graph_check(self._h, ops.stream()))

    def __del__(self):
        try:
            if self._h:
                lib().gsage_graph_destroy(self._h)
                self._h = None
        except Exception:
            pass
