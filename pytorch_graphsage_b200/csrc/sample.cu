// sample.cu -- neighbour sampling kernels.
//
// Replaces  SparseUniformNeighborSampler.__call__  (/root/reference/nn_modules.py:80-101)
//           UniformNeighborSampler.__call__        (/root/reference/nn_modules.py:42-49)
//
// Sparse semantics reproduced exactly: output p = i*S + j takes the p-th bounded draw u in
// [0, n_cols), reduces it `u % degree(ids[i])` (numpy: x % 0 == 0) -- modulo-biased, with replacement -- and
// reads A[ids[i], c]; an absent entry (empty row / dummy node) reads 0.  One thread per sample: the S
// threads of a parent hit the same indptr pair (a warp broadcast), `sel` and `out` are coalesced, the only
// random access is the 4-byte neighbour value.
#include "graph.cuh"
#include "mt19937.cuh"

namespace gsage {

template <typename V>
__global__ void __launch_bounds__(256) sample_fast_kernel(const int64_t* __restrict__ indptr, const V* __restrict__ val,
                                                          int64_t n_rows, const int64_t* __restrict__ ids, int64_t total,
                                                          int S, const uint32_t* __restrict__ sel,
                                                          int64_t* __restrict__ out, int* __restrict__ err) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const int64_t id = ids[p / S];
    if (id < 0 || id >= n_rows) { *err = 1; out[p] = 0; return; }
    const int64_t lo = indptr[id], hi = indptr[id + 1];
    const uint32_t deg = (uint32_t)(hi - lo);
    int64_t r = 0;
    if (deg) r = (int64_t)val[lo + (sel[p] % deg)];
    out[p] = r;
}

template <typename V>
__global__ void __launch_bounds__(256) sample_general_kernel(const int64_t* __restrict__ indptr, const V* __restrict__ val,
                                                             const int32_t* __restrict__ col, const int32_t* __restrict__ deg_arr,
                                                             int64_t n_rows, const int64_t* __restrict__ ids, int64_t total,
                                                             int S, const uint32_t* __restrict__ sel,
                                                             int64_t* __restrict__ out, int* __restrict__ err) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const int64_t id = ids[p / S];
    if (id < 0 || id >= n_rows) { *err = 1; out[p] = 0; return; }
    const uint32_t deg = (uint32_t)deg_arr[id];                 // non-zero stored values, not row length
    const int32_t c = deg ? (int32_t)(sel[p] % deg) : 0;
    int64_t lo = indptr[id], hi = indptr[id + 1];
    while (lo < hi) {                                            // lower_bound on the sorted column ids
        const int64_t mid = (lo + hi) >> 1;
        if (col[mid] < c) lo = mid + 1; else hi = mid;
    }
    int64_t r = 0;
    if (lo < indptr[id + 1] && col[lo] == c) r = (int64_t)val[lo];
    out[p] = r;
}

__global__ void __launch_bounds__(256) sample_dense_kernel(const int64_t* __restrict__ adj, int64_t n_rows, int K,
                                                           const int64_t* __restrict__ ids, int64_t total, int S,
                                                           const int64_t* __restrict__ perm, int64_t* __restrict__ out,
                                                           int* __restrict__ err) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const int64_t id = ids[p / S];
    if (id < 0 || id >= n_rows) { if (err) *err = 1; out[p] = 0; return; }
    out[p] = adj[id * K + perm[p % S]];
}

int sample_sparse_launch(gsage_graph* g, const int64_t* ids, int64_t n, int S, const uint32_t* sel, int64_t* out,
                         cudaStream_t s) {
    const int64_t total = n * (int64_t)S;
    if (total == 0) return GSAGE_OK;
    const int64_t grid = ceil_div(total, 256);
    GS_CHECK_ARG(grid < (1LL << 31), "sample: batch too large");
    if (g->fast) {
        if (g->val64) sample_fast_kernel<int64_t><<<(unsigned)grid, 256, 0, s>>>(g->indptr, (const int64_t*)g->val, g->n_rows, ids, total, S, sel, out, g->err_flag);
        else          sample_fast_kernel<int32_t><<<(unsigned)grid, 256, 0, s>>>(g->indptr, (const int32_t*)g->val, g->n_rows, ids, total, S, sel, out, g->err_flag);
    } else {
        if (g->val64) sample_general_kernel<int64_t><<<(unsigned)grid, 256, 0, s>>>(g->indptr, (const int64_t*)g->val, g->col, g->deg, g->n_rows, ids, total, S, sel, out, g->err_flag);
        else          sample_general_kernel<int32_t><<<(unsigned)grid, 256, 0, s>>>(g->indptr, (const int32_t*)g->val, g->col, g->deg, g->n_rows, ids, total, S, sel, out, g->err_flag);
    }
    GS_LAUNCHED();
    return GSAGE_OK;
}

}  // namespace gsage

using namespace gsage;

extern "C" {

int gsage_sample_sparse(gsage_graph* g, const int64_t* ids_dev, int64_t n, int S, const uint32_t* sel_dev,
                        int64_t* out_dev, void* stream) {
    GS_CHECK_ARG(g, "sample_sparse: NULL graph");
    GS_CHECK_ARG(S > 0, "SparseUniformNeighborSampler: n_samples must be set explicitly (> 0)");
    GS_CHECK_ARG(n >= 0 && (n == 0 || (ids_dev && sel_dev && out_dev)), "sample_sparse: NULL buffer");
    return sample_sparse_launch(g, ids_dev, n, S, sel_dev, out_dev, as_stream(stream));
}

int gsage_sample_sparse_rng(gsage_graph* g, gsage_rng* r, const int64_t* ids_dev, int64_t n, int S, int64_t* out_dev,
                            void* stream) {
    GS_CHECK_ARG(g && r, "sample_sparse_rng: NULL graph / rng");
    GS_CHECK_ARG(S > 0, "SparseUniformNeighborSampler: n_samples must be set explicitly (> 0)");
    GS_CHECK_ARG(n >= 0 && (n == 0 || (ids_dev && out_dev)), "sample_sparse_rng: NULL buffer");
    GS_CHECK_ARG(g->n_cols >= 1 && g->n_cols <= 0xFFFFFFFFLL, "sample_sparse_rng: adjacency width out of range");
    const int64_t total = n * (int64_t)S;
    if (total == 0) return GSAGE_OK;
    // stream-ordered scratch for the n*S bounded draws
    uint32_t* sel = nullptr;
    GS_CUDA(cudaMallocAsync((void**)&sel, sizeof(uint32_t) * total, as_stream(stream)));
    int st = rng_randint_internal(r, (uint32_t)g->n_cols, total, sel, as_stream(stream));
    if (st == GSAGE_OK) st = sample_sparse_launch(g, ids_dev, n, S, sel, out_dev, as_stream(stream));
    cudaFreeAsync(sel, as_stream(stream));
    return st;
}

int gsage_sample_dense(const int64_t* adj_dev, int64_t n_rows, int K, const int64_t* ids_dev, int64_t n,
                       const int64_t* perm_dev, int S, int64_t* out_dev, void* stream) {
    GS_CHECK_ARG(adj_dev && perm_dev && K > 0 && n_rows > 0, "sample_dense: bad table");
    GS_CHECK_ARG(n >= 0 && (n == 0 || (ids_dev && out_dev)), "sample_dense: NULL buffer");
    if (S < 0) S = K + S < 0 ? 0 : K + S;                        // python slice semantics of tmp[:, :n_samples]
    if (S > K) S = K;
    const int64_t total = n * (int64_t)S;
    if (total == 0) return GSAGE_OK;
    sample_dense_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(adj_dev, n_rows, K, ids_dev, total, S,
                                                                                      perm_dev, out_dev, nullptr);
    GS_LAUNCHED();
    return GSAGE_OK;
}

}  // extern "C"
