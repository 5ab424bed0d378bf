// backward.cuh -- launch wrappers of the gradient kernels (backward.cu)
#pragma once
#include "common.cuh"
#include <algorithm>

namespace gsage {
// dW[o, k] = sum_r G[r, o] * A[ids ? ids[r] : r, k]   (dW zeroed first; fp32)
int wgrad_launch(const float* G, int64_t ldg, int O, const void* A, int a_dtype, int64_t lda, const int64_t* ids, int d,
                 int64_t n, float* dW, int64_t lddw, cudaStream_t s, bool accumulate = false);
// dense gradient of a learned embedding table (nn.Embedding's backward): table_grad[ids[i], :] += scale * rows[i / S, :]
// for i < n_ids (S consecutive ids share one source row: the mean aggregator's broadcast; S = 1 for self rows)
int embedding_scatter_launch(const float* rows, int64_t ld, int d, const int64_t* ids, int64_t n_ids, int S, float scale,
                             float* table_grad, int64_t ld_table, int64_t table_rows, cudaStream_t s);
// tensor-core version (wgrad_umma.cu): bf16 G and A, O == 128; up to four jobs per launch (all with the same n and d)
struct WgradJob {
    const void* G; int g_dtype; int64_t ldg; int O;       // G: (n, O) slice of the output gradient
    const void* A; int a_dtype; int64_t lda; const int64_t* ids; int d;
    int64_t n; float* dW; int64_t lddw;
    int64_t a_rows = 0;                                   // rows of the table behind A when known (ids outside it read as zero rows)
};
bool wgrad_umma_eligible(const WgradJob& j);
int wgrad_umma_launch(const WgradJob* jobs, int n_jobs, cudaStream_t s, bool accumulate = false);   // accumulate: dW += instead of dW =
int l2_normalize_bwd_launch(const float* z, const float* dzn, int64_t n, int d, int act, float* dz, cudaStream_t s, void* dz_bf16 = nullptr);
int layer1_grad_launch(const float* dh0, const float* dm2, const void* H, int h_dtype, int64_t ldh, int64_t n0, int64_t n1, int S,
                       int width, int act, void* dH, int dh_dtype, cudaStream_t s);
int colsum_launch(const float* x, int64_t n, int d, float* out, cudaStream_t s);
// attention aggregator, backwards (see backward.cu)
int attention_dw_launch(const void* table, int64_t ld, int64_t n_table_rows, int d, const int64_t* ids, int64_t n_parents, int S,
                        const float* dM, int64_t ld_dm, const float* w, float* dw, float* dN, int64_t ld_dn, cudaStream_t s);
int attention_softmax_bwd_launch(const float* w, const float* dw, const float* na, const float* xa, int H, int64_t n_parents, int S,
                                 float* dA, float* dXA, cudaStream_t s);
int tanh_bwd_launch(float* dt1, const float* t1, int64_t n, cudaStream_t s, int H = 0, void* padded_bf16 = nullptr);
int sum_act_grad_launch(const float* a0, const float* b0, const float* a1, const float* b1, const void* H, int h_dtype, int64_t ldh,
                        int64_t n0, int64_t n1, int width, int act, void* dH, int dh_dtype, cudaStream_t s);
// column sums of a bf16 (n, ld) matrix over its first d columns -> fp32
int colsum_bf16_launch(const void* x, int64_t ld, int64_t n, int d, float* out, cudaStream_t s);
}
