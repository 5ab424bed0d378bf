// linear.cuh -- parameter block shared by the projection kernels (linear_simt.cu, linear_umma.cu)
#pragma once
#include "common.cuh"

namespace gsage {

struct LinearSeg {
    const void* a; int a_dtype; int64_t lda; const int64_t* ids;
    const void* w; int w_dtype; int64_t ldw; int d; int O;
    const float* bias; int64_t col0;
    int S = 1;      // > 1: A row r = mean_j A[ids[r*S + j]] (fused gather+mean)
    int w_trans = 0;   // 1: W is stored (d x O): element (o, k) at w[k * ldw + o]  (backward: dX = dY . W; FFMA kernel only)
    const void* w_hi = nullptr; const void* w_lo = nullptr;   // fp32 operands only: tf32-exact split of W (w = w_hi + w_lo up to 2^-21): lets the
                          // weight-stationary kernel run the projection as 3 x TF32 (fp32-level accuracy on the tensor cores)
    int O_store = 0;      // > 0: only the first O_store of the O computed columns are stored (W rows O_store.. are padding:
                          // the classifier's 41 classes ride in a 48-row operand on the tensor-core kernel)
    int64_t a_rows = 0;   // rows of the table behind `a` when known (> 0): ids outside it then read as ZERO rows in the TMA
                          // gathers (like gather_reduce.cu) instead of touching memory past the table
};

struct LinearParams {
    LinearSeg seg[2];
    int n_segs; int64_t n; int act;
    void* out; int out_dtype; int64_t ld_out;
    int pool_S = 1;    // > 1 (tensor-core kernel only): reduce every S consecutive output rows, store one row per parent
    int pool_max = 1;  // 1 = max, 0 = mean
};

int linear_simt_launch(const LinearParams& P, cudaStream_t s);
// tcgen05 kernel (linear_umma.cu): bf16 operands, 16-byte aligned rows, O % 16 == 0, <= 256 accumulator columns
bool linear_umma_eligible(const LinearParams& P);
int linear_umma_launch(const LinearParams& P, cudaStream_t s);
// weight-stationary tcgen05 kernel (linear_ws_umma.cu): same operands, W of a phase resident in shared memory (preferred when it fits)
bool linear_ws_umma_eligible(const LinearParams& P);
int linear_ws_umma_launch(const LinearParams& P, cudaStream_t s);
// 3 x TF32 variant (fp32 operands with w_hi / w_lo given, weights fully resident): same kernel, fp32-level accuracy
bool linear_ws_umma_x3_eligible(const LinearParams& P);
int linear_ws_umma_x3_launch(const LinearParams& P, cudaStream_t s);
// gather + mean + neighbour projection in one kernel (gather_mean_project_umma.cu)
bool gather_mean_project_eligible(const void* table, int dtype, int64_t ld, int d, int S, const void* w, int w_dtype, int64_t ldw, int O);
int gather_mean_project_launch(const void* table, int64_t ld, int64_t table_rows, int d, const int64_t* ids, int64_t n, int S,
                               const void* w, int64_t ldw, int O, const float* bias, int act, void* out, int out_dtype, int64_t ld_out,
                               int64_t col0, cudaStream_t s);
// w -> (w_hi, w_lo): w_hi = w with the 13 low mantissa bits cleared (exact in tf32), w_lo = the same of (w - w_hi)
int split_tf32_launch(const float* w, int64_t n, float* w_hi, float* w_lo, cudaStream_t s);
// pool aggregators: MLP + pool over the S rows of a parent in one tcgen05 kernel (linear_pool_umma.cu, swap-AB)
bool linear_pool_umma_eligible(const LinearParams& P);
int linear_pool_umma_launch(const LinearParams& P, cudaStream_t s);
// the attention aggregator's reduction in one kernel (attention_umma.cu): scores on tcgen05, softmax + weighted sum on chip
bool attention_fused_eligible(const void* a, int a_dtype, int64_t lda, int d, const void* w1, int w1_dtype, int64_t ldw, int H, int S,
                              int64_t n_parents, const void* out, int64_t ld_out, int out_dtype);
int attention_fused_launch(const void* a, int64_t lda, const int64_t* ids, int d, const void* w1, int64_t ldw, const float* b1,
                           const float* w2, const float* xa, int64_t n_parents, int S, void* out, int out_dtype, int64_t ld_out,
                           cudaStream_t s, int64_t table_rows = 0);
// weight-stationary, double-buffered version of the pool kernel (linear_pool_ws_umma.cu): relu, S <= 64
bool linear_pool_ws_umma_eligible(const LinearParams& P);
int linear_pool_ws_umma_launch(const LinearParams& P, cudaStream_t s);
int linear_pool_ws_umma_backward_launch(const LinearParams& P, const float* dP, int64_t ld_dp, void* dhid, int64_t ld_dhid, float* db,
                                        cudaStream_t s);
// picks the tensor-core kernel when every operand qualifies and `exact` == 0, else the FFMA kernel
int linear_dispatch(const LinearParams& P, int exact, cudaStream_t s);

}  // namespace gsage
