// gather_reduce.cu -- the fused gather + aggregate kernel (HBM-bound; the kernel the roofline is quoted on)
//                     plus the small per-row kernels around it (attention weights, L2 normalise).
//
// Replaces  feats[ids]                         (/root/reference/models.py:76,80)
//           neibs.view(N,S,d).mean(dim=1)      (/root/reference/nn_modules.py:197-198)
//           h.view(N,S,H).max(dim=1)[0] / mean (/root/reference/nn_modules.py:225-226,240,252)
//           softmax(bmm(...)) + weighted sum   (/root/reference/nn_modules.py:307-315)
//           F.normalize(dim=1)                 (/root/reference/models.py:90)
//
// Design (B200): a *row group* of LPR lanes (4..32, a power of two >= the row's 16-byte chunk count, capped at
// a warp) owns one parent.  The group fetches its S neighbour ids with one coalesced load, broadcasts them by
// shuffle, and streams the S rows with 16-byte `ld.global.nc.L1::no_allocate` loads -- every lane keeps
// U x CPL independent 16-byte loads in flight (U rows unrolled, CPL chunks per lane per row), fp32 accumulators
// in registers, no shared memory, no atomics, one vector store per chunk at the end.  The neighbour rows are
// never materialised in HBM: algorithmic traffic per parent = S*d*e (rows) + S*8 (ids) + d*e_out (result).
#include "common.cuh"
#include <float.h>

namespace gsage {

enum { kRedSum = 0, kRedMax = 1 };

// four CTAs per SM (<= 64 registers): the kernel lives on bytes in flight -- at 70 registers only three CTAs fit
#ifndef GS_GR_MIN_CTAS
#define GS_GR_MIN_CTAS 4
#endif
template <typename T, int LPR, int CPL, int RED>
__global__ void __launch_bounds__(256, (CPL <= 3 ? GS_GR_MIN_CTAS : 2))
gather_reduce_kernel(const T* __restrict__ table, int64_t ld, int64_t n_table_rows, int d,
                     const int64_t* __restrict__ ids, int64_t n_parents, int S, const float* __restrict__ weights,
                     float scale, void* __restrict__ out, int out_bf16, int64_t ld_out, int vec_store, int l2_hint) {
    constexpr int VEC = ElemTraits<T>::kPerVec;
    const uint64_t pol_in = l2_hint ? l2_policy_evict_first() : 0, pol_out = l2_hint ? l2_policy_evict_last() : 0;
    constexpr int U = (CPL == 1) ? 8 : (CPL == 2 ? 4 : 2);
    constexpr int GROUPS = 256 / LPR;
    const int lane_g = threadIdx.x & (LPR - 1);
    const int64_t parent = (int64_t)blockIdx.x * GROUPS + (threadIdx.x / LPR);
    if (parent >= n_parents) return;                       // whole groups leave together
    const unsigned gmask = (LPR == 32) ? 0xFFFFFFFFu : (((1u << LPR) - 1u) << ((threadIdx.x & 31) & ~(LPR - 1)));
    const int nchunks = (d + VEC - 1) / VEC;
    const int chunk0 = blockIdx.y * (LPR * CPL) + lane_g;

    float acc[CPL][VEC];
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[c][e] = (RED == kRedMax) ? -FLT_MAX : 0.0f;

    const int64_t first = parent * (int64_t)S;
    for (int j0 = 0; j0 < S; j0 += LPR) {
        const int cnt = min(LPR, S - j0);
        int64_t my_id = -1;
        float my_w = 1.0f;
        if (lane_g < cnt) {
            my_id = ids ? ids[first + j0 + lane_g] : (first + j0 + lane_g);
            if (weights) my_w = weights[first + j0 + lane_g];
        }
        for (int jj = 0; jj < cnt; jj += U) {
            uint4 v[U][CPL];
            float w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int src = min(jj + u, LPR - 1);
                const int64_t id = __shfl_sync(gmask, my_id, src, LPR);
                w[u] = __shfl_sync(gmask, my_w, src, LPR);
                const bool live = (jj + u < cnt) && ((uint64_t)id < (uint64_t)n_table_rows);
                if (!live) w[u] = 0.0f;
                const T* row = table + id * ld;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int ch = chunk0 + c * LPR;
                    v[u][c] = (live && ch < nchunks) ? (l2_hint ? ldg_nc_v4_hint(row + (int64_t)ch * VEC, pol_in) : ldg_nc_v4(row + (int64_t)ch * VEC)) : make_uint4(0, 0, 0, 0);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (RED == kRedMax && !(jj + u < cnt)) continue;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    float f[VEC];
                    ElemTraits<T>::unpack(v[u][c], f);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        if (RED == kRedMax) acc[c][e] = fmaxf(acc[c][e], f[e]);
                        else acc[c][e] = fmaf(w[u], f[e], acc[c][e]);
                    }
                }
            }
        }
    }

#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        const int ch = chunk0 + c * LPR;
        if (ch >= nchunks) continue;
        float f[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) f[e] = acc[c][e] * scale;
        const int col = ch * VEC;
        if (vec_store && col + VEC <= d) {
            if (out_bf16) {
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + parent * ld_out + col;
                if (VEC == 8) {
                    if (l2_hint) stg_v4_hint(o, ElemTraits<__nv_bfloat16>::pack(f), pol_out);
                    else *reinterpret_cast<uint4*>(o) = ElemTraits<__nv_bfloat16>::pack(f);
                } else {
                    *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
                }
            } else {
                float* o = reinterpret_cast<float*>(out) + parent * ld_out + col;
#pragma unroll
                for (int q = 0; q < VEC / 4; ++q)
                    *reinterpret_cast<float4*>(o + 4 * q) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if (col + e >= d) break;
                if (out_bf16) reinterpret_cast<__nv_bfloat16*>(out)[parent * ld_out + col + e] = __float2bfloat16_rn(f[e]);
                else reinterpret_cast<float*>(out)[parent * ld_out + col + e] = f[e];
            }
        }
    }
}

template <typename T, int LPR, int CPL>
static int launch_shape(const void* table, int64_t ld, int64_t rows, int d, const int64_t* ids, int64_t n_parents, int S,
                        int red, const float* weights, float scale, void* out, int out_bf16, int64_t ld_out,
                        int vec_store, int l2_hint, cudaStream_t s) {
    constexpr int VEC = ElemTraits<T>::kPerVec;
    const int nchunks = (d + VEC - 1) / VEC;
    dim3 grid((unsigned)ceil_div(n_parents, 256 / LPR), (unsigned)ceil_div(nchunks, LPR * CPL));
    if (red == kRedMax)
        gather_reduce_kernel<T, LPR, CPL, kRedMax><<<grid, 256, 0, s>>>((const T*)table, ld, rows, d, ids, n_parents, S,
                                                                        weights, scale, out, out_bf16, ld_out, vec_store, l2_hint);
    else
        gather_reduce_kernel<T, LPR, CPL, kRedSum><<<grid, 256, 0, s>>>((const T*)table, ld, rows, d, ids, n_parents, S,
                                                                        weights, scale, out, out_bf16, ld_out, vec_store, l2_hint);
    GS_LAUNCHED();
    return GSAGE_OK;
}

template <typename T>
static int launch_dtype(const void* table, int64_t ld, int64_t rows, int d, const int64_t* ids, int64_t n_parents, int S,
                        int red, const float* weights, float scale, void* out, int out_bf16, int64_t ld_out,
                        int vec_store, int l2_hint, cudaStream_t s) {
    constexpr int VEC = ElemTraits<T>::kPerVec;
    const int nchunks = (d + VEC - 1) / VEC;
#define GS_SHAPE(L, C) return launch_shape<T, L, C>(table, ld, rows, d, ids, n_parents, S, red, weights, scale, out, out_bf16, ld_out, vec_store, l2_hint, s)
    if (nchunks <= 4) GS_SHAPE(4, 1);
    if (nchunks <= 8) GS_SHAPE(8, 1);
    if (nchunks <= 16) GS_SHAPE(16, 1);
    if (nchunks <= 32) GS_SHAPE(32, 1);
    if (nchunks <= 64) GS_SHAPE(32, 2);
    if (nchunks <= 96) GS_SHAPE(32, 3);
    // wider rows: 128 chunks per pass, ceil(nchunks/128) column blocks in grid.y
    GS_SHAPE(32, 4);
#undef GS_SHAPE
}

// --------------------------------------------------------------------------------------------------
// attention weights: one warp per parent, lane j <-> neighbour j
// --------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) attention_weights_kernel(const T* __restrict__ na, const T* __restrict__ xa,
                                                                int64_t ld, int H, int64_t n_parents, int S,
                                                                float* __restrict__ w) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (p >= n_parents) return;
    const T* x = xa + p * ld;
    float mx = -FLT_MAX;
    for (int j = lane; j < S; j += 32) {
        const T* n = na + (p * S + j) * ld;
        float s = 0.0f;
        for (int h = 0; h < H; ++h) s = fmaf(ElemTraits<T>::load(n + h), ElemTraits<T>::load(x + h), s);
        w[p * S + j] = s;
        mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    float sum = 0.0f;
    for (int j = lane; j < S; j += 32) {
        const float e = expf(w[p * S + j] - mx);
        w[p * S + j] = e;
        sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    const float inv = 1.0f / sum;
    for (int j = lane; j < S; j += 32) w[p * S + j] *= inv;
}

// --------------------------------------------------------------------------------------------------
// L2 row normalise: one warp per row
// --------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) l2_normalize_kernel(const T* __restrict__ x, int64_t ld, int64_t n, int d,
                                                           float* __restrict__ out, int64_t ld_out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return;
    float ss = 0.0f;
    for (int c = lane; c < d; c += 32) {
        const float v = ElemTraits<T>::load(x + r * ld + c);
        ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    for (int c = lane; c < d; c += 32) out[r * ld_out + c] = ElemTraits<T>::load(x + r * ld + c) * inv;
}

// --------------------------------------------------------------------------------------------------
// normalise + classifier in one launch (models.py:90-91): zn = F.normalize(z, dim=1); logits = fc(zn).
// The head is 16 K rows x 256 -> 41 classes: two launches of nothing, but under the sample-ahead pipeline each of them
// queues behind the next batch's sampling kernels (0.06-0.07 ms of a 0.7-1.4 ms step, 0.024 alone).  One warp takes kHeadRows
// rows: a lane keeps VPL = D / 32 values of each, reads its VPL values of a classifier row once for all of them, and the
// class's dot products are reduced by shuffles.  fp32 FFMA throughout (exact mode and bf16 mode alike); zn is written for the
// backward pass.
// --------------------------------------------------------------------------------------------------
static constexpr int kHeadRows = 4;

template <int VPL>
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ z, int64_t ld, int64_t n, const float* __restrict__ w,
                                                   const float* __restrict__ b, int n_classes, float* __restrict__ zn, int64_t ld_zn,
                                                   float* __restrict__ logits, int64_t ld_logits) {
    constexpr int D = 32 * VPL;
    const int lane = threadIdx.x & 31;
    const int64_t r0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * kHeadRows;
    if (r0 >= n) return;
    float x[kHeadRows][VPL];
#pragma unroll
    for (int i = 0; i < kHeadRows; ++i) {
        const int64_t r = r0 + i < n ? r0 + i : n - 1;          // (a ragged last group repeats its last row; nothing of it is stored)
        float ss = 0.0f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) { x[i][v] = z[r * ld + lane + 32 * v]; ss = fmaf(x[i][v], x[i][v], ss); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
        const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            x[i][v] *= inv;
            if (r0 + i < n) zn[(r0 + i) * ld_zn + lane + 32 * v] = x[i][v];
        }
    }
    for (int c = 0; c < n_classes; ++c) {
        float wv[VPL], acc[kHeadRows];
#pragma unroll
        for (int v = 0; v < VPL; ++v) wv[v] = __ldg(w + (int64_t)c * D + lane + 32 * v);
#pragma unroll
        for (int i = 0; i < kHeadRows; ++i) {
            acc[i] = 0.0f;
#pragma unroll
            for (int v = 0; v < VPL; ++v) acc[i] = fmaf(x[i][v], wv[v], acc[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int i = 0; i < kHeadRows; ++i) acc[i] += __shfl_xor_sync(0xFFFFFFFFu, acc[i], o);
        }
        if (lane < kHeadRows && r0 + lane < n) {
            float mine = acc[0];
#pragma unroll
            for (int i = 1; i < kHeadRows; ++i) mine = lane == i ? acc[i] : mine;
            logits[(r0 + lane) * ld_logits + c] = mine + __ldg(b + c);
        }
    }
}

// few classes only: with 41 the per-class shuffle reductions make it slower than the two launches it replaces (reddit head 0.068 ->
// 0.098 ms, big10m 0.070 -> 0.101; pokec regression, 1 output: 0.079 -> 0.021)
bool head_fused_eligible(int d, int n_classes) {
    return d % 32 == 0 && d >= 32 && d <= 256 && n_classes <= 8 && getenv("GSAGE_NO_FUSED_HEAD") == nullptr;
}

int head_fused_launch(const float* z, int64_t ld, int64_t n, int d, const float* w, const float* b, int n_classes, float* zn, int64_t ld_zn,
                      float* logits, int64_t ld_logits, cudaStream_t s) {
    if (n == 0) return GSAGE_OK;
    const unsigned grid = (unsigned)ceil_div(ceil_div(n, (int64_t)kHeadRows), 8);
#define GS_HEAD(V) head_kernel<V><<<grid, 256, 0, s>>>(z, ld, n, w, b, n_classes, zn, ld_zn, logits, ld_logits)
    switch (d / 32) {
        case 1: GS_HEAD(1); break;
        case 2: GS_HEAD(2); break;
        case 3: GS_HEAD(3); break;
        case 4: GS_HEAD(4); break;
        case 5: GS_HEAD(5); break;
        case 6: GS_HEAD(6); break;
        case 7: GS_HEAD(7); break;
        default: GS_HEAD(8); break;
    }
#undef GS_HEAD
    GS_LAUNCHED();
    return GSAGE_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int gather_reduce_launch(const void* table, int dtype, int64_t ld, int64_t rows, int d, const int64_t* ids,
                         int64_t n_parents, int S, int reduce, const float* weights, void* out, int out_dtype,
                         int64_t ld_out, cudaStream_t s, int l2_hint) {
    const int vec = dtype == GSAGE_BF16 ? 8 : 4;
    GS_CHECK_ARG(table && out && d > 0 && S > 0 && n_parents >= 0, "gather_reduce: bad arguments");
    GS_CHECK_ARG(dtype == GSAGE_F32 || dtype == GSAGE_BF16, "gather_reduce: dtype must be f32 or bf16");
    GS_CHECK_ARG(out_dtype == GSAGE_F32 || out_dtype == GSAGE_BF16, "gather_reduce: out dtype must be f32 or bf16");
    GS_CHECK_ARG(aligned16(table) && (ld * dtype_size(dtype)) % 16 == 0 && ld >= ceil_div(d, vec) * vec,
                 "gather_reduce: table rows must be 16-byte aligned and padded to a whole 16-byte chunk "
                 "(ld=%lld, d=%d)", (long long)ld, d);
    GS_CHECK_ARG(ld_out >= d, "gather_reduce: ld_out < d");
    if (n_parents == 0) return GSAGE_OK;
    GS_CHECK_ARG(ceil_div(n_parents, 8) < (1LL << 31), "gather_reduce: too many parents for one launch");
    const int vec_store = aligned16(out) && (ld_out * dtype_size(out_dtype)) % 16 == 0;
    const float scale = (reduce == GSAGE_RED_MEAN) ? 1.0f / (float)S : 1.0f;
    const int red = (reduce == GSAGE_RED_MAX) ? kRedMax : kRedSum;
    if (dtype == GSAGE_BF16)
        return launch_dtype<__nv_bfloat16>(table, ld, rows, d, ids, n_parents, S, red, weights, scale, out,
                                           out_dtype == GSAGE_BF16, ld_out, vec_store, l2_hint, s);
    return launch_dtype<float>(table, ld, rows, d, ids, n_parents, S, red, weights, scale, out, out_dtype == GSAGE_BF16,
                               ld_out, vec_store, l2_hint, s);
}

}  // namespace gsage

using namespace gsage;

extern "C" {

int gsage_gather_reduce(const void* table_dev, int dtype, int64_t ld, int64_t n_table_rows, int d, const int64_t* ids_dev,
                        int64_t n_parents, int S, int reduce, const float* weights_dev, void* out_dev, int out_dtype,
                        int64_t ld_out, void* stream) {
    GS_CHECK_ARG(reduce == GSAGE_RED_MEAN || reduce == GSAGE_RED_MAX || reduce == GSAGE_RED_SUM, "gather_reduce: bad reduce op");
    return gather_reduce_launch(table_dev, dtype, ld, n_table_rows, d, ids_dev, n_parents, S, reduce, weights_dev, out_dev,
                                out_dtype, ld_out, as_stream(stream), 0);
}

int gsage_gather_rows(const void* table_dev, int dtype, int64_t ld, int64_t n_table_rows, int d, const int64_t* ids_dev,
                      int64_t n, void* out_dev, int out_dtype, int64_t ld_out, void* stream) {
    return gather_reduce_launch(table_dev, dtype, ld, n_table_rows, d, ids_dev, n, 1, GSAGE_RED_SUM, nullptr, out_dev,
                                out_dtype, ld_out, as_stream(stream), 0);
}

int gsage_attention_weights(const void* na_dev, const void* xa_dev, int dtype, int64_t ld, int H, int64_t n_parents, int S,
                            float* w_dev, void* stream) {
    GS_CHECK_ARG(na_dev && xa_dev && w_dev && H > 0 && ld >= H, "attention_weights: bad arguments");
    GS_CHECK_ARG(S > 1, "attention aggregator: S must be > 1 (the reference's squeeze() is ill-defined at S == 1)");
    if (n_parents == 0) return GSAGE_OK;
    const unsigned grid = (unsigned)ceil_div(n_parents, 8);
    if (dtype == GSAGE_BF16)
        attention_weights_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)na_dev, (const __nv_bfloat16*)xa_dev, ld, H, n_parents, S, w_dev);
    else
        attention_weights_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)na_dev, (const float*)xa_dev, ld, H, n_parents, S, w_dev);
    GS_LAUNCHED();
    return GSAGE_OK;
}

int gsage_l2_normalize(const void* x_dev, int dtype, int64_t ld, int64_t n, int d, float* out_dev, int64_t ld_out,
                       void* stream) {
    GS_CHECK_ARG(x_dev && out_dev && d > 0 && ld >= d && ld_out >= d, "l2_normalize: bad arguments");
    if (n == 0) return GSAGE_OK;
    const unsigned grid = (unsigned)ceil_div(n, 8);
    if (dtype == GSAGE_BF16)
        l2_normalize_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x_dev, ld, n, d, out_dev, ld_out);
    else
        l2_normalize_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)x_dev, ld, n, d, out_dev, ld_out);
    GS_LAUNCHED();
    return GSAGE_OK;
}

}  // extern "C"
