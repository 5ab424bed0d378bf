// backward.cu -- gradient kernels of the hot path (fp32 accumulate; the exact, FFMA generation).
//
// The reference gets these from autograd through `loss.backward()` (/root/reference/models.py:101); the
// gradients are the payload of the one collective of the seed-sharded multi-GPU run.
//   wgrad_simt_kernel        dW[o, k] += sum_r G[r, o] * A[row(r), k]      (fc_x / fc_neib / fc weight gradients;
//                            the self rows are gathered by id straight from the feature table, like the forward)
//   l2_normalize_bwd_kernel  gradient through F.normalize(dim=1)            (models.py:90)
//   layer1_grad_kernel       dH = [dh0 ; broadcast(dm2)/S] * act'(H)       (mean over S + concat + relu, backwards)
//   colsum_kernel            bias gradient
// Tensor-core (tcgen05, MN-major operands) versions of wgrad are the next step; these are correct first.
#include "backward.cuh"

namespace gsage {

__device__ __forceinline__ float ld_any(const void* base, int dtype, int64_t idx) {
    return dtype == GSAGE_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx])
                               : reinterpret_cast<const float*>(base)[idx];
}

// 64 (o) x 64 (k) tile of dW per CTA, a chunk of rows per CTA (grid.z), fp32 atomics into dW
static constexpr int WT = 64, WR = 16;

__global__ void __launch_bounds__(256) wgrad_simt_kernel(const float* __restrict__ G, int64_t ldg, int O,
                                                         const void* __restrict__ A, int a_dtype, int64_t lda,
                                                         const int64_t* __restrict__ ids, int d, int64_t n, int64_t rows_per_cta,
                                                         float* __restrict__ dW, int64_t lddw) {
    __shared__ float Gs[WR][WT + 4];
    __shared__ float As[WR][WT + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int o0 = blockIdx.y * WT, k0 = blockIdx.x * WT;
    const int64_t r_begin = (int64_t)blockIdx.z * rows_per_cta;
    const int64_t r_end = min(n, r_begin + rows_per_cta);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    // loader mapping: 16 rows x 64 cols = 1024 elements, 4 per thread
    const int lr = tid >> 4, lc = (tid & 15) * 4;
    for (int64_t r0 = r_begin; r0 < r_end; r0 += WR) {
        const int64_t r = r0 + lr;
        int64_t src = -1;
        if (r < r_end) src = ids ? ids[r] : r;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float g = 0.0f, a = 0.0f;
            if (r < r_end) {
                if (o0 + lc + q < O) g = G[r * ldg + o0 + lc + q];
                if (k0 + lc + q < d) a = ld_any(A, a_dtype, src * lda + k0 + lc + q);
            }
            Gs[lr][lc + q] = g;
            As[lr][lc + q] = a;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < WR; ++rr) {
            const float4 g4 = *reinterpret_cast<const float4*>(&Gs[rr][ty * 4]);
            const float4 a4 = *reinterpret_cast<const float4*>(&As[rr][tx * 4]);
            const float g[4] = {g4.x, g4.y, g4.z, g4.w}, a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(g[i], a[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int o = o0 + ty * 4 + i;
        if (o >= O) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < d) atomicAdd(dW + (int64_t)o * lddw + k, acc[i][j]);
        }
    }
}

__global__ void __launch_bounds__(256) l2_normalize_bwd_kernel(const float* __restrict__ z, const float* __restrict__ dzn,
                                                               int64_t n, int d, int act, float* __restrict__ dz,
                                                               __nv_bfloat16* __restrict__ dz_bf16) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return;
    float ss = 0.0f, dot = 0.0f;
    for (int c = lane; c < d; c += 32) {
        const float v = z[r * d + c];
        ss = fmaf(v, v, ss);
        dot = fmaf(v, dzn[r * d + c], dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
        dot += __shfl_xor_sync(0xFFFFFFFFu, dot, o);
    }
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    const float proj = dot * inv * inv;                       // <zn, dzn> / ||z||
    for (int c = lane; c < d; c += 32) {
        const float v = z[r * d + c];
        float g = (dzn[r * d + c] - v * proj) * inv;
        if (act == GSAGE_ACT_RELU) g = v > 0.0f ? g : 0.0f;   // z is the post-activation output of layer 2
        else if (act == GSAGE_ACT_TANH) g *= (1.0f - v * v);
        dz[r * d + c] = g;
        if (dz_bf16) dz_bf16[r * d + c] = __float2bfloat16_rn(g);     // operand copy for the tensor-core gradient kernels
    }
}

// dH[r, c]: r < n0 -> dh0[r, c];  r >= n0 -> dm2[(r - n0) / S, c] / S;  times act'(H[r, c]).
// One warp per row, 8 columns per lane per pass (16-byte loads of H, 16- or 32-byte stores): the first version (one
// element per thread, a 64-bit division each) ran at 0.8 TB/s -- 0.56 ms of the 3.1 ms train step.
__global__ void __launch_bounds__(256) layer1_grad_kernel(const float* __restrict__ dh0, const float* __restrict__ dm2,
                                                          const void* __restrict__ H, int h_dtype, int64_t ldh, int64_t n0,
                                                          int64_t n1, int S, int width, int act, void* __restrict__ dH, int dh_bf16) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n0 + n1) return;
    const bool seed = r < n0;
    const float* g_row = seed ? dh0 + r * width : dm2 + ((r - n0) / S) * width;
    const float scale = seed ? 1.0f : 1.0f / (float)S;
    for (int c = lane * 8; c < width; c += 256) {
        float g[8], h[8];
        if ((width & 7) == 0 && (ldh & 7) == 0) {
            const float4 a = *reinterpret_cast<const float4*>(g_row + c), b = *reinterpret_cast<const float4*>(g_row + c + 4);
            g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
            if (h_dtype == GSAGE_BF16) {
                const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(H) + r * ldh + c);
                ElemTraits<__nv_bfloat16>::unpack(v, h);
            } else {
                const float* hp = reinterpret_cast<const float*>(H) + r * ldh + c;
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = hp[e];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float v = g[e] * scale;
                if (act == GSAGE_ACT_RELU) v = h[e] > 0.0f ? v : 0.0f;
                else if (act == GSAGE_ACT_TANH) v *= (1.0f - h[e] * h[e]);
                g[e] = v;
            }
            if (dh_bf16) {
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dH) + r * width + c) = ElemTraits<__nv_bfloat16>::pack(g);
            } else {
                float* o = reinterpret_cast<float*>(dH) + r * width + c;
                *reinterpret_cast<float4*>(o) = make_float4(g[0], g[1], g[2], g[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(g[4], g[5], g[6], g[7]);
            }
        } else {
            for (int e = 0; e < 8 && c + e < width; ++e) {        // rows that are not 16-byte aligned: scalar
                float v = g_row[c + e] * scale;
                const float hv = ld_any(H, h_dtype, r * ldh + c + e);
                if (act == GSAGE_ACT_RELU) v = hv > 0.0f ? v : 0.0f;
                else if (act == GSAGE_ACT_TANH) v *= (1.0f - hv * hv);
                if (dh_bf16) reinterpret_cast<__nv_bfloat16*>(dH)[r * width + c + e] = __float2bfloat16_rn(v);
                else reinterpret_cast<float*>(dH)[r * width + c + e] = v;
            }
        }
    }
}

__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int64_t n, int d, float* __restrict__ out) {
    const int c = blockIdx.x;
    float s = 0.0f;
    for (int64_t r = threadIdx.x; r < n; r += blockDim.x) s += x[r * d + c];
    __shared__ float sm[256];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = sm[0];
}

// one warp per id: lanes stride over the d columns (coalesced RED requests, fire and forget)
__global__ void __launch_bounds__(256) embedding_scatter_kernel(const float* __restrict__ rows, int64_t ld, int d,
                                                                const int64_t* __restrict__ ids, int64_t n_ids, int S, float scale,
                                                                float* __restrict__ table_grad, int64_t ld_table, int64_t table_rows) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n_ids) return;
    const int64_t id = ids[i];
    if ((uint64_t)id >= (uint64_t)table_rows) return;           // ids outside the table read as zero rows in the forward
    const float* src = rows + (i / S) * ld;
    float* dst = table_grad + id * ld_table;
    for (int c = lane; c < d; c += 32) atomicAdd(dst + c, src[c] * scale);
}

// the same with 16-byte vector reductions (red.global.add.v4.f32, sm_90+): a quarter of the L2 atomic requests.  A thread owns four
// columns of one id; LPI = d / 4 threads per id (d = 64: two ids per warp).  Needs d % 4 == 0 and 16-byte aligned rows.
__global__ void __launch_bounds__(256) embedding_scatter_v4_kernel(const float* __restrict__ rows, int64_t ld, int lpi,
                                                                   const int64_t* __restrict__ ids, int64_t n_ids, int S, float scale,
                                                                   float* __restrict__ table_grad, int64_t ld_table, int64_t table_rows) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / lpi;
    if (i >= n_ids) return;
    const int c = (int)(t - i * lpi) * 4;
    const int64_t id = __ldg(ids + i);
    if ((uint64_t)id >= (uint64_t)table_rows) return;
    const float4 v = *reinterpret_cast<const float4*>(rows + (i / S) * ld + c);
    float* dst = table_grad + id * ld_table + c;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x * scale), "f"(v.y * scale), "f"(v.z * scale), "f"(v.w * scale) : "memory");
}

int embedding_scatter_launch(const float* rows, int64_t ld, int d, const int64_t* ids, int64_t n_ids, int S, float scale,
                             float* table_grad, int64_t ld_table, int64_t table_rows, cudaStream_t s) {
    if (n_ids == 0) return GSAGE_OK;
    const bool vec = d % 4 == 0 && ld % 4 == 0 && ld_table % 4 == 0 && ((uintptr_t)rows & 15) == 0 && ((uintptr_t)table_grad & 15) == 0;
    if (vec) {
        const int lpi = d / 4;
        embedding_scatter_v4_kernel<<<(unsigned)ceil_div(n_ids * lpi, 256), 256, 0, s>>>(rows, ld, lpi, ids, n_ids, S, scale, table_grad, ld_table, table_rows);
    } else {
        embedding_scatter_kernel<<<(unsigned)ceil_div(n_ids, 8), 256, 0, s>>>(rows, ld, d, ids, n_ids, S, scale, table_grad, ld_table, table_rows);
    }
    GS_LAUNCHED();
    return GSAGE_OK;
}

// ---- attention aggregator, backwards (nn_modules.py:307-315) ------------------------------------------------------------
//   m_p = sum_j w_pj n_pj,  w_p = softmax_j(s_pj),  s_pj = <a(n_pj), a(x_p)>,  a(v) = W2 tanh(W1 v)

// dw[p*S + j] = <dM[p], n_pj>  (gradient of the softmax weights), and optionally dN[p*S + j, :] = w_pj * dM[p, :] (the part of
// the neighbour rows' gradient that does not go through the attention MLP; layer 2 only).  One warp per parent: every lane
// keeps its 16-byte chunks of dM[p] in registers and streams the S neighbour rows once.
__global__ void __launch_bounds__(256) attention_dw_kernel(const __nv_bfloat16* __restrict__ table, int64_t ld, int64_t n_table_rows, int d,
                                                           const int64_t* __restrict__ ids, int64_t n_parents, int S,
                                                           const float* __restrict__ dM, int64_t ld_dm, const float* __restrict__ w,
                                                           float* __restrict__ dw, float* __restrict__ dN, int64_t ld_dn) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (p >= n_parents) return;
    const int nch = (d + 7) / 8;
    for (int j = 0; j < S; ++j) {
        const int64_t r = p * S + j;
        const int64_t id = ids ? ids[r] : r;
        const bool live = (uint64_t)id < (uint64_t)n_table_rows;
        const float wj = dN ? w[r] : 0.0f;
        float acc = 0.0f;
        for (int c = lane; c < nch; c += 32) {
            float g[8], f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) g[e] = (c * 8 + e < d) ? dM[p * ld_dm + c * 8 + e] : 0.0f;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (live) v = ldg_nc_v4(table + id * ld + (int64_t)c * 8);
            ElemTraits<__nv_bfloat16>::unpack(v, f);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc = fmaf(g[e], f[e], acc);
            if (dN) {
#pragma unroll
                for (int e = 0; e < 8; ++e) if (c * 8 + e < d) dN[r * ld_dn + c * 8 + e] = wj * g[e];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
        if (lane == 0) dw[r] = acc;
    }
}

// softmax backwards + the score's two factors: ds = w (dw - <w, dw>);  dA[r, :] = ds_r * xa[p, :];  dXA[p, :] = sum_j ds_r * na[r, :]
__global__ void __launch_bounds__(256) attention_softmax_bwd_kernel(const float* __restrict__ w, const float* __restrict__ dw,
                                                                    const float* __restrict__ na, const float* __restrict__ xa, int H,
                                                                    int64_t n_parents, int S, float* __restrict__ dA, float* __restrict__ dXA) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (p >= n_parents) return;
    float dot = 0.0f;
    for (int j = lane; j < S; j += 32) dot = fmaf(w[p * S + j], dw[p * S + j], dot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xFFFFFFFFu, dot, o);
    const float xav = lane < H ? xa[p * H + lane] : 0.0f;          // H <= 32: lane h holds xa[p, h]
    float dxa = 0.0f;
    for (int j = 0; j < S; ++j) {
        const int64_t r = p * S + j;
        const float ds = w[r] * (dw[r] - dot);
        if (lane < H) {
            dA[r * H + lane] = ds * xav;
            dxa = fmaf(ds, na[r * H + lane], dxa);
        }
    }
    if (lane < H) dXA[p * H + lane] = dxa;
}

// dpre = dt1 * (1 - t1^2)   (tanh', in place on dt1)
// `padded` (optional): bf16 copy with rows of 128 elements (columns H.. stay zero) = the operand of the tensor-core weight gradient
__global__ void __launch_bounds__(256) tanh_bwd_kernel(float* __restrict__ dt1, const float* __restrict__ t1, int64_t n, int H,
                                                       __nv_bfloat16* __restrict__ padded) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float t = t1[i];
        const float v = dt1[i] * (1.0f - t * t);
        dt1[i] = v;
        if (padded) padded[(i / H) * 128 + (i % H)] = __float2bfloat16_rn(v);
    }
}

// dH[r, :] = (a[r, :] + b[r, :]) * act'(H[r, :]) for the two row ranges [0, n0) (self rows) and [n0, n0 + n1) (neighbour rows)
__global__ void __launch_bounds__(256) sum_act_grad_kernel(const float* __restrict__ a0, const float* __restrict__ b0, const float* __restrict__ a1,
                                                           const float* __restrict__ b1, const void* __restrict__ H, int h_dtype, int64_t ldh,
                                                           int64_t n0, int64_t n1, int width, int act, void* __restrict__ dH, int dh_bf16) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n0 + n1) * width) return;
    const int64_t r = i / width;
    const int c = (int)(i - r * width);
    float g = r < n0 ? a0[r * width + c] + b0[r * width + c] : a1[(r - n0) * width + c] + b1[(r - n0) * width + c];
    const float h = ld_any(H, h_dtype, r * ldh + c);
    if (act == GSAGE_ACT_RELU) g = h > 0.0f ? g : 0.0f;
    else if (act == GSAGE_ACT_TANH) g *= (1.0f - h * h);
    if (dh_bf16) reinterpret_cast<__nv_bfloat16*>(dH)[i] = __float2bfloat16_rn(g);
    else reinterpret_cast<float*>(dH)[i] = g;
}

int attention_dw_launch(const void* table, int64_t ld, int64_t n_table_rows, int d, const int64_t* ids, int64_t n_parents, int S,
                        const float* dM, int64_t ld_dm, const float* w, float* dw, float* dN, int64_t ld_dn, cudaStream_t s) {
    if (n_parents == 0) return GSAGE_OK;
    attention_dw_kernel<<<(unsigned)ceil_div(n_parents, 8), 256, 0, s>>>((const __nv_bfloat16*)table, ld, n_table_rows, d, ids, n_parents, S, dM,
                                                                        ld_dm, w, dw, dN, ld_dn);
    GS_LAUNCHED();
    return GSAGE_OK;
}

int attention_softmax_bwd_launch(const float* w, const float* dw, const float* na, const float* xa, int H, int64_t n_parents, int S,
                                 float* dA, float* dXA, cudaStream_t s) {
    GS_CHECK_ARG(H <= 32, "attention_softmax_bwd: attention width must be <= 32");
    if (n_parents == 0) return GSAGE_OK;
    attention_softmax_bwd_kernel<<<(unsigned)ceil_div(n_parents, 8), 256, 0, s>>>(w, dw, na, xa, H, n_parents, S, dA, dXA);
    GS_LAUNCHED();
    return GSAGE_OK;
}

int tanh_bwd_launch(float* dt1, const float* t1, int64_t n, cudaStream_t s, int H, void* padded_bf16) {
    if (n == 0) return GSAGE_OK;
    tanh_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(dt1, t1, n, H > 0 ? H : 1, (__nv_bfloat16*)padded_bf16);
    GS_LAUNCHED();
    return GSAGE_OK;
}

int sum_act_grad_launch(const float* a0, const float* b0, const float* a1, const float* b1, const void* H, int h_dtype, int64_t ldh,
                        int64_t n0, int64_t n1, int width, int act, void* dH, int dh_dtype, cudaStream_t s) {
    const int64_t total = (n0 + n1) * width;
    if (total == 0) return GSAGE_OK;
    sum_act_grad_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(a0, b0, a1, b1, H, h_dtype, ldh, n0, n1, width, act, dH,
                                                                      dh_dtype == GSAGE_BF16 ? 1 : 0);
    GS_LAUNCHED();
    return GSAGE_OK;
}

int wgrad_launch(const float* G, int64_t ldg, int O, const void* A, int a_dtype, int64_t lda, const int64_t* ids, int d,
                 int64_t n, float* dW, int64_t lddw, cudaStream_t s, bool accumulate) {
    if (!accumulate) GS_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)O * lddw, s));
    if (n == 0) return GSAGE_OK;
    // enough row chunks to fill the machine a few times over, at least 256 rows each
    const int64_t tiles = ceil_div(O, WT) * ceil_div(d, WT);
    int64_t chunks = std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), (4 * 148 + tiles - 1) / tiles));
    const int64_t rows_per_cta = ceil_div(ceil_div(n, chunks), WR) * WR;
    chunks = ceil_div(n, rows_per_cta);
    dim3 grid((unsigned)ceil_div(d, WT), (unsigned)ceil_div(O, WT), (unsigned)chunks);
    wgrad_simt_kernel<<<grid, 256, 0, s>>>(G, ldg, O, A, a_dtype, lda, ids, d, n, rows_per_cta, dW, lddw);
    GS_LAUNCHED();
    return GSAGE_OK;
}

int l2_normalize_bwd_launch(const float* z, const float* dzn, int64_t n, int d, int act, float* dz, cudaStream_t s, void* dz_bf16) {
    if (n == 0) return GSAGE_OK;
    l2_normalize_bwd_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, s>>>(z, dzn, n, d, act, dz, (__nv_bfloat16*)dz_bf16);
    GS_LAUNCHED();
    return GSAGE_OK;
}

int layer1_grad_launch(const float* dh0, const float* dm2, const void* H, int h_dtype, int64_t ldh, int64_t n0, int64_t n1, int S,
                       int width, int act, void* dH, int dh_dtype, cudaStream_t s) {
    const int64_t total = (n0 + n1) * width;
    if (total == 0) return GSAGE_OK;
    GS_CHECK_ARG(ldh >= width, "layer1_grad: ldh < width");
    layer1_grad_kernel<<<(unsigned)ceil_div(n0 + n1, 8), 256, 0, s>>>(dh0, dm2, H, h_dtype, ldh, n0, n1, S, width, act, dH,
                                                                     dh_dtype == GSAGE_BF16 ? 1 : 0);
    GS_LAUNCHED();
    return GSAGE_OK;
}

// bf16 rows: a block takes a strip of rows, a thread one column pair; fp32 atomics combine the strips
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, int64_t n, int d,
                                                          int64_t rows_per_block, float* __restrict__ out) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(n, r0 + rows_per_block);
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        float s = 0.0f;
        for (int64_t r = r0; r < r1; ++r) s += __bfloat162float(x[r * ld + c]);
        atomicAdd(out + c, s);
    }
}

int colsum_bf16_launch(const void* x, int64_t ld, int64_t n, int d, float* out, cudaStream_t s) {
    GS_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * d, s));
    if (n == 0) return GSAGE_OK;
    const int64_t rows_per_block = std::max<int64_t>(64, ceil_div(n, 148 * 8));
    colsum_bf16_kernel<<<(unsigned)ceil_div(n, rows_per_block), 256, 0, s>>>((const __nv_bfloat16*)x, ld, n, d, rows_per_block, out);
    GS_LAUNCHED();
    return GSAGE_OK;
}

int colsum_launch(const float* x, int64_t n, int d, float* out, cudaStream_t s) {
    colsum_kernel<<<d, 256, 0, s>>>(x, n, d, out);
    GS_LAUNCHED();
    return GSAGE_OK;
}

}  // namespace gsage

using namespace gsage;

// dW (O x d, fp32) = G^T . A[ids]  -- the weight gradient of a Linear whose input rows are A[ids] (or A in place) and
// whose output gradient is G.  exact != 0: fp32 FFMA kernel (G must be fp32).  exact == 0: tcgen05 kernel when the
// operands qualify (bf16 G and A, O == 128, 16-byte aligned rows), else an error -- never a silent fallback in tests.
extern "C" int gsage_wgrad(const void* g_dev, int g_dtype, int64_t ldg, int O, const void* a_dev, int a_dtype, int64_t lda,
                           int64_t n_table_rows, const int64_t* ids_dev, int d, int64_t n, float* dw_dev, int64_t lddw, int exact,
                           void* stream) {
    GS_CHECK_ARG(g_dev && a_dev && dw_dev && O > 0 && d > 0 && n >= 0 && lddw >= d && ldg >= O, "wgrad: bad arguments");
    cudaStream_t s = as_stream(stream);
    if (exact) {
        GS_CHECK_ARG(g_dtype == GSAGE_F32, "wgrad: the exact (FFMA) kernel takes an fp32 output gradient");
        return wgrad_launch((const float*)g_dev, ldg, O, a_dev, a_dtype, lda, ids_dev, d, n, dw_dev, lddw, s);
    }
    WgradJob j{g_dev, g_dtype, ldg, O, a_dev, a_dtype, lda, ids_dev, d, n, dw_dev, lddw, ids_dev ? n_table_rows : 0};
    GS_CHECK_ARG(wgrad_umma_eligible(j), "wgrad: operands do not qualify for the tensor-core kernel (bf16 G and A, O == 128, aligned rows)");
    return wgrad_umma_launch(&j, 1, s);
}
