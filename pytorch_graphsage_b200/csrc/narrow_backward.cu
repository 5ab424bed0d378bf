// narrow_backward.cu -- gradient kernels behind the NARROW plug-in API (aggregator `forward(x, neibs)`, prep `forward`).
//
// The reference back-propagates `loss.backward()` (/root/reference/models.py:100-101) through its plug-ins with torch
// autograd: MeanAggregator (nn_modules.py:196-204), PoolAggregator (:223-232), AttentionAggregator (:305-321),
// NodeEmbeddingPrep / LinearPrep (:144-155, :165-166), F.normalize + fc (models.py:90-91).  operators.py wraps its library
// calls in torch.autograd.Functions whose backward passes are these kernels plus gsage_wgrad (weight gradients) and
// gsage_linear with a transposed weight (data gradients) -- all fp32, rows in place (the narrow API receives gathered rows).
#include "backward.cuh"

namespace gsage {

// dpre[r, c] = dout[r, c] * act'(out[r, c])   (out is the POST-activation value: relu' = [out > 0], tanh' = 1 - out^2)
__global__ void __launch_bounds__(256) act_backward_kernel(const float* __restrict__ dout, int64_t ld_dout, const float* __restrict__ out,
                                                           int64_t ld_out, int64_t n, int width, int act, float* __restrict__ dpre, int64_t ld_dpre) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * width) return;
    const int64_t r = i / width;
    const int c = (int)(i - r * width);
    float g = dout[r * ld_dout + c];
    const float h = out[r * ld_out + c];
    if (act == GSAGE_ACT_RELU) g = h > 0.0f ? g : 0.0f;
    else if (act == GSAGE_ACT_TANH) g *= (1.0f - h * h);
    dpre[r * ld_dpre + c] = g;
}

// dst[p*S + j, :] = scale * src[p, :]   (the mean over S rows, backwards: scale = 1/S)
__global__ void __launch_bounds__(256) segment_broadcast_kernel(const float* __restrict__ src, int64_t ld_src, int64_t n, int d, int S, float scale,
                                                                float* __restrict__ dst, int64_t ld_dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * S * d) return;
    const int64_t r = i / d;
    const int c = (int)(i - r * d);
    dst[r * ld_dst + c] = scale * src[(r / S) * ld_src + c];
}

// max over the S rows of a parent, backwards: the FIRST row that attains the maximum receives the gradient (torch.max(dim)
// returns that index on ties), every other row zero.  One thread per (parent, column).
__global__ void __launch_bounds__(256) segment_max_backward_kernel(const float* __restrict__ h, int64_t ld_h, const float* __restrict__ dpooled,
                                                                   int64_t ld_dp, int64_t n, int S, int H, float* __restrict__ dh, int64_t ld_dh) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * H) return;
    const int64_t p = i / H;
    const int c = (int)(i - p * H);
    float best = -INFINITY; int arg = 0;
    for (int j = 0; j < S; ++j) {
        const float v = h[(p * S + j) * ld_h + c];
        if (v > best) { best = v; arg = j; }
    }
    const float g = dpooled[p * ld_dp + c];
    for (int j = 0; j < S; ++j) dh[(p * S + j) * ld_dh + c] = j == arg ? g : 0.0f;
}

// attention reduction, backwards, for fp32 rows in place.  One warp per parent.
//   m_p = sum_j w_j n_j ;  w = softmax_j(s_j) ;  s_j = <na_j, xa_p>
//   dw_j = <dM_p, n_j> ;  ds_j = w_j (dw_j - sum_k w_k dw_k) ;  dN_j = w_j dM_p (the direct path) ;
//   dNA_j = ds_j xa_p ;  dXA_p = sum_j ds_j na_j
__global__ void __launch_bounds__(256) attention_sum_backward_kernel(const float* __restrict__ nb, int64_t ld_nb, int d, int64_t n, int S,
                                                                     const float* __restrict__ dM, int64_t ld_dm, const float* __restrict__ w,
                                                                     const float* __restrict__ na, const float* __restrict__ xa, int H,
                                                                     float* __restrict__ dN, int64_t ld_dn, float* __restrict__ dNA,
                                                                     float* __restrict__ dXA, float* __restrict__ dw_scratch) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (p >= n) return;
    float dot = 0.0f;                                   // sum_k w_k dw_k
    for (int j = 0; j < S; ++j) {
        const int64_t r = p * S + j;
        float acc = 0.0f;
        const float wj = w[r];
        for (int c = lane; c < d; c += 32) {
            const float g = dM[p * ld_dm + c];
            acc = fmaf(g, nb[r * ld_nb + c], acc);
            dN[r * ld_dn + c] = wj * g;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
        if (lane == 0) dw_scratch[r] = acc;
        dot = fmaf(wj, acc, dot);
    }
    __syncwarp();
    for (int h0 = 0; h0 < H; h0 += 32) {
        const int hh = h0 + lane;
        const float xav = hh < H ? xa[p * H + hh] : 0.0f;
        float dxa = 0.0f;
        for (int j = 0; j < S; ++j) {
            const int64_t r = p * S + j;
            const float ds = w[r] * (dw_scratch[r] - dot);
            if (hh < H) {
                dNA[r * H + hh] = ds * xav;
                dxa = fmaf(ds, na[r * H + hh], dxa);
            }
        }
        if (hh < H) dXA[p * H + hh] = dxa;
    }
}

}  // namespace gsage

using namespace gsage;

extern "C" {

int gsage_act_backward(const float* dout_dev, int64_t ld_dout, const float* out_dev, int64_t ld_out, int64_t n, int width, int act,
                       float* dpre_dev, int64_t ld_dpre, void* stream) {
    GS_CHECK_ARG(dout_dev && out_dev && dpre_dev && n >= 0 && width > 0 && ld_dout >= width && ld_out >= width && ld_dpre >= width,
                 "act_backward: bad arguments");
    if (n == 0) return GSAGE_OK;
    act_backward_kernel<<<(unsigned)ceil_div(n * width, 256), 256, 0, as_stream(stream)>>>(dout_dev, ld_dout, out_dev, ld_out, n, width, act, dpre_dev, ld_dpre);
    GS_LAUNCHED();
    return GSAGE_OK;
}

int gsage_segment_broadcast(const float* src_dev, int64_t ld_src, int64_t n, int d, int S, float scale, float* dst_dev, int64_t ld_dst, void* stream) {
    GS_CHECK_ARG(src_dev && dst_dev && n >= 0 && d > 0 && S > 0 && ld_src >= d && ld_dst >= d, "segment_broadcast: bad arguments");
    if (n == 0) return GSAGE_OK;
    segment_broadcast_kernel<<<(unsigned)ceil_div(n * S * d, 256), 256, 0, as_stream(stream)>>>(src_dev, ld_src, n, d, S, scale, dst_dev, ld_dst);
    GS_LAUNCHED();
    return GSAGE_OK;
}

int gsage_segment_max_backward(const float* h_dev, int64_t ld_h, const float* dpooled_dev, int64_t ld_dp, int64_t n, int S, int H, float* dh_dev,
                               int64_t ld_dh, void* stream) {
    GS_CHECK_ARG(h_dev && dpooled_dev && dh_dev && n >= 0 && S > 0 && H > 0 && ld_h >= H && ld_dp >= H && ld_dh >= H, "segment_max_backward: bad arguments");
    if (n == 0) return GSAGE_OK;
    segment_max_backward_kernel<<<(unsigned)ceil_div(n * H, 256), 256, 0, as_stream(stream)>>>(h_dev, ld_h, dpooled_dev, ld_dp, n, S, H, dh_dev, ld_dh);
    GS_LAUNCHED();
    return GSAGE_OK;
}

int gsage_attention_sum_backward(const float* neibs_dev, int64_t ld_nb, int d, int64_t n, int S, const float* dm_dev, int64_t ld_dm,
                                 const float* w_dev, const float* na_dev, const float* xa_dev, int H, float* dn_dev, int64_t ld_dn,
                                 float* dna_dev, float* dxa_dev, float* scratch_dev, void* stream) {
    GS_CHECK_ARG(neibs_dev && dm_dev && w_dev && na_dev && xa_dev && dn_dev && dna_dev && dxa_dev && scratch_dev && n >= 0 && S > 0 && d > 0 && H > 0 &&
                 ld_nb >= d && ld_dm >= d && ld_dn >= d, "attention_sum_backward: bad arguments");
    if (n == 0) return GSAGE_OK;
    attention_sum_backward_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, as_stream(stream)>>>(neibs_dev, ld_nb, d, n, S, dm_dev, ld_dm, w_dev, na_dev, xa_dev,
                                                                                            H, dn_dev, ld_dn, dna_dev, dxa_dev, scratch_dev);
    GS_LAUNCHED();
    return GSAGE_OK;
}

int gsage_colsum(const float* x_dev, int64_t n, int d, float* out_dev, void* stream) {
    GS_CHECK_ARG(x_dev && out_dev && n >= 0 && d > 0, "colsum: bad arguments");
    return colsum_launch(x_dev, n, d, out_dev, as_stream(stream));
}

int gsage_embedding_backward(const float* drows_dev, int64_t ld, int d, const int64_t* ids_dev, int64_t n_ids, float* table_grad_dev,
                             int64_t ld_table, int64_t table_rows, void* stream) {
    GS_CHECK_ARG(drows_dev && ids_dev && table_grad_dev && n_ids >= 0 && d > 0 && ld >= d && ld_table >= d, "embedding_backward: bad arguments");
    return embedding_scatter_launch(drows_dev, ld, d, ids_dev, n_ids, 1, 1.0f, table_grad_dev, ld_table, table_rows, as_stream(stream));
}

int gsage_l2_normalize_backward(const float* z_dev, const float* dzn_dev, int64_t n, int d, float* dz_dev, void* stream) {
    GS_CHECK_ARG(z_dev && dzn_dev && dz_dev && n >= 0 && d > 0, "l2_normalize_backward: bad arguments");
    if (n == 0) return GSAGE_OK;
    return l2_normalize_bwd_launch(z_dev, dzn_dev, n, d, GSAGE_ACT_NONE, dz_dev, as_stream(stream));
}

}  // extern "C"
