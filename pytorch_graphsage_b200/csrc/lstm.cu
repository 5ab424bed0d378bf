// lstm.cu -- the pointwise half of the LSTM aggregator (nn_modules.py:259-286).  The two matrix products of every time step
// (x_t . W_ih^T and h_{t-1} . W_hh^T, 4H gate columns each) run on the projection kernels; this file holds what is left:
// the cell update.  The input half of the gates is computed for ALL S steps of a block of parents in one projection (the
// rows of a parent are adjacent, so step t of parent p is row p*S + t: a row stride of S*4H for the cell kernel).
#include "common.cuh"

namespace gsage {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// torch's cell, gate order (i, f, g, o) along the 4H gate columns:
//   c' = sigmoid(f) c + sigmoid(i) tanh(g),   h' = sigmoid(o) tanh(c')
// one thread per (row, 4 hidden units): float4 loads of the four gate quarters of gx (+ gh) and of both biases
template <typename HT>
__global__ void __launch_bounds__(256) lstm_cell_kernel(const float* __restrict__ gx, int64_t ldgx, const float* __restrict__ gh, int64_t ldgh,
                                                        const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                                        float* __restrict__ c, HT* __restrict__ h, int64_t ldh, int64_t n, int H,
                                                        int first) {
    const int q = H >> 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * q) return;
    const int64_t r = i / q;
    const int u = (int)(i % q) * 4;
    float g4[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(gx + r * ldgx + (int64_t)k * H + u);
        const float4 bi = __ldg(reinterpret_cast<const float4*>(b_ih + k * H + u));
        const float4 bh = __ldg(reinterpret_cast<const float4*>(b_hh + k * H + u));
        g4[k][0] = a.x + bi.x + bh.x; g4[k][1] = a.y + bi.y + bh.y; g4[k][2] = a.z + bi.z + bh.z; g4[k][3] = a.w + bi.w + bh.w;
        if (!first) {
            const float4 b = *reinterpret_cast<const float4*>(gh + r * ldgh + (int64_t)k * H + u);
            g4[k][0] += b.x; g4[k][1] += b.y; g4[k][2] += b.z; g4[k][3] += b.w;
        }
    }
    float4 cv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(c + r * (int64_t)H + u);
    float cc[4] = {cv.x, cv.y, cv.z, cv.w}, hh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        cc[j] = sigmoidf_(g4[1][j]) * cc[j] + sigmoidf_(g4[0][j]) * tanhf(g4[2][j]);
        hh[j] = sigmoidf_(g4[3][j]) * tanhf(cc[j]);
    }
    *reinterpret_cast<float4*>(c + r * (int64_t)H + u) = make_float4(cc[0], cc[1], cc[2], cc[3]);
    HT* ho = h + r * ldh + u;
#pragma unroll
    for (int j = 0; j < 4; ++j) ElemTraits<HT>::store(ho + j, hh[j]);
}

// the cell, backwards (one time step of back-propagation through time).  Recomputes the gate activations from the same
// pre-activations the forward saw (gx + gh + biases) and the previous cell state:
//   do = dh tanh(c);  dc += dh o (1 - tanh(c)^2);  di = dc g;  dg = dc i;  df = dc c_prev;  dc_prev = dc f
//   d(pre-activation) = (di i(1-i), df f(1-f), dg (1-g^2), do o(1-o))  -> dgates (n x 4H, same gate order)
// dc holds d loss / d c_t on entry and d loss / d c_{t-1} on exit (in place).
__global__ void __launch_bounds__(256) lstm_cell_backward_kernel(const float* __restrict__ gx, int64_t ldgx, const float* __restrict__ gh, int64_t ldgh,
                                                                 const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                                                 const float* __restrict__ c_prev, const float* __restrict__ dh,
                                                                 float* __restrict__ dc, float* __restrict__ dgates, int64_t ldg, int64_t n, int H,
                                                                 int first) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * H) return;
    const int64_t r = idx / H;
    const int u = (int)(idx - r * H);
    float a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[k] = gx[r * ldgx + (int64_t)k * H + u] + b_ih[k * H + u] + b_hh[k * H + u];
        if (!first) a[k] += gh[r * ldgh + (int64_t)k * H + u];
    }
    const float i = sigmoidf_(a[0]), f = sigmoidf_(a[1]), g = tanhf(a[2]), o = sigmoidf_(a[3]);
    const float cp = first ? 0.0f : c_prev[r * (int64_t)H + u];
    const float c = f * cp + i * g;
    const float tc = tanhf(c);
    const float dhv = dh[r * (int64_t)H + u];
    const float dcv = dc[r * (int64_t)H + u] + dhv * o * (1.0f - tc * tc);
    float* out = dgates + r * ldg + u;
    out[0] = dcv * g * i * (1.0f - i);
    out[H] = dcv * cp * f * (1.0f - f);
    out[2 * (int64_t)H] = dcv * i * (1.0f - g * g);
    out[3 * (int64_t)H] = dhv * tc * o * (1.0f - o);
    dc[r * (int64_t)H + u] = dcv * f;
}

}  // namespace gsage

using namespace gsage;

extern "C" int gsage_lstm_cell_backward(const float* gx_dev, int64_t ldgx, const float* gh_dev, int64_t ldgh, const float* b_ih_dev,
                                        const float* b_hh_dev, const float* c_prev_dev, const float* dh_dev, float* dc_dev, float* dgates_dev,
                                        int64_t ldg, int64_t n, int H, int first, void* stream) {
    GS_CHECK_ARG(gx_dev && b_ih_dev && b_hh_dev && dh_dev && dc_dev && dgates_dev && n >= 0 && H > 0, "lstm_cell_backward: bad arguments");
    GS_CHECK_ARG(first || (gh_dev && c_prev_dev), "lstm_cell_backward: the recurrent gates / previous cell state are missing");
    GS_CHECK_ARG(ldgx >= 4 * (int64_t)H && ldg >= 4 * (int64_t)H && (first || ldgh >= 4 * (int64_t)H), "lstm_cell_backward: bad row strides");
    if (n == 0) return GSAGE_OK;
    lstm_cell_backward_kernel<<<(unsigned)ceil_div(n * H, 256), 256, 0, as_stream(stream)>>>(gx_dev, ldgx, gh_dev, ldgh, b_ih_dev, b_hh_dev, c_prev_dev,
                                                                                              dh_dev, dc_dev, dgates_dev, ldg, n, H, first ? 1 : 0);
    GS_LAUNCHED();
    return GSAGE_OK;
}

extern "C" int gsage_lstm_cell(const float* gx_dev, int64_t ldgx, const float* gh_dev, int64_t ldgh, const float* b_ih_dev, const float* b_hh_dev,
                               float* c_dev, void* h_dev, int h_dtype, int64_t ldh, int64_t n, int H, int first, void* stream) {
    GS_CHECK_ARG(gx_dev && b_ih_dev && b_hh_dev && c_dev && h_dev && n >= 0 && H > 0, "lstm_cell: bad arguments");
    GS_CHECK_ARG(first || gh_dev, "lstm_cell: the recurrent gates are missing");
    GS_CHECK_ARG(H % 4 == 0 && ldgx % 4 == 0 && ldgx >= 4 * (int64_t)H && ldh >= H, "lstm_cell: H and the gate row strides must be multiples of 4");
    GS_CHECK_ARG(first || (ldgh % 4 == 0 && ldgh >= 4 * (int64_t)H), "lstm_cell: bad row stride of the recurrent gates");
    GS_CHECK_ARG(h_dtype == GSAGE_F32 || h_dtype == GSAGE_BF16, "lstm_cell: bad h dtype");
    GS_CHECK_ARG(((uintptr_t)gx_dev & 15) == 0 && ((uintptr_t)gh_dev & 15) == 0 && ((uintptr_t)c_dev & 15) == 0 &&
                 ((uintptr_t)b_ih_dev & 15) == 0 && ((uintptr_t)b_hh_dev & 15) == 0, "lstm_cell: operands must be 16-byte aligned");
    if (n == 0) return GSAGE_OK;
    cudaStream_t s = as_stream(stream);
    const unsigned grid = (unsigned)ceil_div(n * (H / 4), 256);
    if (h_dtype == GSAGE_F32)
        lstm_cell_kernel<float><<<grid, 256, 0, s>>>(gx_dev, ldgx, gh_dev, ldgh, b_ih_dev, b_hh_dev, c_dev, (float*)h_dev, ldh, n, H, first ? 1 : 0);
    else
        lstm_cell_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(gx_dev, ldgx, gh_dev, ldgh, b_ih_dev, b_hh_dev, c_dev, (__nv_bfloat16*)h_dev, ldh, n, H,
                                                             first ? 1 : 0);
    GS_LAUNCHED();
    return GSAGE_OK;
}
