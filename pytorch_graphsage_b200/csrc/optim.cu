// optim.cu -- clip_grad_norm + Adam over ONE flat parameter / gradient buffer.
//
// Replaces  torch.nn.utils.clip_grad_norm(self.parameters(), 5); self.optimizer.step()   (/root/reference/models.py:102-103)
// for models whose parameters and gradients live in one contiguous fp32 buffer each (parallel.FlatParams /
// FlatGradBucket).  Stock torch runs ~25 tiny launches per step for a 7-tensor model (one multi-tensor kernel per
// elementwise op of the norm, the clip and Adam) -- 0.3 ms of a 2.4 ms train step; here it is two:
//   sumsq_kernel   total gradient norm^2 (block partial sums -> one fp32 atomic each)
//   adam_kernel    clip coefficient min(1, max_norm / (norm + 1e-6)) folded into the gradient read, then torch.optim.Adam's
//                  update (L2 weight decay added to the gradient, bias-corrected moments, eps added after the square root)
// Same arithmetic as torch.optim.Adam(amsgrad=False, maximize=False) + clip_grad_norm_(norm_type=2) up to the order of
// the norm's summation.
#include "common.cuh"
#include <algorithm>
#include <math.h>

namespace gsage {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ out) {
    float s = 0.0f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            const float4 v = *reinterpret_cast<const float4*>(g + i);
            s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        } else {
            for (int64_t k = i; k < n; ++k) s += g[k] * g[k];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    __shared__ float ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += ws[w];
        atomicAdd(out, t);
    }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   int64_t n, const float* __restrict__ sumsq, float max_norm, float lr, float beta1,
                                                   float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt) {
    float coef = 1.0f;
    if (max_norm > 0.0f) coef = fminf(1.0f, max_norm / (sqrtf(*sumsq) + 1e-6f));      // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max = 1)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gc = g[i] * coef;
        g[i] = gc;                                                                   // the reference leaves the clipped gradient in p.grad
        const float gd = weight_decay != 0.0f ? fmaf(weight_decay, p[i], gc) : gc;
        const float mi = beta1 * m[i] + (1.0f - beta1) * gd;
        const float vi = beta2 * v[i] + (1.0f - beta2) * gd * gd;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= (lr / bc1) * (mi / denom);
    }
}

}  // namespace gsage

using namespace gsage;

// One optimiser step on flat fp32 buffers (all on the device): clip the gradient to `max_norm` (<= 0: no clipping), then
// Adam with step count `step` (1-based, for the bias corrections).  `scratch_dev`: one float, overwritten.
extern "C" int gsage_adam_step(float* param_dev, float* grad_dev, float* m_dev, float* v_dev, int64_t n, float lr, float beta1,
                               float beta2, float eps, float weight_decay, int64_t step, float max_norm, float* scratch_dev, void* stream) {
    GS_CHECK_ARG(param_dev && grad_dev && m_dev && v_dev && scratch_dev && n >= 0 && step >= 1, "adam_step: bad arguments");
    GS_CHECK_ARG((reinterpret_cast<uintptr_t>(grad_dev) & 15u) == 0, "adam_step: the gradient buffer must be 16-byte aligned");
    if (n == 0) return GSAGE_OK;
    cudaStream_t s = as_stream(stream);
    const int grid = (int)std::min<int64_t>(ceil_div(n, 256 * 4), (int64_t)sm_count() * 8);
    if (max_norm > 0.0f) {
        GS_CUDA(cudaMemsetAsync(scratch_dev, 0, sizeof(float), s));
        sumsq_kernel<<<grid, 256, 0, s>>>(grad_dev, n, scratch_dev);
        GS_LAUNCHED();
    }
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_kernel<<<(int)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 16), 256, 0, s>>>(
        param_dev, grad_dev, m_dev, v_dev, n, scratch_dev, max_norm, lr, beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2));
    GS_LAUNCHED();
    return GSAGE_OK;
}
