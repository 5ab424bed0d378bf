// metrics.cu -- the per-batch training metric on the device.
//
// Replaces  problem.metric_fn(to_numpy(targets), to_numpy(preds))   (/root/reference/train.py:150) for the three tasks of
// /root/reference/problem.py:44-64 (sklearn micro / macro F1 for `classification` and `multilabel_classification`, mean
// absolute error for `regression_mae`): the reference copies the (B, n_classes) predictions to the host every batch and
// runs sklearn there; here one pass over the logits counts (tp, fp, fn) per label on the device and only the two scalars travel.
//
// F1 conventions (sklearn.metrics.f1_score, what the reference calls):
//   classification: y_pred = argmax over classes (first maximum, numpy.argmax); the label set is every class that occurs in
//       y_true or y_pred; macro = unweighted mean of the per-label F1 over THAT set; micro = global 2tp / (2tp + fp + fn).
//   multilabel:     y_pred = preds > 0; the label set is all L columns; a label with no true and no predicted positive has
//       F1 = 0 (sklearn's zero_division default, with a warning); macro = mean over the L columns.
#include "common.cuh"

namespace gsage {

// one warp per row: argmax of the C logits (first maximum), then tp / fp / fn counts.  counts = [tp[C] | fp[C] | fn[C]]
__global__ void __launch_bounds__(256) metric_argmax_count_kernel(const float* __restrict__ preds, int64_t ld, const int64_t* __restrict__ y,
                                                                  int64_t n, int C, unsigned long long* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return;
    const float* p = preds + r * ld;
    float best = -INFINITY; int arg = C;                    // NaN never wins (numpy would return the first NaN; logits are finite)
    for (int c = lane; c < C; c += 32) {
        const float v = p[c];
        if (v > best) { best = v; arg = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        const int oa = __shfl_xor_sync(0xFFFFFFFFu, arg, o);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) {
        if (arg >= C) arg = 0;
        const int64_t t = y[r];
        if (t == arg) atomicAdd(counts + arg, 1ULL);
        else {
            atomicAdd(counts + C + arg, 1ULL);                                   // predicted, not true
            if (t >= 0 && t < C) atomicAdd(counts + 2 * C + (int)t, 1ULL);       // true, not predicted
        }
    }
}

// multilabel: element (r, c) predicted iff preds > 0, true iff y != 0
__global__ void __launch_bounds__(256) metric_multilabel_count_kernel(const float* __restrict__ preds, int64_t ld, const float* __restrict__ y,
                                                                      int64_t ldy, int64_t n, int C, unsigned long long* __restrict__ counts) {
    const int c = blockIdx.y * 32 + (threadIdx.x & 31);
    unsigned long long tp = 0, fp = 0, fn = 0;
    if (c < C) {
        for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += (int64_t)gridDim.x * 8) {
            const bool pred = preds[r * ld + c] > 0.0f, tru = y[r * ldy + c] != 0.0f;
            tp += pred && tru; fp += pred && !tru; fn += !pred && tru;
        }
        if (tp) atomicAdd(counts + c, tp);
        if (fp) atomicAdd(counts + C + c, fp);
        if (fn) atomicAdd(counts + 2 * C + c, fn);
    }
}

// out[0] = micro F1, out[1] = macro F1.  present_only: average over the labels that occur (single-label classification)
__global__ void metric_f1_finalize_kernel(const unsigned long long* __restrict__ counts, int C, int present_only, double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double tp = 0, fp = 0, fn = 0, macro = 0;
    int labels = 0;
    for (int c = 0; c < C; ++c) {
        const double a = (double)counts[c], b = (double)counts[C + c], d = (double)counts[2 * C + c];
        tp += a; fp += b; fn += d;
        const double den = 2 * a + b + d;
        if (den > 0) { macro += 2 * a / den; ++labels; }
        else if (!present_only) ++labels;                  // multilabel: an empty label counts with F1 = 0
    }
    const double den = 2 * tp + fp + fn;
    out[0] = den > 0 ? 2 * tp / den : 0.0;
    out[1] = labels > 0 ? macro / labels : 0.0;
}

// sum_i |a_i - b_i| into out[0] (double), block-reduced
__global__ void __launch_bounds__(256) metric_abs_err_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, double* __restrict__ out) {
    __shared__ double part[8];
    double s = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) s += fabs((double)a[i] - (double)b[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += part[w];
        atomicAdd(out, t);
    }
}

__global__ void metric_scale_kernel(double* out, double scale) { if (threadIdx.x == 0 && blockIdx.x == 0) out[0] *= scale; }

}  // namespace gsage

using namespace gsage;

extern "C" {

int gsage_metric_f1(const float* preds_dev, int64_t ld, const void* targets_dev, int64_t ld_targets, int64_t n, int n_classes, int multilabel,
                    void* scratch_dev, double* out_dev, void* stream) {
    GS_CHECK_ARG(preds_dev && targets_dev && scratch_dev && out_dev && n >= 0 && n_classes > 0 && ld >= n_classes, "metric_f1: bad arguments");
    cudaStream_t s = as_stream(stream);
    unsigned long long* counts = (unsigned long long*)scratch_dev;
    GS_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 3 * (size_t)n_classes, s));
    if (n > 0) {
        if (multilabel) {
            GS_CHECK_ARG(ld_targets >= n_classes, "metric_f1: multilabel targets are a (n, n_classes) float matrix");
            const unsigned gx = (unsigned)std::min<int64_t>(ceil_div(n, 8), 4 * (int64_t)sm_count());
            metric_multilabel_count_kernel<<<dim3(gx, (unsigned)ceil_div(n_classes, 32)), 256, 0, s>>>(preds_dev, ld, (const float*)targets_dev, ld_targets, n,
                                                                                                          n_classes, counts);
        } else {
            metric_argmax_count_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, s>>>(preds_dev, ld, (const int64_t*)targets_dev, n, n_classes, counts);
        }
        GS_LAUNCHED();
    }
    metric_f1_finalize_kernel<<<1, 32, 0, s>>>(counts, n_classes, multilabel ? 0 : 1, out_dev);
    GS_LAUNCHED();
    return GSAGE_OK;
}

int gsage_metric_mae(const float* preds_dev, const float* targets_dev, int64_t n, double* out_dev, void* stream) {
    GS_CHECK_ARG(preds_dev && targets_dev && out_dev && n > 0, "metric_mae: bad arguments");
    cudaStream_t s = as_stream(stream);
    GS_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), s));
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256), 2 * (int64_t)sm_count());
    metric_abs_err_kernel<<<grid, 256, 0, s>>>(preds_dev, targets_dev, n, out_dev);
    GS_LAUNCHED();
    metric_scale_kernel<<<1, 32, 0, s>>>(out_dev, 1.0 / (double)n);
    GS_LAUNCHED();
    return GSAGE_OK;
}

}  // extern "C"
