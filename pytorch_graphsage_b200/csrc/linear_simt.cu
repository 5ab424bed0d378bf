// linear_simt.cu -- fp32-exact projection kernel (FFMA, fp32 accumulate): the parity-grade path.
//
// Replaces the nn.Linear calls on the hot path: fc_x / fc_neib + concat (/root/reference/nn_modules.py:200,
// 228, 317), the pool MLP (:224), the attention MLP (:307-308), the prep affines (:150,166) and the final
// classifier (/root/reference/models.py:91).
//
//   out[r, col0 + o] = act( sum_k A[row(r), k] * W[o, k] + bias[o] ),   row(r) = ids ? ids[r] : r
//
// Up to two segments per launch (blockIdx.z): the reference's `torch.cat([fc_x(x), fc_neib(agg)], dim=1)`
// is one launch writing two column ranges of one buffer -- the concat never exists as a copy.
// 64x64 output tile per CTA, 16-wide k steps through shared memory, 4x4 register tile per thread.
// The self rows are gathered straight from the feature table by id inside the A-tile load.
#include "linear.cuh"

namespace gsage {

static constexpr int BM = 64, BN = 64, BK = 16;

template <typename T>
__device__ __forceinline__ float load_elem(const void* base, int64_t idx) {
    return ElemTraits<T>::load(reinterpret_cast<const T*>(base) + idx);
}

__device__ __forceinline__ float load_any(const void* base, int dtype, int64_t idx) {
    return dtype == GSAGE_BF16 ? load_elem<__nv_bfloat16>(base, idx) : load_elem<float>(base, idx);
}

__global__ void __launch_bounds__(256) linear_simt_kernel(LinearParams P) {
    const LinearSeg& sg = P.seg[blockIdx.z];
    if ((int)(blockIdx.y * BN) >= sg.O) return;
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;              // 16 x 16 threads, 4x4 outputs each
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int col0 = blockIdx.y * BN;

    // loader mapping: 64 rows x 16 k = 1024 elements, 4 per thread (same k-quad)
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    int64_t a_row = -1;
    const int S = sg.S > 1 ? sg.S : 1;
    if (row0 + lrow < P.n) {
        a_row = row0 + lrow;
        if (S == 1 && sg.ids) a_row = sg.ids[a_row];
    }
    const float inv_S = 1.0f / (float)S;
    const int w_row = col0 + lrow;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < sg.d; k0 += BK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + lk + q;
            float a = 0.0f, w = 0.0f;
            if (a_row >= 0 && k < sg.d) {
                if (S == 1) {
                    a = load_any(sg.a, sg.a_dtype, a_row * sg.lda + k);
                } else {                                       // fused gather+mean operand (fp32, same order as gather_reduce)
                    for (int j = 0; j < S; ++j) {
                        const int64_t q = a_row * S + j;
                        a += load_any(sg.a, sg.a_dtype, (sg.ids ? sg.ids[q] : q) * sg.lda + k);
                    }
                    a *= inv_S;
                }
            }
            if (w_row < sg.O && k < sg.d) w = load_any(sg.w, sg.w_dtype, sg.w_trans ? (int64_t)k * sg.ldw + w_row : (int64_t)w_row * sg.ldw + k);
            As[lk + q][lrow] = a;
            Ws[lk + q][lrow] = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 w4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = row0 + ty * 4 + i;
        if (r >= P.n) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = col0 + tx * 4 + j;
            if (o >= sg.O) continue;
            float v = acc[i][j];
            if (sg.bias) v += sg.bias[o];
            v = apply_act(v, P.act);
            const int64_t at = r * P.ld_out + sg.col0 + o;
            if (P.out_dtype == GSAGE_BF16) reinterpret_cast<__nv_bfloat16*>(P.out)[at] = __float2bfloat16_rn(v);
            else reinterpret_cast<float*>(P.out)[at] = v;
        }
    }
}

int linear_simt_launch(const LinearParams& P, cudaStream_t s) {
    int maxO = 0;
    for (int i = 0; i < P.n_segs; ++i) maxO = P.seg[i].O > maxO ? P.seg[i].O : maxO;
    dim3 grid((unsigned)ceil_div(P.n, BM), (unsigned)ceil_div(maxO, BN), (unsigned)P.n_segs);
    linear_simt_kernel<<<grid, 256, 0, s>>>(P);
    GS_LAUNCHED();
    return GSAGE_OK;
}

}  // namespace gsage

using namespace gsage;

// stream-ordered scratch for the (hi, lo) weight halves of exact == 2 calls: a private pool that keeps its memory between
// calls (the default pool hands everything back to the driver at every synchronisation point)
static cudaMemPool_t split_pool() {
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    cudaMemPool_t& pool = pools[dev & 63];                     // one pool per device of the process
    if (!pool) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); cudaDeviceGetDefaultMemPool(&pool, dev); return pool; }
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return pool;
}

extern "C" int gsage_linear(const gsage_linear_seg* segs, int n_segs, int64_t n, int act, void* out_dev, int out_dtype,
                            int64_t ld_out, int exact, void* stream) {
    GS_CHECK_ARG(segs && (n_segs == 1 || n_segs == 2) && n >= 0 && out_dev, "linear: bad arguments");
    GS_CHECK_ARG(out_dtype == GSAGE_F32 || out_dtype == GSAGE_BF16, "linear: bad out dtype");
    GS_CHECK_ARG(act == GSAGE_ACT_NONE || act == GSAGE_ACT_RELU || act == GSAGE_ACT_TANH, "linear: bad activation");
    LinearParams P;
    P.n_segs = n_segs; P.n = n; P.act = act; P.out = out_dev; P.out_dtype = out_dtype; P.ld_out = ld_out;
    for (int i = 0; i < n_segs; ++i) {
        const gsage_linear_seg& g = segs[i];
        GS_CHECK_ARG(g.a_dev && g.w_dev && g.d > 0 && g.O > 0 && g.lda >= g.d && g.ldw >= (g.w_transposed ? g.O : g.d) && g.col0 >= 0 &&
                     g.col0 + g.O <= ld_out, "linear: bad segment %d", i);
        P.seg[i].a = g.a_dev; P.seg[i].a_dtype = g.a_dtype; P.seg[i].lda = g.lda; P.seg[i].ids = g.ids_dev;
        P.seg[i].w = g.w_dev; P.seg[i].w_dtype = g.w_dtype; P.seg[i].ldw = g.ldw; P.seg[i].d = g.d; P.seg[i].O = g.O;
        P.seg[i].bias = g.bias_dev; P.seg[i].col0 = g.col0; P.seg[i].S = g.reduce_S > 1 ? g.reduce_S : 1; P.seg[i].w_trans = g.w_transposed ? 1 : 0;
        P.seg[i].a_rows = g.ids_dev ? g.a_rows : 0;
    }
    if (n == 0) return GSAGE_OK;
    cudaStream_t s = as_stream(stream);
    if (exact != 2) return linear_dispatch(P, exact, s);
    // exact == 2: fp32 accuracy on the tensor cores where the call qualifies (3 x TF32), the FFMA kernel where it does not.
    // The (hi, lo) halves of the weights live in stream-ordered scratch for the duration of the call.
    float* scratch[2] = {nullptr, nullptr};
    for (int i = 0; i < n_segs; ++i) {
        LinearSeg& g = P.seg[i];
        if (g.a_dtype != GSAGE_F32 || g.w_dtype != GSAGE_F32 || g.w_trans) continue;
        const int64_t cnt = (int64_t)g.O * g.ldw;
        GS_CUDA(cudaMallocFromPoolAsync((void**)&scratch[i], 2 * cnt * sizeof(float), split_pool(), s));
        GS_TRY(split_tf32_launch((const float*)g.w, cnt, scratch[i], scratch[i] + cnt, s));
        g.w_hi = scratch[i]; g.w_lo = scratch[i] + cnt;
    }
    const int st = linear_dispatch(P, 1, s);
    for (int i = 0; i < n_segs; ++i) if (scratch[i]) cudaFreeAsync(scratch[i], s);
    return st;
}

namespace gsage {
// bf16 x bf16 segments go to the tcgen05 kernel (<= 256 accumulator columns per launch: wide segments are
// split into column blocks, two-segment calls that do not fit are issued one segment at a time);
// anything else runs on the fp32 FFMA kernel -- and so does everything when `exact` is set, except fp32 calls that come with
// the (hi, lo) tf32 halves of W and fit the weight-stationary kernel: those run as 3 x TF32 (GSAGE_FP32_FFMA=1 turns that off).
int linear_dispatch(const LinearParams& P, int exact, cudaStream_t s) {
    if (P.pool_S > 1) {
        GS_CHECK_ARG(!exact && linear_pool_umma_eligible(P), "linear: the fused MLP+pool needs operands that qualify for the tensor-core kernel "
                     "(bf16 or fp32-as-TF32, 16-byte aligned rows, one segment)");
        if (linear_pool_ws_umma_eligible(P)) return linear_pool_ws_umma_launch(P, s);
        return linear_pool_umma_launch(P, s);
    }
    bool tc = !exact;
    for (int i = 0; i < P.n_segs; ++i) {
        tc = tc && !P.seg[i].w_trans;
        if (P.seg[i].O_store > 0 && P.seg[i].O_store < P.seg[i].O) {       // padded operand: only the weight-stationary kernel trims the store
            GS_CHECK_ARG(!exact && linear_ws_umma_eligible(P), "linear: a padded projection (O_store < O) needs the weight-stationary tensor-core kernel");
            return linear_ws_umma_launch(P, s);
        }
    }
    for (int i = 0; i < P.n_segs && tc; ++i) {
        LinearParams one = P;
        one.n_segs = 1; one.seg[0] = P.seg[i];
        if (one.seg[0].O > 256) one.seg[0].O = 256;
        tc = linear_umma_eligible(one);
    }
    if (exact && linear_ws_umma_x3_eligible(P)) return linear_ws_umma_x3_launch(P, s);     // fp32 accuracy on the tensor cores (3 x TF32)
    if (exact) {
        // 3 x TF32 needs W resident twice (hi + lo): a projection whose weights do not fit (the pool aggregators' 512-unit MLP, the
        // d = 602 layer) runs in blocks of 128 output columns, each with its own slice of the weights -- the rows are re-read per
        // block, which costs far less than the FFMA kernel this replaces (fp32-exact pool / attention / wide mean models)
        bool all_split = true;
        for (int i = 0; i < P.n_segs; ++i) all_split = all_split && P.seg[i].w_hi && P.seg[i].w_lo && !P.seg[i].w_trans && P.seg[i].S <= 1;
        if (all_split && !getenv("GSAGE_FP32_FFMA")) {
            bool any = false;
            for (int i = 0; i < P.n_segs; ++i) {
                const LinearSeg& g = P.seg[i];
                // the widest column block whose (hi, lo) weights fit: 128, else 64 / 32 (layer 2's 256-wide rows), else FFMA
                int blk = 0;
                for (int cand = 128; cand >= 32 && !blk; cand >>= 1) {
                    LinearParams probe = P;
                    probe.n_segs = 1; probe.seg[0] = g; probe.seg[0].O = g.O < cand ? g.O : cand;
                    if (linear_ws_umma_x3_eligible(probe)) blk = cand;
                }
                for (int c = 0; c < g.O; c += (blk ? blk : 128)) {
                    const int width = blk ? blk : 128;
                    LinearParams one = P;
                    one.n_segs = 1;
                    one.seg[0] = g;
                    one.seg[0].O = g.O - c < width ? g.O - c : width;
                    one.seg[0].w = (const char*)g.w + (size_t)c * g.ldw * sizeof(float);
                    one.seg[0].w_hi = (const char*)g.w_hi + (size_t)c * g.ldw * sizeof(float);
                    one.seg[0].w_lo = (const char*)g.w_lo + (size_t)c * g.ldw * sizeof(float);
                    one.seg[0].bias = g.bias ? g.bias + c : nullptr;
                    one.seg[0].col0 = g.col0 + c;
                    if (blk && linear_ws_umma_x3_eligible(one)) { GS_TRY(linear_ws_umma_x3_launch(one, s)); any = true; }
                    else GS_TRY(linear_simt_launch(one, s));
                }
            }
            (void)any;
            return GSAGE_OK;
        }
    }
    if (!tc) { GS_CHECK_ARG(P.pool_S <= 1, "linear: operands do not qualify for the tensor-core kernel (pooled epilogue)"); return linear_simt_launch(P, s); }
    if (linear_ws_umma_eligible(P)) return linear_ws_umma_launch(P, s);      // weights stationary in smem: half the L2 -> SM bytes
    if (linear_umma_eligible(P)) return linear_umma_launch(P, s);
    for (int i = 0; i < P.n_segs; ++i) {
        const LinearSeg& g = P.seg[i];
        for (int c = 0; c < g.O; c += 256) {
            LinearParams one = P;
            one.n_segs = 1;
            one.seg[0] = g;
            one.seg[0].O = g.O - c < 256 ? g.O - c : 256;
            one.seg[0].w = (const char*)g.w + (size_t)c * g.ldw * dtype_size(g.w_dtype);
            one.seg[0].bias = g.bias ? g.bias + c : nullptr;
            one.seg[0].col0 = g.col0 + c;
            if (linear_ws_umma_eligible(one)) GS_TRY(linear_ws_umma_launch(one, s));
            else if (linear_umma_eligible(one)) GS_TRY(linear_umma_launch(one, s));
            else { GS_CHECK_ARG(P.pool_S <= 1, "linear: column block does not qualify for the tensor-core kernel (pooled epilogue)"); GS_TRY(linear_simt_launch(one, s)); }
        }
    }
    return GSAGE_OK;
}
}  // namespace gsage

// out[p, col0 + o] = reduce_{j<S} act( A[row(p*S+j)] . W[o] + bias[o] )   -- the pool aggregators' MLP + pool in one launch
// (tensor-core kernel only: bf16 or fp32-as-TF32 operands with 16-byte aligned rows, O % 16 == 0)
extern "C" int gsage_linear_pooled(const gsage_linear_seg* seg, int64_t n_parents, int S, int reduce, int act, void* out_dev,
                                   int out_dtype, int64_t ld_out, void* stream) {
    GS_CHECK_ARG(seg && out_dev && n_parents >= 0 && S >= 1, "linear_pooled: bad arguments");
    GS_CHECK_ARG(reduce == GSAGE_RED_MAX || reduce == GSAGE_RED_MEAN, "linear_pooled: reduce must be max or mean");
    LinearParams P;
    P.n_segs = 1; P.n = n_parents * S; P.act = act; P.out = out_dev; P.out_dtype = out_dtype; P.ld_out = ld_out;
    P.seg[0] = LinearSeg{seg->a_dev, seg->a_dtype, seg->lda, seg->ids_dev, seg->w_dev, seg->w_dtype, seg->ldw, seg->d, seg->O,
                         seg->bias_dev, seg->col0};
    P.seg[0].a_rows = seg->ids_dev ? seg->a_rows : 0;
    P.pool_S = S; P.pool_max = reduce == GSAGE_RED_MAX ? 1 : 0;
    if (P.n == 0) return GSAGE_OK;
    if (S == 1) { P.pool_S = 1; return linear_dispatch(P, 0, as_stream(stream)); }
    return linear_dispatch(P, 0, as_stream(stream));
}
