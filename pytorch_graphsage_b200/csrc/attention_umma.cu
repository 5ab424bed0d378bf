// attention_umma.cu -- the attention aggregator's reduction as ONE kernel: every neighbour row is read from HBM once.
//
// Replaces  (/root/reference/nn_modules.py:307-315)
//     neib_att = self.att(neibs)                      att = Linear(d, 32, no bias) -> tanh -> Linear(32, 32, no bias)
//     ws       = softmax_j( bmm(neib_att.view(N, S, 32), x_att.view(N, 32, 1)) )
//     agg      = sum_j ws[:, j] * neibs.view(N, S, d)[:, j]
// i.e. per parent i:  m_i = sum_j softmax_j( <a(n_ij), a(x_i)> ) n_ij,  a(v) = W2 tanh(W1 v [+ b1]).
// The unfused chain (engine.cu round 1) read the neighbour rows twice (once for a(n), once for the weighted sum) and
// round-tripped three (N*S, 32) fp32 intermediates through HBM: ~3.3 GB per layer-1 application where 1.05 GB is the
// rows themselves.  Here a tile of R = floor(128/S)*S neighbour rows (whole parents) lands in shared memory once, by TMA
// (tile::gather4 for rows by id, all k-chunks of the tile), and everything else happens on chip:
//   warp  8     MMA issue   D1[128 rows, 32] = tile . W1^T on tcgen05 (W1 resident in smem, accumulator in TMEM)
//   warps 0-7   compute     tcgen05.ld -> +b1 -> tanh -> W2 (fp32 FFMA, W2 broadcast from smem) -> score against a(x_i)
//                           -> softmax over the S rows of each parent (scores exchanged through smem)
//                           -> weighted sum of the RAW rows, re-read from the swizzled smem tile -> m_i to HBM
//   warps 9-12  producers   TMA loads of the next tile into the other buffer while this one is being reduced
// a(x_i) (N x 32) is computed by the caller with the ordinary projection kernels: N rows, not N*S.
// bf16 operands only (the fp32-exact engine keeps the unfused FFMA chain); attention width H == 32; tanh.approx.f32.
#include "linear.cuh"
#include "umma_ptx.cuh"
#include <string.h>
#include <stdlib.h>
#include <float.h>

namespace gsage {

static constexpr int AH = 32;                    // attention width (nn_modules.py:293-297 hidden_dim = 32)
static constexpr int kAtComputeWarps = 8;
static constexpr int kAtTmaWarps = 4;
static constexpr int kAtThreads = 32 * (kAtComputeWarps + 1 + kAtTmaWarps);
static constexpr int kAtChunk = 128 * 128;       // one k-chunk of the tile: 128 rows x 128 bytes
static constexpr int kAtW1Chunk = AH * 128;      // one k-chunk of W1: 32 rows x 128 bytes
static constexpr int kAtSmemLimit = 227 * 1024;

struct AttParams {
    const void* a; int64_t lda; const int64_t* ids;      // neighbour rows: a[ids[r]] or a[r]
    const float* b1;                                     // 32 floats or NULL (folded prep bias)
    const float* w2;                                     // (32, 32) fp32 row-major (out, in)
    const float* xa;                                     // a(x_i): (n_parents, 32) fp32
    int d, S, R, PT;                                     // R = PT * S rows of a tile are used
    int64_t n_parents, n_rows;
    int kchunks, n_tiles, nbuf, buf_bytes;
    void* out; int out_bf16; int64_t ld_out;
    int* err;
};

struct AttMaps { CUtensorMap w1; CUtensorMap a; };

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(kAtThreads, 1) attention_fused_kernel(const AttParams P, const __grid_constant__ AttMaps M) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [nbuf x tile (kchunks x 16 KB)] [W1: kchunks x 4 KB] [W2 4 KB] [b1 128 B] [scores 2 x 128] [weights 128] [barriers]
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* w1s = smem + (size_t)P.nbuf * P.buf_bytes;
    float* w2s = (float*)(w1s + (size_t)P.kchunks * kAtW1Chunk);
    float* b1s = w2s + AH * AH;
    float* sc = b1s + AH;                                 // [2][128] partial scores (one per half of the W2 outputs)
    float* wt = sc + 256;                                 // [128] softmax weights
    uint64_t* bars = (uint64_t*)(wt + 128);
    uint32_t* tmem_slot = (uint32_t*)(bars + 8);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int b) { return bar_base + 8u * b; };          // tile b has landed (tx bytes)
    auto empty_bar = [&](int b) { return bar_base + 8u * (2 + b); };   // tile b (smem + accumulator) has been consumed
    auto dfull_bar = [&](int b) { return bar_base + 8u * (4 + b); };   // D1 of tile b is complete
    const uint32_t wfull_bar = bar_base + 8u * 6;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(full_bar(b), 1); mbar_init(empty_bar(b), 32 * kAtComputeWarps); mbar_init(dfull_bar(b), 1); }
        mbar_init(wfull_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kAtComputeWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp < kAtComputeWarps) {
        for (int i = threadIdx.x; i < AH * AH; i += 32 * kAtComputeWarps) w2s[i] = P.w2[i];
        if (threadIdx.x < AH) b1s[threadIdx.x] = P.b1 ? P.b1[threadIdx.x] : 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nbuf = P.nbuf;

    if (warp < kAtComputeWarps) {
        // =========================== COMPUTE ===========================
        const int quarter = warp & 3, half = warp >> 2;       // TMEM lane quarter; which 16 of the 32 W2 outputs
        const int row = quarter * 32 + lane;                  // tile row == TMEM lane
        const int S = P.S, PT = P.PT;
        const int my_parent = row / S;                        // parent of this row inside the tile (rows >= R are unused)
        const int nch = (P.d + 7) / 8;                        // 16-byte chunks per row
        int it = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
            const int b = nbuf == 2 ? (it & 1) : 0;
            const uint32_t par = nbuf == 2 ? ((it >> 1) & 1) : (it & 1);
            mbar_wait(dfull_bar(b), par, P.err);
            tc_fence_after();
            const int64_t parent0 = (int64_t)tile * PT;
            const bool live = row < P.R && parent0 + my_parent < P.n_parents;
            // ---- a(n) for this row: tanh(D1 + b1), then this half's 16 outputs of W2, dotted with a(x_parent) ----
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * AH), r);
            tmem_ld_wait();
            float t[AH];
#pragma unroll
            for (int k = 0; k < AH; ++k) t[k] = tanh_approx(__uint_as_float(r[k]) + b1s[k]);
            float part = 0.0f;
            if (live) {
                const float4* xav = reinterpret_cast<const float4*>(P.xa + (parent0 + my_parent) * AH + half * 16);
                float xa[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) { const float4 v = __ldg(xav + q); xa[4 * q] = v.x; xa[4 * q + 1] = v.y; xa[4 * q + 2] = v.z; xa[4 * q + 3] = v.w; }
                // 16 independent accumulators (the first version ran one 32-long dependent FFMA chain per output and was
                // latency-bound: 60 % of the kernel's samples); W2 rows come from smem as warp-uniform broadcast LDS.128
                float acc[16];
#pragma unroll
                for (int o = 0; o < 16; ++o) acc[o] = 0.0f;
                const uint32_t w2_u = smem_u32(w2s) + (uint32_t)(half * 16 * AH * 4);
#pragma unroll
                for (int q = 0; q < AH / 4; ++q) {
#pragma unroll
                    for (int o = 0; o < 16; ++o) {
                        float4 w;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w)
                                     : "r"(w2_u + (uint32_t)((o * AH + 4 * q) * 4)));
                        acc[o] = fmaf(w.x, t[4 * q], acc[o]); acc[o] = fmaf(w.y, t[4 * q + 1], acc[o]);
                        acc[o] = fmaf(w.z, t[4 * q + 2], acc[o]); acc[o] = fmaf(w.w, t[4 * q + 3], acc[o]);
                    }
                }
#pragma unroll
                for (int o = 0; o < 16; ++o) part = fmaf(acc[o], xa[o], part);
            }
            sc[half * 128 + row] = part;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // ---- softmax over the S rows of this row's parent (dummy rows are NOT masked: nn_modules.py:311) ----
            if (half == 0) {
                float w = 0.0f;
                if (live) {
                    const int first = my_parent * S;
                    float mx = -FLT_MAX;
                    for (int j = 0; j < S; ++j) mx = fmaxf(mx, sc[first + j] + sc[128 + first + j]);
                    float sum = 0.0f;
                    for (int j = 0; j < S; ++j) sum += __expf(sc[first + j] + sc[128 + first + j] - mx);
                    w = __expf(sc[row] + sc[128 + row] - mx) / sum;
                }
                wt[row] = w;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // ---- m_p = sum_j w_pj n_pj from the raw rows in the swizzled smem tile: item = (parent, 16-byte chunk) ----
            const uint32_t tile_u = smem_u32(smem + (size_t)b * P.buf_bytes);
            for (int i = threadIdx.x; i < PT * nch; i += 32 * kAtComputeWarps) {
                const int p = i / nch, c = i - p * nch;
                if (parent0 + p >= P.n_parents) continue;
                const int kc = c >> 3, c8 = c & 7;
                const uint32_t chunk_u = tile_u + (uint32_t)kc * kAtChunk;
                float acc[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
                for (int j = 0; j < S; ++j) {
                    const int rr = p * S + j;
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                                 : "r"(chunk_u + (uint32_t)rr * 128u + (uint32_t)((c8 ^ (rr & 7)) << 4)));
                    const float w = wt[rr];
                    float f[8];
                    ElemTraits<__nv_bfloat16>::unpack(v, f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, f[e], acc[e]);
                }
                const int64_t at = (parent0 + p) * P.ld_out + (int64_t)c * 8;
                if (P.out_bf16) {
                    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(P.out) + at) = ElemTraits<__nv_bfloat16>::pack(acc);
                } else {
                    float* o = reinterpret_cast<float*>(P.out) + at;
                    *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                }
            }
            tc_fence_before();
            mbar_arrive(empty_bar(b));                        // smem tile b and accumulator b are free again
        }
    } else if (warp == kAtComputeWarps) {
        // =========================== MMA ISSUER (one elected lane: umma_ptx.cuh, elect_one) ===========================
        if (elect_one()) {
            // D = f32, A = B = bf16, both K-major, N = 32, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AH >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t desc_hi = umma_desc(0);
            const uint32_t w16 = (smem_u32(w1s) & 0x3FFFF) >> 4;
            mbar_wait(wfull_bar, 0, P.err);
            int it = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
                const int b = nbuf == 2 ? (it & 1) : 0;
                const uint32_t par = nbuf == 2 ? ((it >> 1) & 1) : (it & 1);
                mbar_wait(full_bar(b), par, P.err);           // (the producers only refill b after the compute warps released it)
                tc_fence_after();
                const uint32_t a16 = (smem_u32(smem + (size_t)b * P.buf_bytes) & 0x3FFFF) >> 4;
                const uint32_t d_tmem = tmem_base + (uint32_t)(b * AH);
                for (int kc = 0; kc < P.kchunks; ++kc) {
                    const uint64_t adesc = desc_hi | (uint64_t)(a16 + kc * (kAtChunk >> 4));
                    const uint64_t bdesc = desc_hi | (uint64_t)(w16 + kc * (kAtW1Chunk >> 4));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                }
                umma_commit(dfull_bar(b));
            }
        }
        __syncwarp();
    } else {
        // =========================== TMA PRODUCERS ===========================
        const int pw = warp - (kAtComputeWarps + 1);
        const bool lead = pw == 0 && lane == 0;
        if (lead) {
            mbar_arrive_expect_tx(wfull_bar, (uint32_t)(P.kchunks * kAtW1Chunk));
            for (int kc = 0; kc < P.kchunks; ++kc) tma_load_2d(smem_u32(w1s) + kc * kAtW1Chunk, &M.w1, kc * 64, 0, wfull_bar);
        }
        const int my_row = 32 * pw + 4 * lane;                // gather: lanes 0..7 own tile rows my_row .. my_row + 3
        int it = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
            const int b = nbuf == 2 ? (it & 1) : 0;
            const uint32_t par = (nbuf == 2 ? ((it >> 1) & 1) : (it & 1)) ^ 1;
            const int64_t row0 = (int64_t)tile * P.R;
            int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
            if (P.ids && lane < 8) {
                const int64_t base = row0 + my_row;
                if (my_row + 0 < P.R && base + 0 < P.n_rows) r0 = (int)__ldg(P.ids + base + 0);
                if (my_row + 1 < P.R && base + 1 < P.n_rows) r1 = (int)__ldg(P.ids + base + 1);
                if (my_row + 2 < P.R && base + 2 < P.n_rows) r2 = (int)__ldg(P.ids + base + 2);
                if (my_row + 3 < P.R && base + 3 < P.n_rows) r3 = (int)__ldg(P.ids + base + 3);
            }
            if (P.ids || lead) mbar_wait(empty_bar(b), par, P.err);
            const uint32_t tile_u = smem_u32(smem + (size_t)b * P.buf_bytes), fb = full_bar(b);
            if (lead) {
                mbar_arrive_expect_tx(fb, (uint32_t)(P.kchunks * kAtChunk));
                if (!P.ids)
                    for (int kc = 0; kc < P.kchunks; ++kc) tma_load_2d(tile_u + kc * kAtChunk, &M.a, kc * 64, (int)row0, fb);
            }
            if (P.ids && lane < 8)
                for (int kc = 0; kc < P.kchunks; ++kc)
                    tma_gather4(tile_u + kc * kAtChunk + (uint32_t)my_row * 128u, &M.a, kc * 64, r0, r1, r2, r3, fb);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kAtComputeWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
    }
}

// ---- host -------------------------------------------------------------------------------------------------
static bool at_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool attention_fused_eligible(const void* a, int a_dtype, int64_t lda, int d, const void* w1, int w1_dtype, int64_t ldw, int H, int S,
                              int64_t n_parents, const void* out, int64_t ld_out, int out_dtype) {
    if (getenv("GSAGE_NO_FUSED_ATTENTION")) return false;
    if (a_dtype != GSAGE_BF16 || w1_dtype != GSAGE_BF16 || H != AH || S < 2 || S > 128 || n_parents < 1 || d < 8) return false;
    if (!at_aligned16(a) || !at_aligned16(w1) || !at_aligned16(out) || (lda * 2) % 16 != 0 || (ldw * 2) % 16 != 0) return false;
    if (lda < (d + 7) / 8 * 8 || ldw < (d + 7) / 8 * 8) return false;
    const int64_t es_out = out_dtype == GSAGE_BF16 ? 2 : 4;
    if ((ld_out * es_out) % 16 != 0 || ld_out < (d + 7) / 8 * 8) return false;           // whole 16-byte chunks are stored
    const int kchunks = (d + 63) / 64;
    const int fixed = 1024 + kchunks * kAtW1Chunk + AH * AH * 4 + AH * 4 + 384 * 4 + 256;
    return kchunks * kAtChunk + fixed <= kAtSmemLimit && n_parents * (int64_t)S < (1LL << 31);
}

static int* g_att_err = nullptr;

int attention_fused_launch(const void* a, int64_t lda, const int64_t* ids, int d, const void* w1, int64_t ldw, const float* b1,
                           const float* w2, const float* xa, int64_t n_parents, int S, void* out, int out_dtype, int64_t ld_out,
                           cudaStream_t s, int64_t table_rows) {
    AttParams U;
    memset(&U, 0, sizeof(U));
    U.a = a; U.lda = lda; U.ids = ids; U.b1 = b1; U.w2 = w2; U.xa = xa;
    U.d = d; U.S = S; U.PT = 128 / S; U.R = U.PT * S;
    U.n_parents = n_parents; U.n_rows = n_parents * S;
    U.kchunks = (d + 63) / 64;
    U.n_tiles = (int)ceil_div(n_parents, U.PT);
    U.buf_bytes = U.kchunks * kAtChunk;
    const int fixed = 1024 + U.kchunks * kAtW1Chunk + AH * AH * 4 + AH * 4 + 384 * 4 + 256;
    U.nbuf = (2 * U.buf_bytes + fixed <= kAtSmemLimit) ? 2 : 1;
    U.out = out; U.out_bf16 = out_dtype == GSAGE_BF16; U.ld_out = ld_out;
    if (!g_att_err) {
        GS_CUDA(cudaMalloc((void**)&g_att_err, sizeof(int)));
        GS_CUDA(cudaMemset(g_att_err, 0, sizeof(int)));
    }
    U.err = g_att_err;
    AttMaps maps;
    memset(&maps, 0, sizeof(maps));
    GS_TRY(make_map(&maps.w1, w1, AH, d, ldw, AH, 2));
    if (ids) GS_TRY(make_map(&maps.a, a, table_rows > 0 ? table_rows : 0x7FFFFFFF, d, lda, 1, 2));
    else GS_TRY(make_map(&maps.a, a, U.n_rows, d, lda, 128, 2));
    const size_t smem = (size_t)U.nbuf * U.buf_bytes + fixed;
    static bool attr_set = false;
    if (!attr_set) {
        GS_CUDA(cudaFuncSetAttribute(attention_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemLimit));
        attr_set = true;
    }
    // grid = waves x SMs, one CTA resident per SM at a time (profiling knob): > 1 hands the tiles out in smaller static shares; measured slower for this kernel
    // (profiles/r02_persistent_waves.txt), unlike the single-phase projections (linear_ws_umma.cu)
    static const int waves = getenv("GSAGE_ATT_WAVES") ? atoi(getenv("GSAGE_ATT_WAVES")) : 1;
    const int slots = sm_count() * (waves < 1 ? 1 : waves);
    const int grid = U.n_tiles < slots ? U.n_tiles : slots;
    attention_fused_kernel<<<grid, kAtThreads, smem, s>>>(U, maps);
    GS_LAUNCHED();
    return GSAGE_OK;
}

}  // namespace gsage

using namespace gsage;

// m[p, :d] = sum_j softmax_j(<a(n_pj), xa[p]>) n_pj  with a(v) = W2 tanh(W1 v + b1): the attention aggregator's reduction
// (nn_modules.py:307-315) in one launch.  n_pj = table[ids[p*S + j]] (ids NULL: row p*S + j of `table` itself).
// bf16 table and W1, attention width 32; GSAGE_ERR_INVALID for anything else (no silent fallback).
extern "C" int gsage_attention_aggregate(const void* table_dev, int dtype, int64_t ld, int64_t n_table_rows, int d,
                                         const int64_t* ids_dev, int64_t n_parents, int S, const void* w1_dev, int w1_dtype, int64_t ldw, int H, const float* b1_dev,
                                         const float* w2_dev, const float* xa_dev, void* out_dev, int out_dtype, int64_t ld_out,
                                         void* stream) {
    GS_CHECK_ARG(table_dev && w1_dev && w2_dev && xa_dev && out_dev, "attention_aggregate: NULL argument");
    GS_CHECK_ARG(S > 1, "attention aggregator: S must be > 1 (the reference's squeeze() is ill-defined at S == 1)");
    if (n_parents == 0) return GSAGE_OK;
    GS_CHECK_ARG(attention_fused_eligible(table_dev, dtype, ld, d, w1_dev, w1_dtype, ldw, H, S, n_parents, out_dev, ld_out, out_dtype),
                 "attention_aggregate: needs a bf16 table and W1 with 16-byte aligned rows, attention width 32, 2 <= S <= 128, "
                 "an output whose rows hold whole 16-byte chunks");
    return attention_fused_launch(table_dev, ld, ids_dev, d, w1_dev, ldw, b1_dev, w2_dev, xa_dev, n_parents, S, out_dev, out_dtype,
                                  ld_out, as_stream(stream), ids_dev ? n_table_rows : 0);
}
