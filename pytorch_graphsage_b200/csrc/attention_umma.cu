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
//   warps 0-7   compute     u_p = W2^T a(x_p) for the tile's parents (while the tile is still landing)
//                           tcgen05.ld -> +b1 -> tanh -> score = <tanh(.), u_p>
//                           -> softmax over the S rows of each parent (scores exchanged through smem)
//                           -> weighted sum of the RAW rows, re-read from the swizzled smem tile -> m_i to HBM
// The score is  <W2 t, a(x)> = <t, W2^T a(x)>: W2 moves from the N*S neighbour rows onto the N parents (32 x 32 FFMA per PARENT
// instead of per row).  Round 1's kernel applied W2 per row -- 1024 FFMA + 128 LDS per row, and ncu showed it bound by FP32 issue
// (tensor pipe 3 %, DRAM 9 %, 257 M instructions): 0.41 of the HBM peak.  Same value up to fp32 rounding (re-association).
//   warps 12-27 producers   a warp per PARENT, walking the CTA's parents across tile boundaries (so several tiles fill at once): the S rows
//                           of the parent by coalesced 16-byte loads (ld.global.nc, up to 10 rows in flight per lane, ids fetched one
//                           parent ahead), parked in the swizzled K-major tile layout by st.shared.  NOT tile::gather4: a gather4 moves 4 x 128 bytes per
//                           instruction and one SM's TMA unit retires one every ~63 cycles (measured here and in linear_ws_umma:
//                           1.02 M gather4 over 148 SMs = 229 us), i.e. ~2.3 TB/s chip-wide however many warps issue them -- exactly
//                           where round 1's version of this kernel sat (0.41 of the HBM peak, with the tensor pipe 3 % busy).
// a(x_i) (N x 32) is computed by the caller with the ordinary projection kernels: N rows, not N*S.
// bf16 operands only (the fp32-exact engine keeps the unfused FFMA chain); attention width H == 32; tanh.approx.f32.
#include "linear.cuh"
#include "umma_ptx.cuh"
#include <string.h>
#include <stdlib.h>
#include <float.h>
#include <algorithm>

namespace gsage {

static constexpr int AH = 32;                    // attention width (nn_modules.py:293-297 hidden_dim = 32)
static constexpr int kAtComputeWarps = 8;           // warpgroups 0-1
static constexpr int kAtIssueWarps = 4;             // warpgroup 2: warp 8 issues the MMAs, 9-11 idle (setmaxnreg is per warpgroup)
static constexpr int kAtTmaWarps = 16;              // warpgroups 3-6: producer warps (LSU row loads; the name is historical)
static constexpr int kAtThreads = 32 * (kAtComputeWarps + kAtIssueWarps + kAtTmaWarps);      // 896: launched with 72 registers per thread
template <int N> __device__ __forceinline__ void at_setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void at_setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
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
    int64_t src_rows;                                    // rows of the table behind `a` (ids outside read as zero rows)
    int kchunks, n_tiles, nbuf, buf_bytes;
    int n_mb;                                            // 128-feature blocks of the weighted sum (ceil(kchunks / 2))
    void* out; int out_bf16; int64_t ld_out;
    int* err;
};

struct AttMaps { CUtensorMap w1; CUtensorMap a; };

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// MN-major, SWIZZLE_128B operand (see wgrad_umma.cu): atoms of [8 k-rows x 128 bytes]; LBO = bytes between 64-element atoms
// along MN, SBO = between 8-row atoms along K
__device__ __forceinline__ uint64_t att_desc_mn(uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

static constexpr int kAtWN = 16;                 // parents per tile as the N of the weighted-sum MMA (PT <= 16 takes the tensor path)
static constexpr int kAtWtBytes = 2 * kAtWN * 128;   // the softmax weights as a K-major bf16 operand: 2 k-chunks x 16 rows x 128 B
static constexpr int kAtMaxMB = 6;               // 128-feature blocks: d <= 768

// CPL = 16-byte loads per lane per row in the producers (ceil(kchunks * 8 / 32))
template <int CPL>
__global__ void __launch_bounds__(kAtThreads, 1) attention_fused_kernel(const AttParams P, const __grid_constant__ AttMaps M) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [nbuf x tile (kchunks x 16 KB)] [W1: kchunks x 4 KB] [Wt: 2 x 4 KB (one per accumulator parity)] [W2 4 KB] [b1] [scores] [u] [barriers]
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* w1s = smem + (size_t)P.nbuf * P.buf_bytes;
    uint8_t* wts = w1s + (size_t)P.kchunks * kAtW1Chunk;  // (kchunks * 4 KB keeps the 1024-byte alignment)
    float* w2s = (float*)(wts + 2 * kAtWtBytes);
    float* b1s = w2s + AH * AH;
    float* sc = b1s + AH;                                 // [2][128] partial scores (one per half of the 32 attention units)
    float* us = sc + 256;                                 // [64][32] u_p = W2^T a(x_p) of the tile's parents (PT <= 64)
    float* wt = us + 64 * AH;                             // [128] softmax weights (fallback weighted sum, PT > 16)
    uint64_t* bars = (uint64_t*)(wt + 128);
    uint32_t* tmem_slot = (uint32_t*)(bars + 16);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int b) { return bar_base + 8u * b; };            // tile buffer b has landed (tx bytes)            [0..2]
    auto empty_bar = [&](int b) { return bar_base + 8u * (3 + b); };     // tile buffer b has been consumed                [3..5]
    auto dfull_bar = [&](int q) { return bar_base + 8u * (6 + q); };     // D1 (scores operand) of accumulator parity q    [6..7]
    auto wready_bar = [&](int q) { return bar_base + 8u * (8 + q); };    // softmax weights of parity q are in smem        [8..9]
    auto mfull_bar = [&](int q) { return bar_base + 8u * (10 + q); };    // D2 (weighted sums) of parity q is complete     [10..11]
    auto tfree_bar = [&](int q) { return bar_base + 8u * (12 + q); };    // accumulators + Wt of parity q are free again   [12..13]
    const uint32_t wfull_bar = bar_base + 8u * 14;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nbuf = P.nbuf;
    const bool tensor_sum = P.PT <= kAtWN && P.n_mb <= kAtMaxMB;          // the weighted sum as a second MMA
    // TMEM columns: D1 of parity q at [32 q, 32 q + 32); D2 of parity q, feature block mb at [64 + (q * kAtMaxMB + mb) * 16, + 16)

    if (threadIdx.x == 0) {
        for (int b = 0; b < 3; ++b) { mbar_init(full_bar(b), (uint32_t)P.PT); mbar_init(empty_bar(b), 32 * kAtComputeWarps); }
        for (int q = 0; q < 2; ++q) {
            mbar_init(dfull_bar(q), 1); mbar_init(wready_bar(q), 32 * kAtComputeWarps); mbar_init(mfull_bar(q), 1);
            mbar_init(tfree_bar(q), 32 * kAtComputeWarps);
        }
        mbar_init(wfull_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kAtComputeWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp < kAtComputeWarps) {
        for (int i = threadIdx.x; i < AH * AH; i += 32 * kAtComputeWarps) w2s[i] = P.w2[i];
        if (threadIdx.x < AH) b1s[threadIdx.x] = P.b1 ? P.b1[threadIdx.x] : 0.0f;
        for (int i = threadIdx.x; i < 2 * kAtWtBytes / 4; i += 32 * kAtComputeWarps) reinterpret_cast<uint32_t*>(wts)[i] = 0u;   // weights of unused rows stay 0
        // tile rows R..127 are never written by the producers: zero them once in every buffer and k-chunk (weight 0 x stale NaN = NaN)
        const int tail = 128 - P.R;
        for (int i = threadIdx.x; i < P.nbuf * P.kchunks * tail * 8; i += 32 * kAtComputeWarps) {
            const int unit = i & 7, r = P.R + (i >> 3) % tail, plane = (i >> 3) / tail;        // plane = buffer * kchunks + k-chunk
            st_shared_v4(smem_u32(smem) + (uint32_t)plane * kAtChunk + (uint32_t)r * 128u + (uint32_t)(unit << 4), make_uint4(0, 0, 0, 0));
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kAtComputeWarps) {
        // =========================== COMPUTE ===========================
        at_setmaxnreg_inc<104>();                             // 256 x 104 + 128 x 24 + 512 x 64 <= 896 x 72
        const int quarter = warp & 3, half = warp >> 2;       // TMEM lane quarter; which 16 of the 32 attention units / which feature block of a pair
        const int row = quarter * 32 + lane;                  // tile row == TMEM lane (scores); feature within a 128-block (weighted sum)
        const int S = P.S, PT = P.PT;
        const int my_parent = row / S;                        // parent of this row inside the tile (rows >= R are unused)
        const int nch = (P.d + 7) / 8;                        // 16-byte chunks per row
        // where this row's softmax weight goes in the K-major Wt operand: k-chunk row / 64, operand row my_parent, element row % 64
        const uint32_t wt_off = (uint32_t)(row >> 6) * (kAtWN * 128) + (uint32_t)my_parent * 128u +
                                (uint32_t)(((((row & 63) >> 3) ^ (my_parent & 7)) << 4) + (row & 7) * 2);
        int it = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
            const int b = it % nbuf;
            const int q = it & 1;
            const uint32_t qpar = (it >> 1) & 1;
            const int64_t parent0 = (int64_t)tile * PT;
            const bool live = row < P.R && parent0 + my_parent < P.n_parents;
            // ---- u_p = W2^T a(x_p) for the parents of this tile: u[p][k] = sum_o W2[o][k] xa[p][o]  (independent of the tile's
            //      rows: runs while they are still landing / being multiplied) ----
            for (int i = threadIdx.x; i < PT * AH; i += 32 * kAtComputeWarps) {
                const int p = i >> 5, k = i & 31;
                float acc = 0.0f;
                if (parent0 + p < P.n_parents) {
                    const float* xa = P.xa + (parent0 + p) * AH;
#pragma unroll
                    for (int o = 0; o < AH; ++o) acc = fmaf(w2s[o * AH + k], __ldg(xa + o), acc);
                }
                us[i] = acc;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(dfull_bar(q), qpar, P.err);
            tc_fence_after();
            // ---- this half's 16 attention units of the row: tanh(D1 + b1) . u_parent ----
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(q * AH + half * 16), r);
            tmem_ld_wait();
            float part = 0.0f;
            if (live) {
                const float* u = us + my_parent * AH + half * 16;
#pragma unroll
                for (int k = 0; k < 16; ++k) part = fmaf(tanh_approx(__uint_as_float(r[k]) + b1s[half * 16 + k]), u[k], part);
            }
            sc[half * 128 + row] = part;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // ---- softmax over the S rows of this row's parent (dummy rows are NOT masked: nn_modules.py:311) ----
            float w = 0.0f;
            if (half == 0 && live) {
                const int first = my_parent * S;
                float mx = -FLT_MAX;
                for (int j = 0; j < S; ++j) mx = fmaxf(mx, sc[first + j] + sc[128 + first + j]);
                float sum = 0.0f;
                for (int j = 0; j < S; ++j) sum += __expf(sc[first + j] + sc[128 + first + j] - mx);
                w = __expf(sc[row] + sc[128 + row] - mx) / sum;
            }
            if (tensor_sum) {
                // ---- m_p = sum_j w_pj n_pj as a SECOND MMA: D2^T[feature, parent] = tile^T (MN-major, straight from the swizzled tile)
                //      . Wt (K-major bf16, built here).  The softmax weights are the only thing the threads still write. ----
                if (half == 0 && row < P.R) {
                    const __nv_bfloat16 wb = __float2bfloat16_rn(w);
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(smem_u32(wts) + (uint32_t)q * kAtWtBytes + wt_off), "h"(*reinterpret_cast<const uint16_t*>(&wb)) : "memory");
                }
                fence_proxy_async();
                mbar_arrive(wready_bar(q));
                mbar_wait(mfull_bar(q), qpar, P.err);
                tc_fence_after();
                for (int mb = half; mb < P.n_mb; mb += 2) {
                    uint32_t v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(64 + (q * kAtMaxMB + mb) * kAtWN), v);
                    tmem_ld_wait();
                    const int c = mb * 128 + row;                          // feature column
                    if (c < P.d) {
#pragma unroll
                        for (int p = 0; p < kAtWN; ++p) {
                            if (p < PT && parent0 + p < P.n_parents) {
                                const int64_t at = (parent0 + p) * P.ld_out + c;
                                if (P.out_bf16) reinterpret_cast<__nv_bfloat16*>(P.out)[at] = __float2bfloat16_rn(__uint_as_float(v[p]));
                                else reinterpret_cast<float*>(P.out)[at] = __uint_as_float(v[p]);
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(tfree_bar(q));                       // accumulators and Wt of parity q may be overwritten
                mbar_arrive(empty_bar(b));                        // (the MMAs that read tile b have retired: mfull was committed after them)
            } else {
                // ---- fallback (more than 16 parents per tile, i.e. S < 8): weighted sum from the swizzled smem tile in registers ----
                if (half == 0) wt[row] = w;
                asm volatile("bar.sync 1, 256;" ::: "memory");
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
                        const float ww = wt[rr];
                        float f[8];
                        ElemTraits<__nv_bfloat16>::unpack(v, f);
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[e] = fmaf(ww, f[e], acc[e]);
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
                mbar_arrive(tfree_bar(q));
                mbar_arrive(empty_bar(b));
            }
        }
    } else if (warp < kAtComputeWarps + kAtIssueWarps) {
        // =========================== MMA ISSUER (one thread of warp 8) ===========================
        at_setmaxnreg_dec<24>();
        if (warp == kAtComputeWarps && lane == 0) {
            // scores: D1 = f32, A = B = bf16, both K-major, N = 32, M = 128
            const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AH >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            // weighted sum: D2 = f32, A = tile^T (MN-major: bit 15), B = Wt (K-major), N = 16, M = 128
            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(kAtWN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t desc_hi = umma_desc(0);
            const uint64_t desc_mn = att_desc_mn((uint32_t)kAtChunk, 1024u);        // 64-feature atoms one k-chunk (16 KB) apart, 8-row atoms 1 KB apart
            const uint32_t w16 = (smem_u32(w1s) & 0x3FFFF) >> 4, wt16 = (smem_u32(wts) & 0x3FFFF) >> 4;
            mbar_wait(wfull_bar, 0, P.err);
            auto issue_scores = [&](int it) {
                const int b = it % nbuf, q = it & 1;
                mbar_wait(tfree_bar(q), ((it >> 1) & 1) ^ 1, P.err);          // accumulators of parity q drained (first two tiles pass)
                mbar_wait(full_bar(b), (uint32_t)((it / nbuf) & 1), P.err);
                tc_fence_after();
                const uint32_t a16 = (smem_u32(smem + (size_t)b * P.buf_bytes) & 0x3FFFF) >> 4;
                const uint32_t d_tmem = tmem_base + (uint32_t)(q * AH);
                for (int kc = 0; kc < P.kchunks; ++kc) {
                    const uint64_t adesc = desc_hi | (uint64_t)(a16 + kc * (kAtChunk >> 4));
                    const uint64_t bdesc = desc_hi | (uint64_t)(w16 + kc * (kAtW1Chunk >> 4));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc1, (kc | k) ? 1u : 0u);
                }
                umma_commit(dfull_bar(q));
            };
            int n_my = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) ++n_my;
            // with one tile buffer the next tile cannot land before this tile's weighted sum has been issued: no look-ahead then
            const bool ahead = nbuf >= 2;
            if (n_my > 0 && ahead) issue_scores(0);
            for (int it = 0; it < n_my; ++it) {
                if (!ahead) issue_scores(it);
                else if (it + 1 < n_my) issue_scores(it + 1);                 // the next tile's scores go out before this tile's softmax is waited for
                const int b = it % nbuf, q = it & 1;
                if (tensor_sum) {
                    mbar_wait(wready_bar(q), (uint32_t)((it >> 1) & 1), P.err);
                    tc_fence_after();
                    const uint32_t a16 = (smem_u32(smem + (size_t)b * P.buf_bytes) & 0x3FFFF) >> 4;
                    const uint32_t b16 = wt16 + (uint32_t)q * (kAtWtBytes >> 4);
                    for (int mb = 0; mb < P.n_mb; ++mb) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(64 + (q * kAtMaxMB + mb) * kAtWN);
                        const uint64_t adesc = desc_mn | (uint64_t)(a16 + (uint32_t)(2 * mb) * (kAtChunk >> 4));
#pragma unroll
                        for (int j = 0; j < 8; ++j) {              // 16 tile rows (K) per instruction: 2 KB further down the tile, 32 B along Wt's K
                            const uint64_t bdesc = desc_hi | (uint64_t)(b16 + (uint32_t)(j >> 2) * ((kAtWN * 128) >> 4) + (uint32_t)(j & 3) * 2);
                            umma_bf16(d_tmem, adesc + 128 * j, bdesc, idesc2, j ? 1u : 0u);
                        }
                    }
                    umma_commit(mfull_bar(q));
                }
            }
        }
        __syncwarp();
    } else {
        // =========================== PRODUCERS (LSU row loads -> swizzled tile) ===========================
        at_setmaxnreg_dec<64>();
        const int gw = warp - (kAtComputeWarps + kAtIssueWarps);
        if (gw == 0 && lane == 0) {
            mbar_arrive_expect_tx(wfull_bar, (uint32_t)(P.kchunks * kAtW1Chunk));
            for (int kc = 0; kc < P.kchunks; ++kc) tma_load_2d(smem_u32(w1s) + kc * kAtW1Chunk, &M.w1, kc * 64, 0, wfull_bar);
        }
        const __nv_bfloat16* table = reinterpret_cast<const __nv_bfloat16*>(P.a);
        const int units = (P.d + 7) / 8;                       // 16-byte units that hold data (the table zero-pads the last one)
        const int units_tile = P.kchunks * 8;                  // units per tile row (a multiple of 8 >= units): the rest is written as zeros
        constexpr int RIF = CPL == 1 ? 10 : (CPL == 2 ? 5 : 3); // rows in flight per lane: RIF x CPL <= 10 sixteen-byte loads (64 registers)
        const int S = P.S, PT = P.PT;
        int n_my = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) ++n_my;
        const int n_slots = n_my * PT;                          // (tile, parent slot) pairs of this CTA, in order
        // lane j < S holds the table row of neighbour j of a parent (-1: write zeros)
        auto fetch_src = [&](int k) -> int64_t {
            if (k >= n_slots || lane >= S) return -1;
            const int64_t parent = ((int64_t)blockIdx.x + (int64_t)(k / PT) * gridDim.x) * PT + k % PT;
            if (parent >= P.n_parents) return -1;
            const int64_t at = parent * S + lane;
            const int64_t src = P.ids ? __ldg(P.ids + at) : at;
            return (src >= P.src_rows || src < 0) ? -1 : src;      // ids outside the table read as zero rows (like gather_reduce_kernel)
        };
        int k = gw;                                                // this warp's next (tile, slot) pair
        int64_t next_src = fetch_src(k);
        // every warp walks EVERY tile in order, also those it fills no slot of: its waits on a buffer's `empty` barrier then hit
        // consecutive phases (a warp that skipped tiles could run two phases ahead, where a parity wait is ambiguous)
        for (int it = 0; it < n_my; ++it) {
            const int b = it % nbuf;
            const uint32_t epar = (uint32_t)((it / nbuf) & 1) ^ 1;
            const uint32_t tile_u = smem_u32(smem + (size_t)b * P.buf_bytes);
            bool waited = false;
            while (k < (it + 1) * PT) {
                const int slot = k - it * PT;
                const int64_t my_src = next_src;
                k += kAtTmaWarps;
                next_src = fetch_src(k);                            // one parent ahead
                for (int j0 = 0; j0 < S; j0 += RIF) {
                    uint4 v[RIF][CPL];
                    uint32_t ok = 0;
#pragma unroll
                    for (int u = 0; u < RIF; ++u) {
                        const int64_t src = __shfl_sync(0xFFFFFFFFu, my_src, min(j0 + u, 31));
                        const bool live = j0 + u < S && src >= 0;
                        const __nv_bfloat16* rowp = table + (live ? src : 0) * P.lda;
#pragma unroll
                        for (int c = 0; c < CPL; ++c) {        // unconditional load from a clamped address, masked afterwards (see gather_mean_project_umma.cu)
                            const int ch = lane + 32 * c;
                            v[u][c] = ldg_nc_v4(rowp + (int64_t)min(ch, units - 1) * 8);
                            ok |= (live && ch < units) ? (1u << (u * CPL + c)) : 0u;
                        }
                    }
                    if (!waited) {                             // the loads are in flight: now make sure the buffer is free
                        if (lane == 0) mbar_wait(empty_bar(b), epar, P.err);
                        __syncwarp();
                        waited = true;
                    }
#pragma unroll
                    for (int u = 0; u < RIF; ++u) {
                        if (j0 + u < S) {
                            const int r = slot * S + j0 + u;
#pragma unroll
                            for (int c = 0; c < CPL; ++c) {
                                const int ch = lane + 32 * c;
                                if (ch < units_tile) {
                                    const uint32_t mask = 0u - ((ok >> (u * CPL + c)) & 1u);
                                    uint4 w = v[u][c];
                                    w.x &= mask; w.y &= mask; w.z &= mask; w.w &= mask;
                                    st_shared_v4(tile_u + (uint32_t)(ch >> 3) * kAtChunk + (uint32_t)r * 128u + (uint32_t)(((ch & 7) ^ (r & 7)) << 4), w);
                                }
                            }
                        }
                    }
                }
                fence_proxy_async();                           // generic-proxy stores -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(b));        // one of the PT arrivals of this tile
            }
            if (!waited) {
                if (lane == 0) mbar_wait(empty_bar(b), epar, P.err);
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kAtComputeWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// ---- host -------------------------------------------------------------------------------------------------
static bool at_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool attention_fused_eligible(const void* a, int a_dtype, int64_t lda, int d, const void* w1, int w1_dtype, int64_t ldw, int H, int S,
                              int64_t n_parents, const void* out, int64_t ld_out, int out_dtype) {
    if (getenv("GSAGE_NO_FUSED_ATTENTION")) return false;
    if (a_dtype != GSAGE_BF16 || w1_dtype != GSAGE_BF16 || H != AH || S < 2 || S > 32 || n_parents < 1 || d < 8 || d > 1024) return false;
    if (!at_aligned16(a) || !at_aligned16(w1) || !at_aligned16(out) || (lda * 2) % 16 != 0 || (ldw * 2) % 16 != 0) return false;
    if (lda < (d + 7) / 8 * 8 || ldw < (d + 7) / 8 * 8) return false;
    const int64_t es_out = out_dtype == GSAGE_BF16 ? 2 : 4;
    if ((ld_out * es_out) % 16 != 0 || ld_out < (d + 7) / 8 * 8) return false;           // whole 16-byte chunks are stored
    const int kchunks = (d + 63) / 64;
    const int fixed = 1024 + kchunks * kAtW1Chunk + 2 * kAtWtBytes + AH * AH * 4 + AH * 4 + 384 * 4 + 64 * AH * 4 + 256;
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
    U.src_rows = ids ? (table_rows > 0 ? table_rows : ((int64_t)1 << 62)) : U.n_rows;
    U.kchunks = (d + 63) / 64;
    U.n_tiles = (int)ceil_div(n_parents, U.PT);
    U.buf_bytes = U.kchunks * kAtChunk;
    const int fixed = 1024 + U.kchunks * kAtW1Chunk + 2 * kAtWtBytes + AH * AH * 4 + AH * 4 + 384 * 4 + 64 * AH * 4 + 256;
    U.nbuf = (3 * U.buf_bytes + fixed <= kAtSmemLimit) ? 3 : (2 * U.buf_bytes + fixed <= kAtSmemLimit) ? 2 : 1;
    if (const char* e = getenv("GSAGE_ATT_NBUF")) U.nbuf = std::max(1, std::min(U.nbuf, atoi(e)));
    U.n_mb = (U.kchunks + 1) / 2;
    U.out = out; U.out_bf16 = out_dtype == GSAGE_BF16; U.ld_out = ld_out;
    if (!g_att_err) {
        GS_CUDA(cudaMalloc((void**)&g_att_err, sizeof(int)));
        GS_CUDA(cudaMemset(g_att_err, 0, sizeof(int)));
    }
    U.err = g_att_err;
    AttMaps maps;
    memset(&maps, 0, sizeof(maps));
    GS_TRY(make_map(&maps.w1, w1, AH, d, ldw, AH, 2));
    const size_t smem = (size_t)U.nbuf * U.buf_bytes + fixed;
    const int grid = U.n_tiles < sm_count() ? U.n_tiles : sm_count();
    const int cpl = (U.kchunks * 8 + 31) / 32;
#define GS_ATT(C)                                                                                                            \
    do {                                                                                                                     \
        static bool attr_set = false;                                                                                        \
        if (!attr_set) {                                                                                                     \
            GS_CUDA(cudaFuncSetAttribute(attention_fused_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemLimit)); \
            attr_set = true;                                                                                                 \
        }                                                                                                                    \
        attention_fused_kernel<C><<<grid, kAtThreads, smem, s>>>(U, maps);                                                   \
    } while (0)
    if (cpl == 1) GS_ATT(1); else if (cpl == 2) GS_ATT(2); else if (cpl == 3) GS_ATT(3); else GS_ATT(4);
#undef GS_ATT
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
