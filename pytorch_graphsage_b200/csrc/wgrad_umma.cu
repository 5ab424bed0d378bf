// wgrad_umma.cu -- weight gradients of the projection on the 5th-gen tensor cores.
//
// The reference gets dW = dY^T . X from autograd through nn.Linear (/root/reference/nn_modules.py:200, backward of
// `loss.backward()` at models.py:101).  Here it is a split-K GEMM whose reduction dimension is the ROW index of both
// operands:   dW[o, k] = sum_r G[r, o] * A[row(r), k],   r over all 26*B parent rows of layer 1.
// Both operands are row-major with r as the slow index, i.e. "MN-major" for the MMA (the contiguous direction is M / N,
// not K).  tcgen05 reads such tiles directly (instruction-descriptor bits 15/16 = MN-major, SWIZZLE_128B canonical
// layout: atoms of 64 elements along MN x 8 along K, LBO = distance between atoms along MN, SBO = between atoms along K),
// so no transposed copy of the activations is ever made: TMA drops [64 rows x 64 columns] boxes of G and A straight into
// the layout the tensor core wants, and the self rows of fc_x are fetched BY ID from the feature table with
// tile::gather4, exactly like the forward projection.
//
//   unit   = (N-tile of <= 512 dW columns, K-range of rows); one CTA per unit runs every gemm of the launch over it, one
//            after the other (equal work per CTA); grid ~ one wave of 148
//   warps 0-3  epilogue   TMEM -> registers -> red.global.add.f32 into dW (fp32; dW is zeroed by the launcher)
//   warp  4    MMA issue  per 64-row stage: 4 k-steps x (N <= 256) tcgen05.mma.kind::f16, M = 128 = O
//   warps 5-12 TMA issue  per stage: 2 boxes of G + q boxes of A (in place; producer 0) or 16 x q gather4 (rows by id; eight
//                         producers x 2 row groups -- a lone warp issues one gather4 per ~50 cycles, see linear_ws_umma.cu)
// HBM-bound: every G / A byte is read once per N-tile (G is re-read by the second N-tile when d > 512).
// Requires bf16 operands, O == 128, 16-byte aligned rows.  Everything else stays on the FFMA kernel (backward.cu).
#include "backward.cuh"
#include "umma_ptx.cuh"
#include <string.h>
#include <stdlib.h>

namespace gsage {

static constexpr int GK = 64;                    // rows (reduction index) per stage
static constexpr int kBox = GK * 128;            // one [64 rows x 64 bf16] box = 8 KB
static constexpr int kWgEpiWarps = 4;
static constexpr int kWgTmaWarps = 8;             // gather4 issue is serial per warp (operands in uniform registers): spread it
static constexpr int kWgThreads = 32 * (kWgEpiWarps + 1 + kWgTmaWarps);
static constexpr int kWgMaxStages = 6;
static constexpr int kWgSmemLimit = 227 * 1024;

struct WgGemm {
    const int64_t* ids;      // NULL: A rows in place
    float* dW; int64_t lddw; int d;
};

static constexpr int kWgMaxJobs = 4;

struct WgParams {
    WgGemm gemm[kWgMaxJobs];
    int n_gemms; int n_tiles_n; int ksplit;
    int q;                   // 64-column boxes of A per N-tile (tile width = 64 * q <= 512)
    int64_t n;               // rows to reduce over
    int64_t ktiles;          // ceil(n / 64)
    int stages; int stage_bytes;
    int swap_offsets;        // debug: exchange LBO and SBO (GSAGE_WGRAD_SWAP=1)
    int* err;
};

struct WgMaps { CUtensorMap g[kWgMaxJobs]; CUtensorMap a[kWgMaxJobs]; };

// MN-major, SWIZZLE_128B operand: atoms of [8 k-rows x 128 bytes]; LBO = bytes between atoms along MN, SBO = along K
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_umma_kernel(const WgParams P, const __grid_constant__ WgMaps M) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + (size_t)P.stages * P.stage_bytes);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kWgMaxStages + 3);
    float* scratch = (float*)(bars + 2 * kWgMaxStages + 4);          // 128 x 33 floats: the epilogue's transpose buffer
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kWgMaxStages + s); };
    const uint32_t done_bar = bar_base + 8u * (2 * kWgMaxStages), drained_bar = bar_base + 8u * (2 * kWgMaxStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // unit decode: blockIdx.x = ntile * ksplit + kpart; the CTA runs EVERY gemm of the launch over its K-range, one after
    // the other (first version: one gemm per CTA -- the CTAs of the in-place gemm finished early and idled while the
    // gather4-issue-bound CTAs of the by-id gemm were still running)
    const int kpart = blockIdx.x % P.ksplit;
    const int ntile = blockIdx.x / P.ksplit;
    const int64_t kt0 = P.ktiles * kpart / P.ksplit, kt1 = P.ktiles * (kpart + 1) / P.ksplit;
    const int n_it = (int)(kt1 - kt0);
    const int col0 = ntile * P.q * 64;                       // first dW column (= A column) of this unit

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(done_bar, 1);
        mbar_init(drained_bar, 32 * kWgEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWgEpiWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kWgEpiWarps) {
        // =========================== EPILOGUE ===========================
        if (n_it > 0) {
            for (int gi = 0; gi < P.n_gemms; ++gi) {
                const WgGemm& G = P.gemm[gi];
                mbar_wait(done_bar, gi & 1, P.err);
                tc_fence_after();
                // TMEM lane == dW row, so a thread holds 32 columns of ONE row: adding them straight to dW would make every
                // warp-level RED touch 32 different 128-byte lines (12 M L2 atomic requests per launch, all CTAs on the same
                // 300 KB).  Transpose each 128 x 32 chunk through shared memory instead: a warp then adds 32 consecutive
                // floats of one row -- one line per request, 32x fewer requests.
                const int o = warp * 32 + lane;
                for (int c0 = 0; c0 < P.q * 64; c0 += 32) {
                    if (col0 + c0 >= G.d) break;
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) scratch[o * 33 + j] = __uint_as_float(r[j]);
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    const int c = col0 + c0 + lane;
                    if (c < G.d) {
#pragma unroll 8
                        for (int rr = warp; rr < 128; rr += kWgEpiWarps)
                            atomicAdd(G.dW + (int64_t)rr * G.lddw + c, scratch[rr * 33 + lane]);
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                tc_fence_before();
                mbar_arrive(drained_bar);                        // the accumulator may be overwritten by the next gemm
            }
        }
    } else if (warp == kWgEpiWarps) {
        // =========================== MMA ISSUER ===========================
        // D (128 x 64q, fp32) += G_tile^T (M = 128, MN-major) . A_tile (N, MN-major); K = 16 rows per instruction.
        // One thread, nothing recomputed per stage (the issue loop is serial latency).
        if (lane == 0 && n_it > 0) {
            const uint32_t lbo = P.swap_offsets ? 1024u : (uint32_t)kBox, sbo = P.swap_offsets ? (uint32_t)kBox : 1024u;
            const int n_first = P.q > 4 ? 4 : P.q, n_second = P.q - n_first;      // N = 64 * n_first (<= 256), then the rest
            const uint32_t base_idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc1 = base_idesc | ((uint32_t)((64 * n_first) >> 3) << 17);
            const uint32_t idesc2 = base_idesc | ((uint32_t)((64 * n_second) >> 3) << 17);
            const uint64_t desc_hi = umma_desc_mn(0, lbo, sbo);
            const uint32_t ring16 = (smem_u32(smem) & 0x3FFFF) >> 4, sb16 = (uint32_t)P.stage_bytes >> 4;
            const uint32_t a_off16 = (2 * kBox) >> 4, b2_off16 = (uint32_t)((2 + n_first) * kBox) >> 4;
            const uint32_t n_stages = (uint32_t)P.stages, d2 = tmem_base + (uint32_t)(64 * n_first);
            uint32_t stage = 0, par = 0, g16 = ring16;
            for (int gi = 0; gi < P.n_gemms; ++gi) {
                if (gi > 0) { mbar_wait(drained_bar, (gi - 1) & 1, P.err); tc_fence_after(); }
                for (int it = 0; it < n_it; ++it) {
                    mbar_wait(full_bar(stage), par, P.err);
                    tc_fence_after();
                    const uint64_t gdesc = desc_hi | (uint64_t)g16;
#pragma unroll
                    for (int j = 0; j < GK / 16; ++j) {          // 16 rows = two 8-row atoms = 2048 bytes (128 x 16 B) further down every box
                        const uint32_t acc = (it | j) ? 1u : 0u;
                        umma_bf16(tmem_base, gdesc + 128 * j, gdesc + a_off16 + 128 * j, idesc1, acc);
                        if (n_second > 0) umma_bf16(d2, gdesc + 128 * j, gdesc + b2_off16 + 128 * j, idesc2, acc);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == n_stages) { stage = 0; par ^= 1; g16 = ring16; } else g16 += sb16;
                }
                umma_commit(done_bar);
            }
        }
        __syncwarp();
    } else {
        // =========================== TMA PRODUCERS ===========================
        const int pw = warp - (kWgEpiWarps + 1);
        const bool lead = pw == 0 && lane == 0;
        const uint32_t ring_u = smem_u32(smem), sb = (uint32_t)P.stage_bytes;
        const uint32_t n_stages = (uint32_t)P.stages;
        // gather: lane -> (row group g of this producer's kGroups, box b): rows r0 + 4 (kGroups pw + g) .. + 3, columns col0 + 64 b
        constexpr int kGroups = GK / 4 / kWgTmaWarps;
        const int gq = lane / P.q, gb = lane - gq * P.q;
        const int grp = kGroups * pw + gq;
        uint32_t stage = 0, par = 1, g_addr = ring_u;
        for (int gi = 0; gi < P.n_gemms; ++gi) {
            const int64_t* ids = P.gemm[gi].ids;
            const bool gather_lane = ids != nullptr && lane < kGroups * P.q;
            for (int it = 0; it < n_it; ++it) {
                const int64_t r0 = (kt0 + it) * GK;
                int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
                if (gather_lane) {
                    const int64_t base = r0 + 4 * grp;
                    if (base + 0 < P.n) i0 = (int)__ldg(ids + base + 0);
                    if (base + 1 < P.n) i1 = (int)__ldg(ids + base + 1);
                    if (base + 2 < P.n) i2 = (int)__ldg(ids + base + 2);
                    if (base + 3 < P.n) i3 = (int)__ldg(ids + base + 3);
                }
                if (ids || lead) mbar_wait(empty_bar(stage), par, P.err);
                const uint32_t fb = full_bar(stage), a_addr = g_addr + 2 * kBox;
                if (lead) {
                    mbar_arrive_expect_tx(fb, (uint32_t)P.stage_bytes);
                    tma_load_2d(g_addr, &M.g[gi], 0, (int)r0, fb);                        // rows past n read as zero
                    tma_load_2d(g_addr + kBox, &M.g[gi], 64, (int)r0, fb);
                    if (!ids)
                        for (int b = 0; b < P.q; ++b) tma_load_2d(a_addr + b * kBox, &M.a[gi], col0 + 64 * b, (int)r0, fb);
                }
                if (gather_lane) tma_gather4(a_addr + gb * kBox + grp * 512, &M.a[gi], col0 + 64 * gb, i0, i1, i2, i3, fb);
                if (++stage == n_stages) { stage = 0; par ^= 1; g_addr = ring_u; } else g_addr += sb;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kWgEpiWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host -------------------------------------------------------------------------------------------------
static bool wg_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool wgrad_umma_eligible(const WgradJob& j) {
    if (getenv("GSAGE_NO_WGRAD_UMMA")) return false;
    if (j.g_dtype != GSAGE_BF16 || j.a_dtype != GSAGE_BF16 || j.O != 128 || j.n < 1 || j.d < 8) return false;
    if (!wg_aligned16(j.G) || !wg_aligned16(j.A) || (j.ldg * 2) % 16 != 0 || (j.lda * 2) % 16 != 0) return false;
    if (j.lda < (j.d + 7) / 8 * 8) return false;
    return j.n < (1LL << 31);
}

static int* g_wg_err = nullptr;

// all jobs of one launch share n (rows), O = 128 and d (the fc_x / fc_neib pair of one layer application, or the four
// 128-unit blocks of a pool MLP's weight gradient)
int wgrad_umma_launch(const WgradJob* jobs, int n_jobs, cudaStream_t s, bool accumulate) {
    GS_CHECK_ARG(n_jobs >= 1 && n_jobs <= kWgMaxJobs, "wgrad_umma: one to four jobs per launch");
    for (int i = 0; i < n_jobs; ++i) {
        GS_CHECK_ARG(wgrad_umma_eligible(jobs[i]), "wgrad_umma: job %d does not qualify (bf16, O == 128, aligned rows)", i);
        GS_CHECK_ARG(jobs[i].n == jobs[0].n && jobs[i].d == jobs[0].d, "wgrad_umma: jobs of one launch must share n and d");
        if (!accumulate) GS_CUDA(cudaMemsetAsync(jobs[i].dW, 0, sizeof(float) * (size_t)128 * jobs[i].lddw, s));
    }
    WgParams U;
    memset(&U, 0, sizeof(U));
    const int d = jobs[0].d;
    const int boxes = (d + 63) / 64;
    U.n_tiles_n = (boxes + 7) / 8;
    U.q = (boxes + U.n_tiles_n - 1) / U.n_tiles_n;
    U.n_gemms = n_jobs; U.n = jobs[0].n; U.ktiles = ceil_div(U.n, GK);
    const int units = U.n_tiles_n;                         // every CTA runs all jobs over its K-range
    int ksplit = sm_count() / units;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > U.ktiles) ksplit = (int)U.ktiles;
    U.ksplit = ksplit;
    U.stage_bytes = (2 + U.q) * kBox;
    const int kScratch = 128 * 33 * 4;
    U.stages = (kWgSmemLimit - 2048 - kScratch) / U.stage_bytes;
    if (U.stages > kWgMaxStages) U.stages = kWgMaxStages;
    GS_CHECK_ARG(U.stages >= 2, "wgrad_umma: stage too large");
    if (const char* e = getenv("GSAGE_WGRAD_SWAP")) U.swap_offsets = atoi(e);
    if (!g_wg_err) {
        GS_CUDA(cudaMalloc((void**)&g_wg_err, sizeof(int)));
        GS_CUDA(cudaMemset(g_wg_err, 0, sizeof(int)));
    }
    U.err = g_wg_err;
    WgMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int i = 0; i < n_jobs; ++i) {
        U.gemm[i].ids = jobs[i].ids; U.gemm[i].dW = jobs[i].dW; U.gemm[i].lddw = jobs[i].lddw; U.gemm[i].d = d;
        GS_TRY(make_map(&maps.g[i], jobs[i].G, U.n, 128, jobs[i].ldg, GK, 2));
        if (jobs[i].ids) GS_TRY(make_map(&maps.a[i], jobs[i].A, jobs[i].a_rows > 0 ? jobs[i].a_rows : 0x7FFFFFFF, d, jobs[i].lda, 1, 2));       // rows by id (tile::gather4)
        else GS_TRY(make_map(&maps.a[i], jobs[i].A, U.n, d, jobs[i].lda, GK, 2));
    }
    const size_t smem = (size_t)U.stages * U.stage_bytes + 1024 + 256 + kScratch;
    static bool attr_set = false;
    if (!attr_set) {
        GS_CUDA(cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemLimit));
        attr_set = true;
    }
    wgrad_umma_kernel<<<units * ksplit, kWgThreads, smem, s>>>(U, maps);
    GS_LAUNCHED();
    return GSAGE_OK;
}

}  // namespace gsage
