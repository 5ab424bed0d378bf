// linear_pool_umma.cu -- the pool aggregators' per-neighbour MLP with the pool fused in, on tcgen05 ("swap-AB").
//
// Replaces  h = relu(mlp(neibs)); h.view(N, S, H).max(dim=1)[0] | .mean(dim=1)   (nn_modules.py:224-226,240,252)
//
//   out[p, col0 + h] = reduce_{j < S} relu( A[row(p*S + j)] . W[h] + bias[h] )
//
// The obvious orientation (rows of A on the M side, as in linear_umma.cu) leaves the S rows of a parent spread over
// TMEM *lanes*, so the pool needs a shared-memory exchange and two barriers per 32 columns (measured 2.3x slower
// than not fusing at all).  Here the operands are swapped: the M side holds 128 hidden units (rows of W), the N
// side holds up to 256 neighbour rows, so the accumulator is D[hidden unit, neighbour row] -- one hidden unit per
// TMEM lane / epilogue thread, and the S rows of a parent are S consecutive COLUMNS: the pool is a running
// max / sum in the thread's registers, and a warp's store of one parent is 32 consecutive hidden units (coalesced).
// The (N*S, H) hidden matrix never exists in HBM.
//
// Tiles: one row block of R = floor(128/S)*S neighbour rows against up to FOUR blocks of 128 hidden units at once
// (4 x 128 fp32 accumulator columns = the whole TMEM), so every neighbour row is gathered exactly once per pass
// (a first version with one hidden block per tile re-gathered every row block 4x and was TMA-request bound:
// 5.5 ms/step on pokec max-pool; unfused 2.76 ms).  Sixteen epilogue warps (four per TMEM lane quarter, one hidden
// block each).  Warp roles, smem ring and all-TMA operand loads (tile::gather4 for rows by id) as in linear_umma.cu.
#include "linear.cuh"
#include "umma_ptx.cuh"
#include <string.h>

namespace gsage {

static constexpr int PM = 128;            // hidden units per block (UMMA M)
static constexpr int kHB = 4;             // hidden blocks per pass: 4 x 128 accumulator columns = all of TMEM
static constexpr int kPoolEpiWarps = 16;          // four warps per TMEM lane quarter: one hidden block each
static constexpr int kPoolThreads = 32 * (kPoolEpiWarps + 2);
static constexpr int kMBytes = PM * 128;                      // 16 KB: 128 W rows x one 128-byte chunk
static constexpr int kNBytes = 128 * 128;                     // 16 KB: up to 128 neighbour rows x one chunk

struct PoolParams {
    const void* a; int64_t lda; const int64_t* ids;            // neighbour rows (gathered by id, or in place)
    const float* bias; int64_t col0;
    int d, H, S, pool_max, act;
    int64_t n_rows, n_parents;                                  // n_rows = n_parents * S
    int R, Nmma, row_blocks, h_blocks, passes, kchunks, uk, tf32;
    void* out; int out_bf16; int64_t ld_out;
    int stages, stage_bytes; int* err;
};

struct PoolMaps { CUtensorMap w; CUtensorMap a; CUtensorMap g; };

template <int ACT, bool POOL_MAX>
__global__ void __launch_bounds__(kPoolThreads, 1) linear_pool_umma_kernel(const PoolParams P, const __grid_constant__ PoolMaps M) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + (size_t)P.stages * P.stage_bytes);
    uint32_t* tmem_slot = (uint32_t*)(bars + 20);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
    const uint32_t tfull_bar = bar_base + 8u * 16, tempty_bar = bar_base + 8u * 17;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tfull_bar, 1); mbar_init(tempty_bar, 32 * kPoolEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kPoolEpiWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles = P.row_blocks * P.passes;

    if (warp < kPoolEpiWarps) {
        // ============ EPILOGUE: one hidden unit per thread, pool along the columns (warps w and w+4 share a lane quarter) ============
        int it = 0;
        const int parents_per_block = P.R / P.S;
        const int quarter = warp & 3, first = warp >> 2;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int rb = tile / P.passes, pass = tile - rb * P.passes;
            mbar_wait(tfull_bar, it & 1, P.err);
            tc_fence_after();
            const int64_t row0 = (int64_t)rb * P.R;
            const int cols = (int)min((int64_t)P.R, P.n_rows - row0);          // valid neighbour rows of this block
            for (int j = first; j < kHB; j += kPoolEpiWarps / 4) {              // this warp's hidden block(s) of the pass
                const int hb = pass * kHB + j;
                if (hb >= P.h_blocks) break;
                const int h = hb * PM + quarter * 32 + lane;
                const float bias = (P.bias && h < P.H) ? __ldg(P.bias + h) : 0.0f;
                float acc = POOL_MAX ? -3.0e38f : 0.0f;
                int cnt = 0;
                int64_t parent = (int64_t)rb * parents_per_block;
                for (int c0 = 0; c0 < cols; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j * 128 + c0), r);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        if (c0 + q < cols) {
                            float v = __uint_as_float(r[q]) + bias;
                            if (ACT == GSAGE_ACT_RELU) v = fmaxf(v, 0.0f);
                            if (ACT == GSAGE_ACT_TANH) v = tanhf(v);
                            acc = POOL_MAX ? fmaxf(acc, v) : acc + v;
                            if (++cnt == P.S) {
                                if (h < P.H) {
                                    const float o = POOL_MAX ? acc : acc * (1.0f / (float)P.S);
                                    const int64_t at = parent * P.ld_out + P.col0 + h;
                                    if (P.out_bf16) reinterpret_cast<__nv_bfloat16*>(P.out)[at] = __float2bfloat16_rn(o);
                                    else reinterpret_cast<float*>(P.out)[at] = o;
                                }
                                ++parent; cnt = 0;
                                acc = POOL_MAX ? -3.0e38f : 0.0f;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar);
        }
    } else if (warp == kPoolEpiWarps) {
        // ============ MMA ISSUER: D_j[hidden, row] += W_block_j . rows^T for the (up to) four hidden blocks of the pass ============
        int item = 0, it = 0;
        const uint32_t fmt = P.tf32 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(P.Nmma >> 3) << 17) | ((uint32_t)(PM >> 4) << 24);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int pass = tile % P.passes;
            const int nb = min(kHB, P.h_blocks - pass * kHB);
            mbar_wait(tempty_bar, (it & 1) ^ 1, P.err);                 // the epilogue has drained the previous tile
            tc_fence_after();
            for (int kc = 0; kc < P.kchunks; ++kc, ++item) {
                const int stage = item % P.stages;
                mbar_wait(full_bar(stage), (item / P.stages) & 1, P.err);
                tc_fence_after();
                if (elect_one()) {                             // (not lane == 0: umma_ptx.cuh)
                    const uint32_t base = smem_u32(smem + (size_t)stage * P.stage_bytes);
                    const uint64_t bdesc = umma_desc(base + kHB * kMBytes);
                    for (int j = 0; j < nb; ++j) {
                        const uint64_t adesc = umma_desc(base + j * kMBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (P.tf32) umma_tf32(tmem_base + j * 128, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                            else umma_bf16(tmem_base + j * 128, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                        }
                    }
                    umma_commit(empty_bar(stage));
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(tfull_bar);
            __syncwarp();
        }
    } else {
        // =========================== TMA ISSUER ===========================
        int item = 0;
        const int groups = (P.R + 3) / 4;                          // 4-row gather groups (<= 32: one per lane)
        const uint32_t n_bytes = P.ids ? (uint32_t)groups * 512u : (uint32_t)kNBytes;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int rb = tile / P.passes, pass = tile - rb * P.passes;
            const int nb = min(kHB, P.h_blocks - pass * kHB);
            const int64_t row0 = (int64_t)rb * P.R;
            int rid[4] = {0, 0, 0, 0};
            if (P.ids && lane < groups) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t r = row0 + 4 * lane + i;
                    if (r < P.n_rows && 4 * lane + i < P.R) rid[i] = (int)__ldg(P.ids + r);
                }
            }
            for (int kc = 0; kc < P.kchunks; ++kc, ++item) {
                const int stage = item % P.stages;
                mbar_wait(empty_bar(stage), ((item / P.stages) & 1) ^ 1, P.err);
                const uint32_t base = smem_u32(smem + (size_t)stage * P.stage_bytes);
                if (elect_one()) {
                    mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(nb * kMBytes) + n_bytes);
                    for (int j = 0; j < nb; ++j) tma_load_2d(base + j * kMBytes, &M.w, kc * P.uk, (pass * kHB + j) * PM, full_bar(stage));
                    if (!P.ids) tma_load_2d(base + kHB * kMBytes, &M.a, kc * P.uk, (int)row0, full_bar(stage));
                }
                if (P.ids) {
                    const bool elected = elect_one();
                    for (int l = 0; l < groups; ++l) {                 // the ids of rows 4l .. 4l+3 live in lane l: hand them to the issuing lane
                        const int a = __shfl_sync(0xFFFFFFFFu, rid[0], l), b = __shfl_sync(0xFFFFFFFFu, rid[1], l);
                        const int c = __shfl_sync(0xFFFFFFFFu, rid[2], l), d = __shfl_sync(0xFFFFFFFFu, rid[3], l);
                        if (elected) tma_gather4(base + kHB * kMBytes + l * 512, &M.g, kc * P.uk, a, b, c, d, full_bar(stage));
                    }
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kPoolEpiWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

static bool pool_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool linear_pool_umma_eligible(const LinearParams& P) {
    if (P.n_segs != 1 || P.pool_S < 2 || P.pool_S > 128 || P.n < 1 || P.n % P.pool_S != 0) return false;
    const LinearSeg& s = P.seg[0];
    if (s.S > 1 || s.w_trans || s.w_dtype != s.a_dtype) return false;
    if (s.a_dtype != GSAGE_BF16 && s.a_dtype != GSAGE_F32) return false;
    const int es = s.a_dtype == GSAGE_BF16 ? 2 : 4, per = 16 / es;
    if (!pool_aligned16(s.a) || !pool_aligned16(s.w) || (s.lda * es) % 16 != 0 || (s.ldw * es) % 16 != 0) return false;
    if (s.lda < (s.d + per - 1) / per * per || s.ldw < (s.d + per - 1) / per * per) return false;
    return true;
}

static int* g_pool_err = nullptr;

int linear_pool_umma_launch(const LinearParams& P, cudaStream_t s) {
    const LinearSeg& g = P.seg[0];
    PoolParams U;
    memset(&U, 0, sizeof(U));
    U.a = g.a; U.lda = g.lda; U.ids = g.ids; U.bias = g.bias; U.col0 = g.col0;
    U.d = g.d; U.H = g.O; U.S = P.pool_S; U.pool_max = P.pool_max; U.act = P.act;
    U.n_rows = P.n; U.n_parents = P.n / P.pool_S;
    U.tf32 = g.a_dtype == GSAGE_F32 ? 1 : 0;
    U.uk = U.tf32 ? 32 : 64;
    U.R = (128 / U.S) * U.S;
    U.Nmma = (U.R + 15) / 16 * 16;
    U.row_blocks = (int)ceil_div(U.n_rows, U.R);
    U.h_blocks = (U.H + PM - 1) / PM;
    U.passes = (U.h_blocks + kHB - 1) / kHB;
    U.stage_bytes = kHB * kMBytes + kNBytes;
    U.kchunks = (U.d + U.uk - 1) / U.uk;
    U.out = P.out; U.out_bf16 = P.out_dtype == GSAGE_BF16; U.ld_out = P.ld_out;
    U.stages = 2;                                             // 2 x 80 KB
    if (!g_pool_err) {
        GS_CUDA(cudaMalloc((void**)&g_pool_err, sizeof(int)));
        GS_CUDA(cudaMemset(g_pool_err, 0, sizeof(int)));
    }
    U.err = g_pool_err;
    const int es = U.tf32 ? 4 : 2;
    PoolMaps maps;
    memset(&maps, 0, sizeof(maps));
    GS_TRY(make_map(&maps.w, g.w, g.O, g.d, g.ldw, PM, es));
    if (g.ids) GS_TRY(make_map(&maps.g, g.a, g.a_rows > 0 ? g.a_rows : 0x7FFFFFFF, g.d, g.lda, 1, es));
    else GS_TRY(make_map(&maps.a, g.a, P.n, g.d, g.lda, 128, es));
    const size_t smem = (size_t)U.stages * U.stage_bytes + 1024 + 256;
    const int tiles = U.row_blocks * U.passes;
    const int grid = tiles < sm_count() ? tiles : sm_count();
#define GS_POOL_LAUNCH(A, MX)                                                                                                   \
    do {                                                                                                                        \
        static bool attr_set = false;                                                                                           \
        if (!attr_set) {                                                                                                        \
            GS_CUDA(cudaFuncSetAttribute(linear_pool_umma_kernel<A, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
            attr_set = true;                                                                                                    \
        }                                                                                                                       \
        linear_pool_umma_kernel<A, MX><<<grid, kPoolThreads, smem, s>>>(U, maps);                                               \
    } while (0)
    if (U.act == GSAGE_ACT_RELU) { if (U.pool_max) GS_POOL_LAUNCH(GSAGE_ACT_RELU, true); else GS_POOL_LAUNCH(GSAGE_ACT_RELU, false); }
    else if (U.act == GSAGE_ACT_TANH) { if (U.pool_max) GS_POOL_LAUNCH(GSAGE_ACT_TANH, true); else GS_POOL_LAUNCH(GSAGE_ACT_TANH, false); }
    else { if (U.pool_max) GS_POOL_LAUNCH(GSAGE_ACT_NONE, true); else GS_POOL_LAUNCH(GSAGE_ACT_NONE, false); }
#undef GS_POOL_LAUNCH
    GS_LAUNCHED();
    return GSAGE_OK;
}

}  // namespace gsage
