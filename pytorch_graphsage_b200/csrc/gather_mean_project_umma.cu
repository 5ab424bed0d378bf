// gather_mean_project_umma.cu -- the neighbour half of the mean aggregator in ONE kernel (GSAGE_FUSED_LAYER=0 turns it off),
//     out[p, col0 : col0 + O] = act( (1/S) sum_j table[ids[p*S + j]] . Wn^T + bias )        (nn_modules.py:197-200)
// so that the reduced rows M never travel to HBM and back (17 % of the bytes of a reddit step: DESIGN.md section 3).
//
// Why it is shaped like this (measurements behind every choice are in DESIGN.md / profiles/):
//   * the gather must keep WHOLE-ROW locality and ~80 KB of loads in flight per SM to run at HBM speed: that is what
//     gather_reduce_kernel does (0.91-0.93 of the measured peak), so the producers here are its row groups verbatim -- a warp
//     owns a parent, broadcasts its S ids by shuffle, keeps U x CPL 16-byte `ld.global.nc` in flight per lane, fp32 accumulators;
//     instead of storing the mean row to HBM it rounds it to bf16 and parks it in the swizzled K-major A tile in shared memory.
//     (The first fused attempt, linear_umma.cu's register path, gathered one 128-byte k-chunk per row per stage with 8 warps:
//     2.3 ms/step, rejected.)
//   * W must be RESIDENT: the L2 -> SM fabric is capped near HBM speed (linear_ws_umma.cu), so re-streaming W per tile would
//     cost what the fusion saves.  Wn (nk x O x 128 B = 160 KB for d = 602) plus a whole-row A tile therefore bound the tile at
//     TR = 48 parents for d = 602 (128 for d <= 256).  UMMA still runs M = 128: the descriptors walk 128 rows, rows >= TR alias
//     the next k-plane / the weights (reads only), their accumulator lanes are never stored.  The tensor pipe is ~5 % busy
//     either way.
//   * A is single-buffered.  A producer warp issues the loads of its first parent of tile i+1 BEFORE it waits for the MMAs of
//     tile i to retire (the wait sits in front of the first shared-memory store), so the HBM pipe does not drain.
//   warps 0-3   epilogue     tcgen05.ld -> bias / activation -> bf16 | fp32 -> HBM      (two TMEM buffers of O columns)
//   warp  4     MMA issue    loads W once (TMA), then per tile nk x 4 tcgen05.mma M=128, N=O, K=16   (warps 5-7: idle, they only
//                            exist so that the issuer sits in a warpgroup of its own for setmaxnreg)
//   warps 8-23  producers    gather + mean -> swizzled A tile
// First measurement (round 2, 20 producer warps x 6 loads in flight per lane, 80 registers for everybody): 4.1 TB/s, warps 39 %
// active, long-scoreboard bound -- 61 KB in flight per SM is not enough to cover HBM latency, and 48 parents over 20 warps leave
// 20 % of the producers idle per tile.  Hence: the register file is re-balanced with setmaxnreg (epilogue 56, issuer group 24,
// producers 96 registers -- the pool a warpgroup can grow from is what the OTHER warpgroups of the CTA released, so the three
// budgets must fit in the 768 x 80 registers the CTA was launched with; a first version asked for 104 and dead-locked in
// setmaxnreg.inc), 16 producers take exactly TR / 16 parents each, every lane keeps U x CPL = 10-12 loads in flight,
// the fanout is a compile-time constant so that the rounds of a parent are unrolled and the next round's loads are issued while the
// previous round is being accumulated, and the ids of the next parent are fetched one parent ahead.
// The self half  act(table[ids_self] . Wx^T)  stays on linear_ws_umma_kernel (one segment).
#include "linear.cuh"
#include "umma_ptx.cuh"
#include <string.h>
#include <stdlib.h>

namespace gsage {

static constexpr int kFgEpiWarps = 4;
static constexpr int kFgIssueWarps = 4;        // warp 4 issues; 5-7 idle (one warpgroup: setmaxnreg is per warpgroup)
static constexpr int kFgGatherWarps = 16;
static constexpr int kFgThreads = 32 * (kFgEpiWarps + kFgIssueWarps + kFgGatherWarps);     // 768: launched with 80 registers per thread
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
static constexpr int kFgSmemLimit = 227 * 1024;
static constexpr int kFgMaxCPL = 4;            // 16-byte units per lane per row: rows up to 128 units = 1024 bf16

struct FgParams {
    const __nv_bfloat16* table; int64_t ld; int64_t table_rows; int d;
    const int64_t* ids; int64_t n; int S; float scale;
    int O; const float* bias; int act;
    void* out; int out_bf16; int64_t ld_out; int64_t col0;
    int nk;               // 128-byte k-chunks per row (ceil(d / 64))
    int units_ld;         // loadable 16-byte units per table row (ld / 8, zero padded past d)
    int TR;               // parents per tile (multiple of 8, <= 128)
    int n_tiles;
    int* err;
};


// CPL = 16-byte units per lane per row (ceil(nk * 8 / 32)), U = neighbour rows in flight per lane, ST = fanout (0: run time)
template <int CPL, int U, int ST>
__global__ void __launch_bounds__(kFgThreads, 1) gather_mean_project_kernel(const FgParams P, const __grid_constant__ CUtensorMap w_map) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [A tile: nk planes of TR x 128 B] [W: nk planes of O x 128 B] [barriers]   (A first: UMMA walks 128 rows per plane,
    // the rows past TR of the last plane must still be addressable shared memory -- they land in the weights)
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t a_plane = (uint32_t)P.TR * 128u, w_plane = (uint32_t)P.O * 128u;
    uint8_t* a_tile = smem;
    uint8_t* w_area = smem + (size_t)P.nk * a_plane;
    uint64_t* bars = (uint64_t*)(w_area + (size_t)P.nk * w_plane);
    uint32_t* tmem_slot = (uint32_t*)(bars + 8);
    const uint32_t bar_base = smem_u32(bars);
    const uint32_t w_full = bar_base, a_full = bar_base + 8, a_free = bar_base + 16;
    auto t_full = [&](int b) { return bar_base + 8u * (3 + b); };
    auto t_empty = [&](int b) { return bar_base + 8u * (5 + b); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(w_full, 1);
        mbar_init(a_full, kFgGatherWarps);
        mbar_init(a_free, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(t_full(b), 1); mbar_init(t_empty(b), 32 * kFgEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kFgEpiWarps) {                               // the MMA warp owns the TMEM allocation: two buffers of 128 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kFgEpiWarps) {
        // =========================== EPILOGUE ===========================
        setmaxnreg_dec<56>();
        const int row_in_tile = warp * 32 + lane;            // TMEM lane == tile row; only rows < TR carry a parent
        int it = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(t_full(buf), (it >> 1) & 1, P.err);
            tc_fence_after();
            const int64_t row = (int64_t)tile * P.TR + row_in_tile;
            const bool live = row_in_tile < P.TR && row < P.n;
            if (warp * 32 < P.TR) {                          // warps whose 32 lanes are all past TR have nothing to read
                for (int c0 = 0; c0 < P.O; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 128 + c0), r);
                    tmem_ld_wait();
                    if (live) {
                        void* o = (char*)P.out + (row * P.ld_out + P.col0 + c0) * (P.out_bf16 ? 2 : 4);
                        const float* bias = P.bias ? P.bias + c0 : nullptr;
                        const int valid = min(32, P.O - c0);
                        if (P.act == GSAGE_ACT_RELU) epilogue_store32<GSAGE_ACT_RELU>(r, bias, valid, o, P.out_bf16);
                        else if (P.act == GSAGE_ACT_TANH) epilogue_store32<GSAGE_ACT_TANH>(r, bias, valid, o, P.out_bf16);
                        else epilogue_store32<GSAGE_ACT_NONE>(r, bias, valid, o, P.out_bf16);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(t_empty(buf));
        }
    } else if (warp < kFgEpiWarps + kFgIssueWarps) {
        // =========================== W LOAD + MMA ISSUER (one thread of warp 4) ===========================
        setmaxnreg_dec<24>();
        if (warp == kFgEpiWarps && elect_one()) {          // (elect_one, not lane == 0: umma_ptx.cuh)
            mbar_arrive_expect_tx(w_full, (uint32_t)P.nk * w_plane);
            for (int kc = 0; kc < P.nk; ++kc) tma_load_2d(smem_u32(w_area) + (uint32_t)kc * w_plane, &w_map, kc * 64, 0, w_full);
            mbar_wait(w_full, 0, P.err);
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N = O, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.O >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t desc_hi = umma_desc(0);
            const uint32_t a16 = (smem_u32(a_tile) & 0x3FFFF) >> 4, w16 = (smem_u32(w_area) & 0x3FFFF) >> 4;
            const uint32_t a_plane16 = a_plane >> 4, w_plane16 = w_plane >> 4;
            int it = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
                const uint32_t buf = it & 1;
                mbar_wait(t_empty(buf), ((it >> 1) & 1) ^ 1, P.err);       // first use of each buffer passes immediately
                mbar_wait(a_full, it & 1, P.err);                          // every producer warp has parked its parents of this tile
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 128u;
                uint64_t adesc = desc_hi | (uint64_t)a16, bdesc = desc_hi | (uint64_t)w16;
                for (int kc = 0; kc < P.nk; ++kc) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                    adesc += a_plane16; bdesc += w_plane16;
                }
                umma_commit(a_free);                                       // the A tile may be overwritten once these MMAs retire
                umma_commit(t_full(buf));                                  // ... and the accumulators are complete
            }
        }
        __syncwarp();
    } else {
        // =========================== PRODUCERS: gather + mean -> swizzled A tile ===========================
        setmaxnreg_inc<96>();                                               // 128 x 56 + 128 x 24 + 512 x 96 <= 768 x 80
        constexpr int VEC = 8;                                             // bf16 per 16-byte unit
        const int gw = warp - (kFgEpiWarps + kFgIssueWarps);
        const uint32_t a_u = smem_u32(a_tile);
        const int units_tile = P.nk * 8;                                   // 16-byte units per A row (a multiple of 8 >= units_ld)
        const int S = ST > 0 ? ST : P.S;
        auto fetch_id = [&](int64_t parent) -> int64_t {                   // lane j holds the id of neighbour j of `parent`
            if (parent >= P.n || lane >= S) return -1;
            const int64_t at = parent * (int64_t)S + lane;
            return P.ids ? __ldg(P.ids + at) : at;
        };
        int it = 0;
        int64_t next_id = fetch_id((int64_t)blockIdx.x * P.TR + gw);
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
            bool may_store = (it == 0);                                    // tile 0: the A tile is free from the start
            for (int r = gw; r < P.TR; r += kFgGatherWarps) {
                const int64_t my_id = next_id;
                {   // one parent ahead: the ids of this warp's next parent (next row of this tile, or its first row of the next tile)
                    const bool wrap = r + kFgGatherWarps >= P.TR;
                    const int64_t ntile = wrap ? (int64_t)tile + gridDim.x : (int64_t)tile;
                    next_id = ntile < P.n_tiles ? fetch_id(ntile * P.TR + (wrap ? gw : r + kFgGatherWarps)) : -1;
                }
                float acc[CPL][VEC];
#pragma unroll
                for (int c = 0; c < CPL; ++c)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[c][e] = 0.0f;
                {
#pragma unroll
                    for (int jj = 0; jj < (ST > 0 ? ST : 32); jj += U) {
                        if (ST == 0 && jj >= S) break;
                        // Every load is issued UNCONDITIONALLY from a clamped (always valid) address and masked afterwards.  The
                        // obvious `ok ? ld : 0` compiles to a predicated load followed by a predicated MOV of zero into the same
                        // register -- and a predicated-off MOV still waits on the scoreboard of the load in flight, which
                        // serialised the loads of a round (ncu source view of v2: ~12 equally hot wait points per parent).
                        uint4 v[U][CPL];
                        uint32_t okbits = 0;
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int64_t id = __shfl_sync(0xFFFFFFFFu, my_id, min(jj + u, 31));
                            const bool live = (jj + u < S) && ((uint64_t)id < (uint64_t)P.table_rows);
                            const __nv_bfloat16* row = P.table + (live ? id : 0) * P.ld;
#pragma unroll
                            for (int c = 0; c < CPL; ++c) {
                                const int ch = lane + 32 * c;
                                if (ST > 0 && jj + u >= ST) { v[u][c] = make_uint4(0, 0, 0, 0); continue; }     // compile-time dead rows: no load at all
                                v[u][c] = ldg_nc_v4(row + (int64_t)min(ch, P.units_ld - 1) * VEC);
                                okbits |= (live && ch < P.units_ld) ? (1u << (u * CPL + c)) : 0u;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int c = 0; c < CPL; ++c) {
                                if (ST > 0 && jj + u >= ST) continue;
                                const uint32_t mask = 0u - ((okbits >> (u * CPL + c)) & 1u);
                                uint4 w = v[u][c];
                                w.x &= mask; w.y &= mask; w.z &= mask; w.w &= mask;
                                float f[VEC];
                                ElemTraits<__nv_bfloat16>::unpack(w, f);
#pragma unroll
                                for (int e = 0; e < VEC; ++e) acc[c][e] = fmaf(1.0f, f[e], acc[c][e]);      // same arithmetic as gather_reduce_kernel
                            }
                    }
                }
                if (!may_store) {                                          // loads of this parent are issued: now wait for tile it-1's MMAs
                    if (lane == 0) mbar_wait(a_free, (it - 1) & 1, P.err);
                    __syncwarp();
                    may_store = true;
                }
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int ch = lane + 32 * c;
                    if (ch < units_tile) {                                 // units past the row's data are written as zeros (W is zero there too)
                        float f[VEC];
#pragma unroll
                        for (int e = 0; e < VEC; ++e) f[e] = acc[c][e] * P.scale;
                        const uint32_t dst = a_u + (uint32_t)(ch >> 3) * a_plane + (uint32_t)r * 128u + (uint32_t)(((ch & 7) ^ (r & 7)) << 4);
                        st_shared_v4(dst, ElemTraits<__nv_bfloat16>::pack(f));
                    }
                }
            }
            if (!may_store) {                                              // a warp without a parent in this tile still keeps the parity in step
                if (lane == 0) mbar_wait(a_free, (it - 1) & 1, P.err);
                __syncwarp();
            }
            fence_proxy_async();                                           // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
    }

    // teardown: everyone done with TMEM before it is freed
    tc_fence_before();
    __syncthreads();
    if (warp == kFgEpiWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// ---- host -------------------------------------------------------------------------------------------------
static int fg_tile_rows(int nk, int O) {
    const int fixed = 1024 /*align slack*/ + 256 /*barriers*/;
    const int left = kFgSmemLimit - fixed - nk * O * 128;
    if (left <= 0) return 0;
    int tr = left / (nk * 128) / 8 * 8;
    if (tr >= kFgGatherWarps) tr = tr / kFgGatherWarps * kFgGatherWarps;     // every producer warp takes the same number of parents
    return tr > 128 ? 128 : tr;
}

bool gather_mean_project_eligible(const void* table, int dtype, int64_t ld, int d, int S, const void* w, int w_dtype, int64_t ldw, int O) {
    // opt-in (GSAGE_FUSED_LAYER=1): measured slower than gather_reduce_kernel + the projection it replaces
    { const char* f = getenv("GSAGE_FUSED_LAYER"); if (!f || atoi(f) == 0) return false; }
    if (dtype != GSAGE_BF16 || w_dtype != GSAGE_BF16 || S < 1 || S > 32 || d < 1) return false;
    if (O % 16 != 0 || O < 16 || O > 128) return false;                     // one TMEM buffer = 128 columns
    if ((reinterpret_cast<uintptr_t>(table) & 15u) || (reinterpret_cast<uintptr_t>(w) & 15u) || ld % 8 != 0 || ldw % 8 != 0) return false;
    if (ld < (d + 7) / 8 * 8 || ldw < (d + 7) / 8 * 8) return false;        // whole 16-byte units readable (zero padded)
    const int nk = (d + 63) / 64;
    if (nk * 8 > 32 * kFgMaxCPL) return false;
    return fg_tile_rows(nk, O) >= 16;
}

static int* g_fg_err = nullptr;

int gather_mean_project_launch(const void* table, int64_t ld, int64_t table_rows, int d, const int64_t* ids, int64_t n, int S,
                               const void* w, int64_t ldw, int O, const float* bias, int act, void* out, int out_dtype, int64_t ld_out,
                               int64_t col0, cudaStream_t s) {
    if (n <= 0) return GSAGE_OK;
    FgParams P;
    memset(&P, 0, sizeof(P));
    P.table = (const __nv_bfloat16*)table; P.ld = ld; P.table_rows = table_rows; P.d = d;
    P.ids = ids; P.n = n; P.S = S; P.scale = 1.0f / (float)S;
    P.O = O; P.bias = bias; P.act = act;
    P.out = out; P.out_bf16 = out_dtype == GSAGE_BF16; P.ld_out = ld_out; P.col0 = col0;
    P.nk = (d + 63) / 64;
    const int units_row = (d + 7) / 8;                                      // units that hold data (the last one zero padded by the table)
    P.units_ld = units_row;
    P.TR = fg_tile_rows(P.nk, O);
    GS_CHECK_ARG(P.TR >= 16, "gather_mean_project: operands do not fit in shared memory");
    P.n_tiles = (int)ceil_div(n, P.TR);
    if (!g_fg_err) {
        GS_CUDA(cudaMalloc((void**)&g_fg_err, sizeof(int)));
        GS_CUDA(cudaMemset(g_fg_err, 0, sizeof(int)));
    }
    P.err = g_fg_err;
    CUtensorMap w_map;
    GS_TRY(make_map(&w_map, w, O, d, ldw, O, 2));
    const size_t smem = (size_t)P.nk * P.TR * 128 + (size_t)P.nk * O * 128 + 1024 + 256;
    GS_CHECK_ARG(smem <= (size_t)kFgSmemLimit, "gather_mean_project: %zu bytes of shared memory needed", smem);
    const int grid = P.n_tiles < sm_count() ? P.n_tiles : sm_count();
    const int cpl = (P.nk * 8 + 31) / 32;
#define GS_FG(C, UU, SS)                                                                                                    \
    do {                                                                                                                    \
        static bool attr_set = false;                                                                                       \
        if (!attr_set) {                                                                                                    \
            GS_CUDA(cudaFuncSetAttribute(gather_mean_project_kernel<C, UU, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFgSmemLimit)); \
            attr_set = true;                                                                                                \
        }                                                                                                                   \
        gather_mean_project_kernel<C, UU, SS><<<grid, kFgThreads, smem, s>>>(P, w_map);                                     \
    } while (0)
    // U x CPL 16-byte loads in flight per lane: 10-12 (40-48 of the producers' 96 registers)
#define GS_FG_S(C, UU)                                                                                                      \
    do {                                                                                                                    \
        if (S == 10) GS_FG(C, UU, 10); else if (S == 25) GS_FG(C, UU, 25); else GS_FG(C, UU, 0);                            \
    } while (0)
    if (cpl == 1) GS_FG_S(1, 10);             // rows <= 512 B: a whole fanout-10 parent in one round
    else if (cpl == 2) GS_FG_S(2, 5);
    else if (cpl == 3) GS_FG_S(3, 4);
    else GS_FG_S(4, 3);
#undef GS_FG_S
#undef GS_FG
    GS_LAUNCHED();
    return GSAGE_OK;
}

}  // namespace gsage

using namespace gsage;

extern "C" int gsage_gather_mean_project(const void* table_dev, int dtype, int64_t ld, int64_t n_table_rows, int d, const int64_t* ids_dev,
                                         int64_t n_parents, int S, const void* w_dev, int w_dtype, int64_t ldw, int O, const float* bias_dev,
                                         int act, void* out_dev, int out_dtype, int64_t ld_out, int64_t col0, void* stream) {
    GS_CHECK_ARG(table_dev && w_dev && out_dev && n_parents >= 0 && col0 >= 0 && col0 + O <= ld_out, "gather_mean_project: bad arguments");
    GS_CHECK_ARG(out_dtype == GSAGE_F32 || out_dtype == GSAGE_BF16, "gather_mean_project: bad out dtype");
    GS_CHECK_ARG(gather_mean_project_eligible(table_dev, dtype, ld, d, S, w_dev, w_dtype, ldw, O),
                 "gather_mean_project: needs a bf16 table and W with 16-byte aligned zero-padded "
                 "rows, S <= 32, O %% 16 == 0 and <= 128, and weights that fit in shared memory next to a tile of >= 16 parents");
    return gather_mean_project_launch(table_dev, ld, n_table_rows, d, ids_dev, n_parents, S, w_dev, ldw, O, bias_dev, act, out_dev, out_dtype,
                                      ld_out, col0, as_stream(stream));
}
