// linear_umma.cu -- the dense W projection on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Replaces the same nn.Linear calls as linear_simt.cu (nn_modules.py:200,224,228,307-308,317) when every
// operand is bf16: out[r, col0_s + o] = act( sum_k A_s[row_s(r), k] * W_s[o, k] + bias_s[o] ) for up to two
// segments s (the reference's concat-with-self, [fc_x(x) | fc_neib(agg)], is two accumulators of one tile).
//
// B200 design
//   * persistent CTAs (one per SM), 128-row output tiles, warps with fixed roles:
//       warps 0-3   epilogue    TMEM -> registers (tcgen05.ld 32x32b) -> bias/activation -> bf16/fp32 -> HBM
//       warp  4     MMA issue   one elected lane issues tcgen05.mma (M=128, N=O, K=16) + tcgen05.commit
//       warp  5     TMA issue   every operand tile arrives by TMA into 128B-swizzled K-major smem:
//                                 W chunk (O x 64) and in-place A chunk (128 x 64): cp.async.bulk.tensor.2d
//                                 A rows gathered BY ID (self rows straight from the feature table):
//                                 cp.async.bulk.tensor.2d ... tile::gather4, lane l owns rows 4l..4l+3 of the tile
//                               completion = byte count on the stage's `full` mbarrier (one expect_tx arrival)
//       warps 6-13  register-path loaders, used only for a segment with reduce_S > 1 (fused gather+mean: S
//                   neighbour rows summed in fp32 registers on the way into the tile -- correct, but 256 threads
//                   cannot keep enough loads in flight, so the engine keeps the standalone gather_reduce kernel)
//   * smem ring of (A 128x64, W Ox64) bf16 chunk pairs, full/empty mbarriers, out-of-bounds rows/columns
//     zero-filled by the TMA unit (K tail 602 -> 640, last row tile);
//   * two TMEM accumulator buffers (2 x 256 fp32 columns = all 512 columns): the epilogue of tile i overlaps
//     the loads and MMAs of tile i+1.
// Every mbarrier wait is bounded (a stuck pipeline traps instead of hanging the GPU).
// Measured history of the loader (reddit layer-1 shape, 204800 x 602 -> 2 x 128, profiles/README.md):
//   cp.async + wait_group + fence.proxy.async 326 us -> LDG/STS register path 640 us -> cp.async + mbarrier
//   arrive-on 287 us (LDGSTS issue bound, ~16 B/clk/SM) -> TMA for W and in-place A 177 us -> all-TMA 197 us.
#include "linear.cuh"
#include "umma_ptx.cuh"
#include <string.h>
#include <stdlib.h>

namespace gsage {

static constexpr int UM = 128;            // rows per tile (UMMA M)
static constexpr int UK = 64;             // bf16 elements per smem chunk row = 128 bytes (one swizzle atom row)
static constexpr int kEpiWarps = 4;
static constexpr int kLoadGroups = 4, kGroupThreads = 64;           // loader groups fill different stages concurrently
static constexpr int kLoadWarps = kLoadGroups * kGroupThreads / 32;
static constexpr int kTmaWarp = kEpiWarps + 1;                       // one lane issues the TMA tile loads (W, contiguous A)
static constexpr int kFirstLoadWarp = kEpiWarps + 2;
static constexpr int kThreads = 32 * (kEpiWarps + 2 + kLoadWarps);
static constexpr int kABytes = UM * UK * 2;                   // 16 KB
static constexpr int kMaxO = 256;

struct UmmaSeg {
    const __nv_bfloat16* a; int64_t lda; const int64_t* ids;
    const __nv_bfloat16* w; int64_t ldw; int d; int O;
    const float* bias; int64_t col0;
    int kchunks;          // ceil(d / 64)
    int acc_col;          // first TMEM column of this segment's accumulator inside a buffer
    int S;                // rows reduced into one A row (1 = plain gather; > 1 = fused gather+mean)
    float scale;          // 1/S for the mean
    int kvalid;           // d rounded up to a whole 16-byte chunk: elements that may be read
};

struct UmmaParams {
    UmmaSeg seg[2];
    int n_segs; int64_t n; int act;
    void* out; int out_bf16; int64_t ld_out;
    int n_tiles; int stages; int stage_bytes; int w_bytes;
    int any_reduce;       // some segment has S > 1: loaders take the register path (fused gather+mean)
    int full_count;       // arrivals that complete a `full` barrier phase
    int tile_rows;        // rows of the input each tile covers: 128, or floor(128/S)*S with the pooled epilogue
    int pool_S;           // > 1: the epilogue reduces every S consecutive rows (one parent) and stores ONE row per parent
    int pool_max;         // 1 = max, 0 = mean
    int tf32;             // operands are fp32 read as TF32 (kind::tf32, 32 elements per 128-byte chunk row) instead of bf16
    int uk;               // elements per 128-byte chunk row: 64 (bf16) | 32 (tf32)
    int prefetch;         // L2-prefetch whole rows of the next tile from the TMA warp (GSAGE_UMMA_PREFETCH=1; off by default)
    int pf;               // tiles of LSU-path L2 prefetch distance from two otherwise idle warps (0 = off; GSAGE_PF)
    int debug;            // GSAGE_UMMA_DEBUG bit0: no A reads, bit1: no W reads, bit2: no MMA issue (timing experiments only)
    int* err;
};

// TMA descriptors (cuTensorMapEncodeTiled): per segment the W matrix, and the A matrix when it is read in place
// (no ids).  Boxes are 64 columns (128 bytes, SWIZZLE_128B) x 128 | O rows; out-of-bounds rows / columns read as
// zero, which is what pads the K tail (d = 602 -> 640) and the last row tile.
struct UmmaMaps { CUtensorMap w[2]; CUtensorMap a[2]; CUtensorMap g[2]; };   // g: 64 x 1 box for tile::gather4 (rows by id)

// 32 accumulator columns of one row: bias, activation (compile-time), convert, 16-byte stores when aligned

// ---- the kernel ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) linear_umma_kernel(const UmmaParams P, const __grid_constant__ UmmaMaps M) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages x (A 16 KB | W w_bytes)] then barriers
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + (size_t)P.stages * P.stage_bytes);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * 8 + 4);
    float* pool_scratch = (float*)(bars + 32);              // 128 x 33 floats, used only by the pooled epilogue
    volatile int* progress = (volatile int*)(tmem_slot + 1);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (16 + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (18 + b); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(full_bar(s), P.full_count); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 32 * kEpiWarps); }
        *progress = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kEpiWarps) {                                 // the MMA warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int items_per_tile = 0;
    for (int s = 0; s < P.n_segs; ++s) items_per_tile += P.seg[s].kchunks;

    if (warp < kEpiWarps) {
        // =========================== EPILOGUE ===========================
        const int row_in_tile = warp * 32 + lane;            // TMEM lane == tile row; warp w may touch lanes 32w..32w+31
        int it = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(tfull_bar(buf), (it >> 1) & 1, P.err);
            tc_fence_after();
            const int64_t row = (int64_t)tile * P.tile_rows + row_in_tile;
            for (int s = 0; s < P.n_segs; ++s) {
                const UmmaSeg& sg = P.seg[s];
                for (int c0 = 0; c0 < sg.O; c0 += 32) {
                    if (P.debug & 16) continue;                      // timing experiment: no epilogue work at all
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 256 + sg.acc_col + c0), r);
                    tmem_ld_wait();
                    if (P.pool_S > 1) {
                        // segmented reduce over the S rows of each parent (nn_modules.py:225-226,240): rows -> smem, then
                        // (parent, column) pairs; the MLP output of the neighbour rows never goes to HBM
                        const int S = P.pool_S, parents = P.tile_rows / S, valid = min(32, sg.O - c0);
                        float* mine = pool_scratch + row_in_tile * 33;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float f = __uint_as_float(r[j]);
                            if (sg.bias && j < valid) f += __ldg(sg.bias + c0 + j);
                            mine[j] = apply_act(f, P.act);
                        }
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        const int64_t parent0 = (int64_t)tile * parents, n_parents = P.n / S;
                        for (int o = threadIdx.x; o < parents * 32; o += 128) {
                            const int pp = o >> 5, c = o & 31;
                            if (parent0 + pp < n_parents && c < valid) {
                                const float* src = pool_scratch + pp * S * 33 + c;
                                float a = src[0];
                                if (P.pool_max) { for (int j = 1; j < S; ++j) a = fmaxf(a, src[j * 33]); }
                                else { for (int j = 1; j < S; ++j) a += src[j * 33]; a *= 1.0f / (float)S; }
                                const int64_t at = (parent0 + pp) * P.ld_out + sg.col0 + c0 + c;
                                if (P.out_bf16) reinterpret_cast<__nv_bfloat16*>(P.out)[at] = __float2bfloat16_rn(a);
                                else reinterpret_cast<float*>(P.out)[at] = a;
                            }
                        }
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                    } else if (row < P.n && !(P.debug & 8)) {
                        void* o = (char*)P.out + (row * P.ld_out + sg.col0 + c0) * (P.out_bf16 ? 2 : 4);
                        const float* bias = sg.bias ? sg.bias + c0 : nullptr;
                        const int valid = min(32, sg.O - c0);
                        if (P.act == GSAGE_ACT_RELU) epilogue_store32<GSAGE_ACT_RELU>(r, bias, valid, o, P.out_bf16);
                        else if (P.act == GSAGE_ACT_TANH) epilogue_store32<GSAGE_ACT_TANH>(r, bias, valid, o, P.out_bf16);
                        else epilogue_store32<GSAGE_ACT_NONE>(r, bias, valid, o, P.out_bf16);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(buf));
        }
    } else if (warp == kEpiWarps) {
        // =========================== MMA ISSUER ===========================
        int item = 0, it = 0;
        for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1, P.err);     // first use of each buffer passes immediately
            tc_fence_after();
            for (int s = 0; s < P.n_segs; ++s) {
                const UmmaSeg& sg = P.seg[s];
                // instruction descriptor: D=f32, A=B=bf16, K-major both, N = O, M = 128
                const uint32_t fmt = P.tf32 ? 2u : 1u;      // a/b format: 1 = bf16, 2 = tf32
                const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(sg.O >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256 + sg.acc_col);
                for (int kc = 0; kc < sg.kchunks; ++kc, ++item) {
                    const int stage = item % P.stages;
                    mbar_wait(full_bar(stage), (item / P.stages) & 1, P.err);
                    tc_fence_after();
                    const bool elected = elect_one();                    // (not lane == 0: umma_ptx.cuh)
                    if (elected && (P.debug & 4)) umma_commit(empty_bar(stage));
                    if (elected && !(P.debug & 4)) {
                        const uint32_t a_addr = smem_u32(smem + (size_t)stage * P.stage_bytes);
                        const uint32_t b_addr = a_addr + kABytes;
                        const uint64_t adesc = umma_desc(a_addr), bdesc = umma_desc(b_addr);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {                    // 4 x (K = 16 bf16 | 8 tf32): +32 bytes inside the swizzle atom
                            if (P.tf32) umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                            else umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                        }
                        umma_commit(empty_bar(stage));                   // smem slot reusable once these MMAs retire
                    }
                    __syncwarp();
                }
            }
            if (elect_one()) { umma_commit(tfull_bar(buf)); progress_publish(progress, it + 1); }   // accumulators of this tile complete
            __syncwarp();
        }
    } else if (warp == kTmaWarp) {
        // =========================== TMA ISSUER ===========================
        // W tile of every item, and the A tile of items whose segment is read in place: one instruction each,
        // completion (byte count) credited to the stage's `full` barrier
        if (!P.any_reduce) {
            int item = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
                for (int sidx = 0; sidx < P.n_segs; ++sidx) {
                    const UmmaSeg& sg = P.seg[sidx];
                    const uint32_t w_bytes = (uint32_t)sg.O * 128u;
                    if (P.prefetch) {                               // whole rows of the NEXT tile of this segment -> L2
                        const int64_t nbase = ((int64_t)tile + gridDim.x) * P.tile_rows + 4 * lane;
                        const uint32_t row_bytes = (uint32_t)sg.kvalid * 2u;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (nbase + q < P.n) {
                                const int64_t src = sg.ids ? __ldg(sg.ids + nbase + q) : nbase + q;
                                l2_prefetch(sg.a + src * sg.lda, row_bytes);
                            }
                        }
                    }
                    int r0 = 0, r1 = 0, r2 = 0, r3 = 0;            // lane l gathers tile rows 4l .. 4l+3 by id
                    if (sg.ids) {
                        const int64_t base = (int64_t)tile * P.tile_rows + 4 * lane;
                        if (base + 0 < P.n) r0 = (int)__ldg(sg.ids + base + 0);
                        if (base + 1 < P.n) r1 = (int)__ldg(sg.ids + base + 1);
                        if (base + 2 < P.n) r2 = (int)__ldg(sg.ids + base + 2);
                        if (base + 3 < P.n) r3 = (int)__ldg(sg.ids + base + 3);
                    }
                    for (int kc = 0; kc < sg.kchunks; ++kc, ++item) {
                        const int stage = item % P.stages;
                        mbar_wait(empty_bar(stage), ((item / P.stages) & 1) ^ 1, P.err);
                        const uint32_t sa_u = smem_u32(smem + (size_t)stage * P.stage_bytes);
                        const bool elected = elect_one();          // one lane issues everything (not lane == 0 / per-lane gather4: umma_ptx.cuh)
                        if (elected) {
                            mbar_arrive_expect_tx(full_bar(stage), w_bytes + ((sg.ids && (P.debug & 1)) ? 0u : (uint32_t)kABytes));
                            tma_load_2d(sa_u + kABytes, &M.w[sidx], kc * P.uk, 0, full_bar(stage));
                            if (!sg.ids) tma_load_2d(sa_u, &M.a[sidx], kc * P.uk, tile * P.tile_rows, full_bar(stage));
                        }
                        if (sg.ids && !(P.debug & 1)) {
                            for (int l = 0; l < 32; ++l) {             // the ids of rows 4l .. 4l+3 live in lane l: hand them to the issuing lane
                                const int a = __shfl_sync(0xFFFFFFFFu, r0, l), b = __shfl_sync(0xFFFFFFFFu, r1, l);
                                const int c = __shfl_sync(0xFFFFFFFFu, r2, l), d = __shfl_sync(0xFFFFFFFFu, r3, l);
                                if (elected) tma_gather4(sa_u + l * 512, &M.g[sidx], kc * P.uk, a, b, c, d, full_bar(stage));
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // =========================== LOADERS ===========================
        if (!P.any_reduce && P.pf > 0 && warp < kFirstLoadWarp + 2) {
            // plain mode: two of these warps pull the A rows of the tile `pf` tiles ahead into L2 (LSU path, whole rows)
            const int t = threadIdx.x - 32 * kFirstLoadWarp, nt = 64;
            const int es = P.tf32 ? 4 : 2;
            int it = 0;
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
                if (it >= P.pf) progress_wait(progress, it - P.pf);
                for (int sidx = 0; sidx < P.n_segs; ++sidx) {
                    const UmmaSeg& sg = P.seg[sidx];
                    prefetch_tile_rows(sg.a, sg.lda * es, sg.ids, (int64_t)tile * P.tile_rows, P.n, P.tile_rows, sg.kvalid * es, t, nt);
                }
            }
        }
        if (P.any_reduce) {
        // group g fills items g, g+G, g+2G, ... (an item = one (tile, segment, k-chunk) stage); the MMA warp consumes
        // items in order.  Thread (rg, c): 16-byte chunk c of rows rg, rg+8, ... of the 128-row tile.
        const int lt = (threadIdx.x - 32 * kFirstLoadWarp) % kGroupThreads;
        const int group = (threadIdx.x - 32 * kFirstLoadWarp) / kGroupThreads;
        const int rg = lt >> 3, c = lt & 7;
        const int my_tiles = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int total_items = my_tiles * items_per_tile;
        for (int item = group; item < total_items; item += kLoadGroups) {
            const int tile = blockIdx.x + (item / items_per_tile) * gridDim.x;
            int rem = item % items_per_tile, sidx = 0;
            while (rem >= P.seg[sidx].kchunks) { rem -= P.seg[sidx].kchunks; ++sidx; }
            const UmmaSeg& sg = P.seg[sidx];
            const int k0 = rem * UK + c * 8;                       // first element of this thread's chunk
            const bool live_k = k0 < sg.kvalid;
            const int stage = item % P.stages;
            mbar_wait(empty_bar(stage), ((item / P.stages) & 1) ^ 1, P.err);
            uint8_t* sa = smem + (size_t)stage * P.stage_bytes;
            const __nv_bfloat16* abase = sg.a + k0;
            const int S = sg.S;
            const uint32_t sa_u = smem_u32(sa);
            if (S == 1) {
                // ---- A, plain gather: ids first, then all 16 row slots of this thread in flight at once ----
                int64_t src[UM / 8];
#pragma unroll
                for (int i = 0; i < UM / 8; ++i) {
                    const int64_t row = (int64_t)tile * UM + rg + 8 * i;
                    src[i] = -1;
                    if (live_k && row < P.n) src[i] = sg.ids ? __ldg(sg.ids + row) : row;
                }
                uint4 v[UM / 8];
#pragma unroll
                for (int i = 0; i < UM / 8; ++i) {
                    v[i] = make_uint4(0, 0, 0, 0);
                    if (src[i] >= 0) v[i] = ldg_nc_v4(abase + src[i] * sg.lda);
                }
#pragma unroll
                for (int i = 0; i < UM / 8; ++i) {
                    const int r = rg + 8 * i;
                    st_shared_v4(sa_u + r * 128 + ((c ^ (r & 7)) << 4), v[i]);
                }
            } else {
                // ---- A, fused gather+mean.  The thread walks its 16 row slots x S neighbours as one flat sequence in
                // batches of 16: the ids of batch b+1 are fetched while the 16 row loads of batch b are in flight.
                const int total = (UM / 8) * S;
                int64_t nxt[16];
                auto fetch_ids = [&](int q0) {
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int q = q0 + u;
                        nxt[u] = -1;
                        if (q < total) {
                            const int slot = q / S, j = q - slot * S;
                            const int64_t row = (int64_t)tile * UM + rg + 8 * slot;
                            if (live_k && row < P.n) { const int64_t at = row * S + j; nxt[u] = sg.ids ? __ldg(sg.ids + at) : at; }
                        }
                    }
                };
                fetch_ids(0);
                float acc[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
                int slot = 0, j = 0;
                for (int q0 = 0; q0 < total; q0 += 16) {
                    uint4 v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        v[u] = make_uint4(0, 0, 0, 0);
                        if (nxt[u] >= 0) v[u] = ldg_nc_v4(abase + nxt[u] * sg.lda);
                    }
                    fetch_ids(q0 + 16);
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        if (q0 + u < total) {
                            float f[8];
                            ElemTraits<__nv_bfloat16>::unpack(v[u], f);
#pragma unroll
                            for (int e = 0; e < 8; ++e) acc[e] += f[e];
                            if (++j == S) {                      // slot complete: scale, round to bf16, park in the swizzled tile
#pragma unroll
                                for (int e = 0; e < 8; ++e) acc[e] *= sg.scale;
                                const int r = rg + 8 * slot;
                                st_shared_v4(sa_u + r * 128 + ((c ^ (r & 7)) << 4), ElemTraits<__nv_bfloat16>::pack(acc));
#pragma unroll
                                for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
                                j = 0; ++slot;
                            }
                        }
                    }
                }
            }
            // ---- W: O rows x this k-chunk (L2-resident after the first tile) ----
            const uint32_t sw_u = sa_u + kABytes;
            const __nv_bfloat16* wbase = sg.w + k0;
            for (int rb = rg; rb < sg.O; rb += 64) {              // 8 row slots per batch: 8 loads in flight, then 8 stores
                uint4 wv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rb + 8 * i;
                    wv[i] = make_uint4(0, 0, 0, 0);
                    if (live_k && r < sg.O) wv[i] = ldg_nc_v4(wbase + (int64_t)r * sg.ldw);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rb + 8 * i;
                    if (r < sg.O) st_shared_v4(sw_u + r * 128 + ((c ^ (r & 7)) << 4), wv[i]);
                }
            }
            fence_proxy_async();                                  // generic-proxy stores -> visible to the tensor core
            mbar_arrive(full_bar(stage));
        }
        }
    }

    // teardown: everyone done with TMEM before it is freed
    tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host -------------------------------------------------------------------------------------------------
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool linear_umma_eligible(const LinearParams& P) {
    if (P.n < 1) return false;
    int cols = 0;
    for (int i = 0; i < P.n_segs; ++i) {
        const LinearSeg& s = P.seg[i];
        if (s.a_dtype != P.seg[0].a_dtype || s.w_dtype != s.a_dtype) return false;         // all bf16, or all fp32 (run as TF32)
        if (s.a_dtype == GSAGE_F32 && s.S > 1) return false;                              // the register path is bf16 only
        if (s.O % 16 != 0 || s.O < 16 || s.O > kMaxO) return false;
        const int es = s.a_dtype == GSAGE_BF16 ? 2 : 4, per = 16 / es;
        if (!aligned16(s.a) || !aligned16(s.w) || (s.lda * es) % 16 != 0 || (s.ldw * es) % 16 != 0) return false;
        if (s.lda < (s.d + per - 1) / per * per || s.ldw < (s.d + per - 1) / per * per) return false;   // whole 16-byte chunks readable
        cols += s.O;
    }
    return cols <= 256;
}

static int* g_umma_err = nullptr;

int linear_umma_launch(const LinearParams& P, cudaStream_t s) {
    UmmaParams U;
    memset(&U, 0, sizeof(U));
    int col = 0, maxO = 0;
    U.tf32 = P.seg[0].a_dtype == GSAGE_F32 ? 1 : 0;
    U.uk = U.tf32 ? 32 : 64;
    for (int i = 0; i < P.n_segs; ++i) {
        const LinearSeg& g = P.seg[i];
        U.seg[i].a = (const __nv_bfloat16*)g.a; U.seg[i].lda = g.lda; U.seg[i].ids = g.ids;
        U.seg[i].w = (const __nv_bfloat16*)g.w; U.seg[i].ldw = g.ldw; U.seg[i].d = g.d; U.seg[i].O = g.O;
        U.seg[i].bias = g.bias; U.seg[i].col0 = g.col0;
        U.seg[i].kchunks = (g.d + U.uk - 1) / U.uk;
        U.seg[i].acc_col = col;
        U.seg[i].S = g.S > 1 ? g.S : 1;
        U.seg[i].scale = g.S > 1 ? 1.0f / (float)g.S : 1.0f;
        U.seg[i].kvalid = (g.d + 7) / 8 * 8;
        col += (g.O + 31) / 32 * 32;
        maxO = g.O > maxO ? g.O : maxO;
    }
    GS_CHECK_ARG(col <= 256, "linear_umma: accumulators need %d TMEM columns (max 256 per buffer)", col);
    U.n_segs = P.n_segs; U.n = P.n; U.act = P.act; U.out = P.out; U.out_bf16 = P.out_dtype == GSAGE_BF16; U.ld_out = P.ld_out;
    U.pool_S = P.pool_S > 1 ? P.pool_S : 1;
    U.pool_max = P.pool_max;
    GS_CHECK_ARG(U.pool_S <= UM, "linear_umma: pooled epilogue needs S <= 128");
    GS_CHECK_ARG(U.pool_S == 1 || P.n % U.pool_S == 0, "linear_umma: pooled epilogue needs n to be a multiple of S");
    U.tile_rows = U.pool_S > 1 ? (UM / U.pool_S) * U.pool_S : UM;
    U.n_tiles = (int)ceil_div(P.n, U.tile_rows);
    for (int i = 0; i < P.n_segs; ++i) U.any_reduce |= (U.seg[i].S > 1) ? 1 : 0;
    U.prefetch = 0;      // measured: 275 us with, 198 us without (reddit layer-1 shape) -- the prefetches queue behind the tile loads
    if (const char* e = getenv("GSAGE_UMMA_PREFETCH")) U.prefetch = atoi(e) != 0;
    U.pf = 1;
    if (const char* e = getenv("GSAGE_PF")) U.pf = atoi(e);
    U.full_count = U.any_reduce ? kGroupThreads : 1;                      // plain mode: the TMA lane's expect_tx arrival + byte count
    if (const char* e = getenv("GSAGE_UMMA_DEBUG")) U.debug = atoi(e);
    U.w_bytes = maxO * UK * 2;
    U.stage_bytes = (kABytes + U.w_bytes + 1023) / 1024 * 1024;
    const int budget = (P.pool_S > 1 ? 180 : 200) * 1024;
    U.stages = budget / U.stage_bytes;
    if (U.stages > 8) U.stages = 8;
    GS_CHECK_ARG(U.stages >= 3, "linear_umma: tile too large for shared memory");
    if (!g_umma_err) {
        GS_CUDA(cudaMalloc((void**)&g_umma_err, sizeof(int)));
        GS_CUDA(cudaMemset(g_umma_err, 0, sizeof(int)));
    }
    U.err = g_umma_err;
    const size_t smem = (size_t)U.stages * U.stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/ + (U.pool_S > 1 ? 128 * 33 * 4 : 0);
    static bool attr_set = false;
    if (!attr_set) {
        GS_CUDA(cudaFuncSetAttribute(linear_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int grid = U.n_tiles < sm_count() ? U.n_tiles : sm_count();
    UmmaMaps maps;
    memset(&maps, 0, sizeof(maps));
    const int es = U.tf32 ? 4 : 2;
    for (int i = 0; i < P.n_segs; ++i) {
        const LinearSeg& g = P.seg[i];
        GS_TRY(make_map(&maps.w[i], g.w, g.O, g.d, g.ldw, g.O, es));
        if (!g.ids && U.seg[i].S == 1) GS_TRY(make_map(&maps.a[i], g.a, P.n, g.d, g.lda, UM, es));
        if (g.ids && U.seg[i].S == 1) GS_TRY(make_map(&maps.g[i], g.a, g.a_rows > 0 ? g.a_rows : 0x7FFFFFFF, g.d, g.lda, 1, es));   // rows by id
    }
    linear_umma_kernel<<<grid, kThreads, smem, s>>>(U, maps);
    GS_LAUNCHED();
    return GSAGE_OK;
}

}  // namespace gsage
