// linear_ws_umma.cu -- the dense W projection on tcgen05 with the WEIGHTS STATIONARY in shared memory.
//
// Same contract as linear_umma.cu (nn_modules.py:200,228,317: out[r, col0_s + o] = act(A_s[row_s(r)] . W_s[o] + bias_s[o])
// for up to two segments s, the reference's concat-with-self), different data movement.  The streaming kernel
// re-reads the W chunk of every (tile, k-chunk) stage from L2, so for the layer-1 shape (d = 602, O = 128) half of
// the bytes an SM ingests are weights -- and the L2 -> SM fabric (~6.5 TB/s chip-wide, B300_MICROARCH.md "LTS
// throughput cap") is what bounds that kernel, not HBM.  Here a CTA loads (most of) the W of a *phase* once and keeps it
// for all of its row tiles: the first `kres` k-chunks of W are resident (kres x O x 128 B), the rest -- whatever does not
// fit next to a ring deep enough to cover the load latency -- streams through the ring with the A chunks.
// Measured (reddit layer 1, B=16384): fully resident W (160 KB) leaves 4 x 16 KB stages = too few bytes in flight, 418 us
// vs 391 us streaming; kres and the ring depth are therefore planned together (GSAGE_WS_STAGES, default 6).
//   phase = the segments whose weights fit in shared memory together (both for d <= 256; one at a time for
//           d = 602: all row tiles with Wx, then all row tiles again with Wn -- each phase writes its own column
//           range of the output, the activation is elementwise, so the result is identical)
//   warps 0-3  epilogue   tcgen05.ld -> bias / activation -> bf16 | fp32 -> HBM
//   warp  4    MMA issue  tcgen05.mma M=128, N=O, K=16 (bf16) | 8 (tf32); A from the stage ring, B from the resident W
//   warps 5-12 TMA issue  W of the phase (one expect_tx for all of it), then per (tile, segment, k-chunk) one A stage:
//                         cp.async.bulk.tensor.2d (in place, producer 0) or 4 x 8 tile::gather4 (rows by id, 8 producers x 4)
// Two TMEM accumulator buffers (2 x 256 columns): the epilogue of tile i overlaps the loads and MMAs of tile i+1.
// Every mbarrier wait is bounded (a stuck pipeline traps instead of hanging the GPU).
#include "linear.cuh"
#include "umma_ptx.cuh"
#include <string.h>
#include <stdlib.h>

namespace gsage {

static constexpr int WM = 128;                 // rows per tile (UMMA M)
static constexpr int kWsEpiWarps = 8;             // two per TMEM lane quarter: they take the 32-column chunks of a tile in turn (with four,
                                                  // the epilogue -- ~3000 cycles per tile -- outlasted the MMAs of every shape with d <= 256)
static constexpr int kWsTmaWarps = 8;             // producers: a lone warp issuing 32 gather4 per stage is the bottleneck (see below)
static constexpr int kWsSplitWarps = 4;           // 3 x TF32 mode only: turn every landed fp32 A chunk into its (hi, lo) pair
static constexpr int kWsThreads = 32 * (kWsEpiWarps + 1 + kWsTmaWarps + kWsSplitWarps);
static constexpr int kWsABytes = WM * 128;     // one A stage: 128 rows x 128 bytes
static constexpr int kWsMaxStages = 12;
static constexpr int kSmemLimit = 227 * 1024;

struct WsSeg {
    const void* a; int64_t lda; const int64_t* ids;
    int d; int O; int O_store; const float* bias; int64_t col0;
    int kchunks;          // ceil(d / uk)
    int acc_col;          // first TMEM column of this segment's accumulator inside a buffer
    int w_off;            // byte offset of this segment's resident W inside the W area (kres slots of O x 128 B; x3: hi slots then lo slots)
    int kres;             // k-chunks of W resident in shared memory; chunks kres.. stream through the ring
};

struct WsParams {
    WsSeg seg[2];
    int n_phases; int phase_first[2]; int phase_count[2];
    int64_t n; int act;
    void* out; int out_bf16; int64_t ld_out;
    int n_tiles; int stages; int w_area;      // bytes reserved for the resident weights
    int stage_bytes;                          // one ring slot: an A chunk (128 x 128 B) or a streamed W chunk (O x 128 B)
    int tf32; int uk;
    int x3;                                   // 3 x TF32: D += A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T (fp32 operands, W fully resident as hi + lo)
    int* err;
    long long* dbg;                           // GSAGE_WS_TIMING builds only: per-role cycle counters of CTA 0
};

struct WsMaps { CUtensorMap w[2]; CUtensorMap a[2]; CUtensorMap g[2]; CUtensorMap wl[2]; };   // wl: the lo halves of W (3 x TF32)


#ifdef GSAGE_WS_TIMING
#define WT_DECL long long wt_a = 0, wt_b = 0, wt_c = 0, wt_t = clock64()
#define WT_LAP(x) do { const long long n_ = clock64(); (x) += n_ - wt_t; wt_t = n_; } while (0)
#define WT_OUT(base, cond) do { if (P.dbg && blockIdx.x == 0 && (cond)) { P.dbg[(base)] = wt_a; P.dbg[(base) + 1] = wt_b; P.dbg[(base) + 2] = wt_c; } } while (0)
#else
#define WT_DECL
#define WT_LAP(x)
#define WT_OUT(base, cond)
#endif

__global__ void __launch_bounds__(kWsThreads, 1) linear_ws_umma_kernel(const WsParams P, const __grid_constant__ WsMaps M) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [resident W: w_area] [stages x A 16 KB] [barriers]
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_ring = smem + P.w_area;
    uint64_t* bars = (uint64_t*)(a_ring + (size_t)P.stages * P.stage_bytes);
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * kWsMaxStages + 6);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kWsMaxStages + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kWsMaxStages + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kWsMaxStages + 2 + b); };
    const uint32_t wfull_bar = bar_base + 8u * (2 * kWsMaxStages + 4);
    const uint32_t wempty_bar = bar_base + 8u * (2 * kWsMaxStages + 5);
    auto split_bar = [&](int s) { return bar_base + 8u * (2 * kWsMaxStages + 6 + s); };     // x3: the stage's lo tile is ready

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(split_bar(s), 32 * kWsSplitWarps); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 32 * kWsEpiWarps); }
        mbar_init(wfull_bar, 1);
        mbar_init(wempty_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWsEpiWarps) {                               // the MMA warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kWsEpiWarps) {
        // =========================== EPILOGUE ===========================
        // (twelve warps -- the idle splitter warps of the non-x3 modes as a third turn -- measured 4-8 % slower than eight)
        const int quarter = warp & 3, turn = warp >> 2;      // TMEM lane quarter; which of every two 32-column chunks is mine
        const int row_in_tile = quarter * 32 + lane;         // TMEM lane == tile row
        int it = 0;
        WT_DECL;
        for (int ph = 0; ph < P.n_phases; ++ph) {
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                WT_LAP(wt_c);
                mbar_wait(tfull_bar(buf), (it >> 1) & 1, P.err);
                tc_fence_after();
                WT_LAP(wt_a);
                const int64_t row = (int64_t)tile * WM + row_in_tile;
                int chunk = 0;
                for (int si = 0; si < P.phase_count[ph]; ++si) {
                    const WsSeg& sg = P.seg[P.phase_first[ph] + si];
                    for (int c0 = 0; c0 < sg.O; c0 += 32, ++chunk) {
                        if ((chunk & 1) != turn) continue;
                        uint32_t r[32];
                        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 256 + sg.acc_col + c0), r);
                        tmem_ld_wait();
                        if (row < P.n) {
                            void* o = (char*)P.out + (row * P.ld_out + sg.col0 + c0) * (P.out_bf16 ? 2 : 4);
                            const float* bias = sg.bias ? sg.bias + c0 : nullptr;
                            const int valid = min(32, sg.O_store - c0);
                            if (valid <= 0) continue;
                            if (P.act == GSAGE_ACT_RELU) epilogue_store32<GSAGE_ACT_RELU>(r, bias, valid, o, P.out_bf16);
                            else if (P.act == GSAGE_ACT_TANH) epilogue_store32<GSAGE_ACT_TANH>(r, bias, valid, o, P.out_bf16);
                            else epilogue_store32<GSAGE_ACT_NONE>(r, bias, valid, o, P.out_bf16);
                        }
                    }
                }
                WT_LAP(wt_b);
                tc_fence_before();
                mbar_arrive(tempty_bar(buf));
            }
        }
        WT_OUT(0, threadIdx.x == 0);                         // wait tfull | tcgen05.ld + act + stores | arrive + loop
    } else if (warp == kWsEpiWarps) {
        // =========================== MMA ISSUER ===========================
        // ONE thread runs this loop, and it is the critical path of the kernel: every instruction between two stages is
        // serial latency (a lone warp issues a dependent instruction every ~5 cycles).  Measured with ncu's source view:
        // the first version (modulo / division by the runtime stage count, descriptors rebuilt per stage, parameter
        // structs indexed dynamically) spent ~1300 cycles of instructions per 16 KB stage -- 2/3 of the kernel.  Hence:
        // ring position and parity by increment, descriptors by addition, everything else hoisted.  Round 2: the thread is
        // picked by elect.sync, not `lane == 0` -- ptxas wrapped every tcgen05.mma / commit of a lane-0 region in an
        // ELECT / BRA.U.ANY loop over the active lanes (umma_ptx.cuh: elect_one).
        if (elect_one()) {
            const uint32_t fmt = P.tf32 ? 2u : 1u;              // a/b format: 1 = bf16, 2 = tf32
            const uint32_t ring16 = (smem_u32(a_ring) & 0x3FFFF) >> 4, sb16 = (uint32_t)P.stage_bytes >> 4;
            const uint64_t desc_hi = umma_desc(0);               // everything but the start address
            const uint32_t n_stages = (uint32_t)P.stages;
            uint32_t stage = 0, par = 0, a16 = ring16;           // ring slot, its parity, its address / 16
            int it = 0;
            WT_DECL;
            for (int ph = 0; ph < P.n_phases; ++ph) {
                mbar_wait(wfull_bar, ph & 1, P.err);             // this phase's resident weights have landed
                const int s_first = P.phase_first[ph], s_count = P.phase_count[ph];
                for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
                    const uint32_t buf = it & 1;
                    WT_LAP(wt_c);
                    mbar_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1, P.err);     // first use of each buffer passes immediately
                    tc_fence_after();
                    WT_LAP(wt_a);
                    for (int si = 0; si < s_count; ++si) {
                        const WsSeg& sg = P.seg[s_first + si];
                        // instruction descriptor: D = f32, A = B = bf16 | tf32, both K-major, N = O, M = 128
                        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(sg.O >> 3) << 17) | ((uint32_t)(WM >> 4) << 24);
                        const uint32_t d_tmem = tmem_base + buf * 256u + (uint32_t)sg.acc_col;
                        const uint32_t w_slot16 = ((uint32_t)sg.O * 128u) >> 4;
                        uint64_t wdesc = desc_hi | (uint64_t)((smem_u32(smem + sg.w_off) & 0x3FFFF) >> 4);
                        const int kchunks = sg.kchunks, kres = sg.kres;
                        for (int kc = 0; kc < kchunks; ++kc) {
                            WT_LAP(wt_c);
                            mbar_wait(P.x3 ? split_bar(stage) : full_bar(stage), par, P.err);
                            WT_LAP(wt_b);
                            const uint32_t a_stage = stage;
                            const uint64_t adesc = desc_hi | (uint64_t)a16;
                            if (++stage == n_stages) { stage = 0; par ^= 1; a16 = ring16; } else a16 += sb16;
                            uint64_t bdesc = wdesc;
                            uint32_t w_stage = 0;
                            const bool streamed = kc >= kres;
                            if (streamed) {                              // this W chunk came through the ring too
                                mbar_wait(full_bar(stage), par, P.err);
                                w_stage = stage;
                                bdesc = desc_hi | (uint64_t)a16;
                                if (++stage == n_stages) { stage = 0; par ^= 1; a16 = ring16; } else a16 += sb16;
                            } else {
                                wdesc += w_slot16;
                            }
                            tc_fence_after();
                            if (P.x3) {
                                // a = a_hi + a_lo, w = w_hi + w_lo (each half exact in tf32): the three products that matter
                                const uint64_t alo = adesc + (kWsABytes >> 4), wlo = bdesc + (uint64_t)(kchunks * w_slot16);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    umma_tf32(d_tmem, alo + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                                    umma_tf32(d_tmem, adesc + 2 * k, wlo + 2 * k, idesc, 1u);
                                    umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                                }
                            } else if (P.tf32) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                            } else {
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                            }
                            umma_commit(empty_bar(a_stage));             // slots reusable once these MMAs retire
                            if (streamed) umma_commit(empty_bar(w_stage));
                        }
                    }
                    umma_commit(tfull_bar(buf));                         // accumulators of this tile complete
                }
                umma_commit(wempty_bar);                                 // every MMA that reads this phase's W has retired
            }
            WT_LAP(wt_c);
            WT_OUT(4, true);                                     // wait tempty | wait for the stage (rows landed / split) | issue
        }
        __syncwarp();
    } else if (warp >= kWsEpiWarps + 1 + kWsTmaWarps) {
        // =========================== SPLITTERS (3 x TF32 only) ===========================
        // every landed fp32 A chunk becomes a (hi, lo) pair in place: hi = the value with its 13 low mantissa bits cleared
        // (exact in tf32 whatever rounding the tensor core applies to its inputs), lo = a - hi.  Elementwise, so the swizzled
        // layout is irrelevant: the 128 threads sweep the 16 KB tile (and its twin 16 KB further up) in 8 passes of 2 KB,
        // consecutive lanes on consecutive 16-byte units (a lane-per-row split would be a 32-way bank conflict: measured,
        // 6900 cycles per stage instead of ~400).
        if (P.x3) {
            const int t = threadIdx.x - 32 * (kWsEpiWarps + 1 + kWsTmaWarps);
            const uint32_t ring_u = smem_u32(a_ring), sb = (uint32_t)P.stage_bytes;
            const uint32_t n_stages = (uint32_t)P.stages;
            uint32_t stage = 0, par = 0, sa_u = ring_u;
            WT_DECL;
            for (int ph = 0; ph < P.n_phases; ++ph) {
                const int s_first = P.phase_first[ph], s_count = P.phase_count[ph];
                for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
                    for (int si = 0; si < s_count; ++si) {
                        const int kchunks = P.seg[s_first + si].kchunks;
                        for (int kc = 0; kc < kchunks; ++kc) {
                            WT_LAP(wt_c);
                            mbar_wait(full_bar(stage), par, P.err);
                            WT_LAP(wt_a);
                            const uint32_t at = sa_u + (uint32_t)t * 16u;
                            // all eight loads first, then the arithmetic and the stores: one round trip through a shared memory the
                            // tensor core is reading 8 KB per MMA from, instead of eight dependent ones (cycle counters: the split was
                            // 3200 cycles per stage and the MMA thread waited for it)
                            uint4 v[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[q].x), "=r"(v[q].y), "=r"(v[q].z), "=r"(v[q].w) : "r"(at + 2048u * q));
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const uint4 hi = make_uint4(v[q].x & 0xFFFFE000u, v[q].y & 0xFFFFE000u, v[q].z & 0xFFFFE000u, v[q].w & 0xFFFFE000u);
                                const uint4 lo = make_uint4(__float_as_uint(__uint_as_float(v[q].x) - __uint_as_float(hi.x)), __float_as_uint(__uint_as_float(v[q].y) - __uint_as_float(hi.y)),
                                                            __float_as_uint(__uint_as_float(v[q].z) - __uint_as_float(hi.z)), __float_as_uint(__uint_as_float(v[q].w) - __uint_as_float(hi.w)));
                                st_shared_v4(at + 2048u * q, hi);
                                st_shared_v4(at + (uint32_t)kWsABytes + 2048u * q, lo);
                            }
                            fence_proxy_async();                     // generic-proxy stores -> visible to the tensor core
                            mbar_arrive(split_bar(stage));
                            WT_LAP(wt_b);
                            if (++stage == n_stages) { stage = 0; par ^= 1; sa_u = ring_u; } else sa_u += sb;
                        }
                    }
                }
            }
            WT_OUT(8, t == 0);                                   // wait for the rows | split + fence + arrive | loop
        }
    } else {
        // =========================== TMA PRODUCERS ===========================
        // Eight warps walk the same stage sequence.  Producer 0 (lane 0) posts every expect_tx and issues the one-instruction
        // loads (resident W, in-place A chunks, streamed W chunks).  A chunk gathered BY ID is 32 tile::gather4 instructions
        // with per-instruction row coordinates; TMA operands live in uniform registers, so a warp issues them one lane at
        // a time (~50 cycles each: ncu shows the R2UR / UTMALDG / BRA.U.ANY loop at 43 % of a lone producer warp's samples,
        // ~1 us per 16 KB stage).  Each of the eight producers therefore gathers 16 of the 128 rows: 4 instructions per warp
        // per stage, two warps on each scheduler.  mbarrier transaction counts may go negative, so the byte completions of the other
        // producers need no ordering against producer 0's expect_tx.
        const int pw = warp - (kWsEpiWarps + 1);
        const uint32_t ring_u = smem_u32(a_ring), sb = (uint32_t)P.stage_bytes;
        const uint32_t n_stages = (uint32_t)P.stages;
        const int uk = P.uk;
        const bool lead = pw == 0 && lane == 0;
        uint32_t stage = 0, par = 1, sa_u = ring_u;              // producer parity starts at 1: fresh slots are free
        for (int ph = 0; ph < P.n_phases; ++ph) {
            const int s_first = P.phase_first[ph], s_count = P.phase_count[ph];
            if (lead) {
                mbar_wait(wempty_bar, (ph & 1) ^ 1, P.err);  // phase 0 passes immediately; later phases wait for the MMAs
                uint32_t total = 0;
                for (int si = 0; si < s_count; ++si) {
                    const WsSeg& sg = P.seg[s_first + si];
                    total += (uint32_t)sg.kres * (uint32_t)sg.O * 128u * (P.x3 ? 2u : 1u);
                }
                mbar_arrive_expect_tx(wfull_bar, total);
                for (int si = 0; si < s_count; ++si) {
                    const WsSeg& sg = P.seg[s_first + si];
                    const uint32_t w_base = smem_u32(smem + sg.w_off);
                    for (int kc = 0; kc < sg.kres; ++kc)
                        tma_load_2d(w_base + (uint32_t)kc * (uint32_t)sg.O * 128u, &M.w[s_first + si], kc * uk, 0, wfull_bar);
                    if (P.x3)
                        for (int kc = 0; kc < sg.kres; ++kc)
                            tma_load_2d(w_base + (uint32_t)(sg.kres + kc) * (uint32_t)sg.O * 128u, &M.wl[s_first + si], kc * uk, 0, wfull_bar);
                }
            }
            // the row ids of a gathered (tile, segment) are loaded one item AHEAD: a dependent global load at the top of every
            // item is ~1-2 us of latency the ring has to hide (measured in the pool kernel: profiles/r02_pool_phase_cycles.txt)
            constexpr int kRowsPerProducer = WM / kWsTmaWarps, kGatherLanes = kRowsPerProducer / 4;
            const int my_row = kRowsPerProducer * pw + 4 * lane;
            int n0 = 0, n1 = 0, n2 = 0, n3 = 0;
            auto load_ids = [&](int tile, int sidx) {             // unconditional loads from a clamped address, masked afterwards
                const int64_t* ids = tile < P.n_tiles ? P.seg[sidx].ids : nullptr;
                n0 = n1 = n2 = n3 = 0;
                if (ids && lane < kGatherLanes) {
                    const int64_t base = (int64_t)tile * WM + my_row, last = P.n - 1;
                    const int v0 = (int)__ldg(ids + min(base + 0, last)), v1 = (int)__ldg(ids + min(base + 1, last));
                    const int v2 = (int)__ldg(ids + min(base + 2, last)), v3 = (int)__ldg(ids + min(base + 3, last));
                    n0 = base + 0 <= last ? v0 : 0; n1 = base + 1 <= last ? v1 : 0;
                    n2 = base + 2 <= last ? v2 : 0; n3 = base + 3 <= last ? v3 : 0;
                }
            };
            load_ids(blockIdx.x, s_first);
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
                for (int si = 0; si < s_count; ++si) {
                    const int sidx = s_first + si;
                    const WsSeg& sg = P.seg[sidx];
                    const int64_t* ids = sg.ids;
                    const int kchunks = sg.kchunks, kres = sg.kres;
                    const uint32_t w_chunk_bytes = (uint32_t)sg.O * 128u;
                    const CUtensorMap* map_a = ids ? &M.g[sidx] : &M.a[sidx];
                    const CUtensorMap* map_w = &M.w[sidx];
                    const int tile_row = tile * WM;
                    const int r0 = n0, r1 = n1, r2 = n2, r3 = n3;
                    if (si + 1 < s_count) load_ids(tile, sidx + 1); else load_ids(tile + (int)gridDim.x, s_first);
                    if (ids) {
                        // ---- gathered operand: the first lanes of producer pw own tile rows kRowsPerProducer pw + 4 lane .. + 3 ----
                        const uint32_t row_off = (uint32_t)my_row * 128u;
                        for (int kc = 0, col = 0; kc < kchunks; ++kc, col += uk) {
                            mbar_wait(empty_bar(stage), par, P.err);
                            const uint32_t fb = full_bar(stage);
                            if (lead) mbar_arrive_expect_tx(fb, (uint32_t)kWsABytes);
                            if (lane < kGatherLanes) tma_gather4(sa_u + row_off, map_a, col, r0, r1, r2, r3, fb);
                            if (++stage == n_stages) { stage = 0; par ^= 1; sa_u = ring_u; } else sa_u += sb;
                            if (kc >= kres) {                      // this W chunk is not resident: it streams through the ring too
                                if (lead) {
                                    mbar_wait(empty_bar(stage), par, P.err);
                                    mbar_arrive_expect_tx(full_bar(stage), w_chunk_bytes);
                                    tma_load_2d(sa_u, map_w, col, 0, full_bar(stage));
                                }
                                if (++stage == n_stages) { stage = 0; par ^= 1; sa_u = ring_u; } else sa_u += sb;
                            }
                        }
                    } else {
                        // ---- operand read in place: one instruction per chunk, producer 0 alone ----
                        for (int kc = 0, col = 0; kc < kchunks; ++kc, col += uk) {
                            if (lead) {
                                mbar_wait(empty_bar(stage), par, P.err);
                                mbar_arrive_expect_tx(full_bar(stage), (uint32_t)kWsABytes);
                                tma_load_2d(sa_u, map_a, col, tile_row, full_bar(stage));
                            }
                            if (++stage == n_stages) { stage = 0; par ^= 1; sa_u = ring_u; } else sa_u += sb;
                            if (kc >= kres) {
                                if (lead) {
                                    mbar_wait(empty_bar(stage), par, P.err);
                                    mbar_arrive_expect_tx(full_bar(stage), w_chunk_bytes);
                                    tma_load_2d(sa_u, map_w, col, 0, full_bar(stage));
                                }
                                if (++stage == n_stages) { stage = 0; par ^= 1; sa_u = ring_u; } else sa_u += sb;
                            }
                        }
                    }
                }
            }
        }
    }

    // teardown: everyone done with TMEM before it is freed
    tc_fence_before();
    __syncthreads();
    if (warp == kWsEpiWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host -------------------------------------------------------------------------------------------------
static bool ws_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int ws_kchunks(const LinearSeg& s) {
    const int es = s.a_dtype == GSAGE_BF16 ? 2 : 4, uk = 128 / es;
    return (s.d + uk - 1) / uk;
}

struct WsPlan { int n_phases, w_area, stages, stage_bytes, kres[2]; };

// plan the phases: both segments fully resident together when they fit next to a deep enough ring, else one segment per
// phase with as many resident k-chunks as the ring leaves room for
static bool ws_plan(const LinearParams& P, bool x3, WsPlan* out) {
    const int fixed = 1024 /*align slack*/ + 512 /*barriers*/;
    int total = 0, cols = 0, maxO = 0;
    for (int i = 0; i < P.n_segs; ++i) {
        total += ws_kchunks(P.seg[i]) * P.seg[i].O * 128 * (x3 ? 2 : 1);       // x3: hi and lo copies
        cols += (P.seg[i].O + 31) / 32 * 32;
        maxO = P.seg[i].O > maxO ? P.seg[i].O : maxO;
    }
    WsPlan p;
    p.stage_bytes = maxO * 128 > kWsABytes ? maxO * 128 : kWsABytes;
    if (x3) p.stage_bytes = 2 * kWsABytes;                                      // the A chunk and its lo twin
    int min_stages = x3 ? 3 : 6;
    if (const char* e = getenv("GSAGE_WS_STAGES")) min_stages = atoi(e);
    if (min_stages < 3) min_stages = 3;
    if (min_stages > kWsMaxStages) min_stages = kWsMaxStages;
    const int avail = kSmemLimit - fixed - min_stages * p.stage_bytes;
    if (avail < 0) return false;
    if (cols <= 256 && total <= avail) {
        p.n_phases = 1; p.w_area = total;
        for (int i = 0; i < P.n_segs; ++i) p.kres[i] = ws_kchunks(P.seg[i]);
    } else if (x3) {
        return false;                                                           // x3 streams no weights
    } else {
        p.n_phases = P.n_segs; p.w_area = 0;
        for (int i = 0; i < P.n_segs; ++i) {
            const int slot = P.seg[i].O * 128;
            int k = avail / slot;
            if (k > ws_kchunks(P.seg[i])) k = ws_kchunks(P.seg[i]);
            p.kres[i] = k;
            if (k * slot > p.w_area) p.w_area = k * slot;
        }
    }
    p.w_area = (p.w_area + 1023) / 1024 * 1024;
    p.stages = (kSmemLimit - fixed - p.w_area) / p.stage_bytes;
    if (p.stages > kWsMaxStages) p.stages = kWsMaxStages;
    if (p.stages < 3) return false;
    *out = p;
    return true;
}

bool linear_ws_umma_eligible(const LinearParams& P) {
    if (P.n < 1 || P.pool_S > 1 || P.n_segs < 1 || P.n_segs > 2) return false;
    if (getenv("GSAGE_NO_WS")) return false;
    for (int i = 0; i < P.n_segs; ++i) {
        const LinearSeg& s = P.seg[i];
        if (s.a_dtype != P.seg[0].a_dtype || s.w_dtype != s.a_dtype) return false;         // all bf16, or all fp32 (run as TF32)
        if (s.S > 1 || s.w_trans) return false;
        if (s.O % 16 != 0 || s.O < 16 || s.O > 256) return false;
        const int es = s.a_dtype == GSAGE_BF16 ? 2 : 4, per = 16 / es;
        if (!ws_aligned16(s.a) || !ws_aligned16(s.w) || (s.lda * es) % 16 != 0 || (s.ldw * es) % 16 != 0) return false;
        if (s.lda < (s.d + per - 1) / per * per || s.ldw < (s.d + per - 1) / per * per) return false;   // whole 16-byte chunks readable
    }
    WsPlan plan;
    return ws_plan(P, false, &plan);
}

bool linear_ws_umma_x3_eligible(const LinearParams& P) {
    if (getenv("GSAGE_FP32_FFMA")) return false;
    for (int i = 0; i < P.n_segs; ++i) {
        const LinearSeg& s = P.seg[i];
        if (s.a_dtype != GSAGE_F32 || !s.w_hi || !s.w_lo || !ws_aligned16(s.w_hi) || !ws_aligned16(s.w_lo)) return false;
    }
    if (!linear_ws_umma_eligible(P)) return false;
    WsPlan plan;
    return ws_plan(P, true, &plan);
}

static int* g_ws_err = nullptr;

static int ws_launch(const LinearParams& P, bool x3, cudaStream_t s) {
    WsParams U;
    memset(&U, 0, sizeof(U));
    U.x3 = x3 ? 1 : 0;
    U.tf32 = P.seg[0].a_dtype == GSAGE_F32 ? 1 : 0;
    U.uk = U.tf32 ? 32 : 64;
    WsPlan plan;
    GS_CHECK_ARG(ws_plan(P, x3, &plan), "linear_ws_umma: operands do not fit in shared memory");
    const int n_phases = plan.n_phases;
    U.n_phases = n_phases; U.stages = plan.stages; U.w_area = plan.w_area; U.stage_bytes = plan.stage_bytes;
    for (int i = 0; i < P.n_segs; ++i) {
        const LinearSeg& g = P.seg[i];
        U.seg[i].a = g.a; U.seg[i].lda = g.lda; U.seg[i].ids = g.ids;
        U.seg[i].d = g.d; U.seg[i].O = g.O; U.seg[i].bias = g.bias; U.seg[i].col0 = g.col0;
        U.seg[i].O_store = (g.O_store > 0 && g.O_store < g.O) ? g.O_store : g.O;
        U.seg[i].kchunks = (g.d + U.uk - 1) / U.uk;
        U.seg[i].kres = plan.kres[i];
    }
    if (n_phases == 1) {
        U.phase_first[0] = 0; U.phase_count[0] = P.n_segs;
        int col = 0, off = 0;
        for (int i = 0; i < P.n_segs; ++i) {
            U.seg[i].acc_col = col; U.seg[i].w_off = off;
            col += (P.seg[i].O + 31) / 32 * 32;
            off += plan.kres[i] * P.seg[i].O * 128 * (x3 ? 2 : 1);
        }
    } else {
        for (int i = 0; i < P.n_segs; ++i) { U.phase_first[i] = i; U.phase_count[i] = 1; U.seg[i].acc_col = 0; U.seg[i].w_off = 0; }
    }
    U.n = P.n; U.act = P.act; U.out = P.out; U.out_bf16 = P.out_dtype == GSAGE_BF16; U.ld_out = P.ld_out;
    U.n_tiles = (int)ceil_div(P.n, WM);
    if (!g_ws_err) {
        GS_CUDA(cudaMalloc((void**)&g_ws_err, sizeof(int)));
        GS_CUDA(cudaMemset(g_ws_err, 0, sizeof(int)));
    }
    U.err = g_ws_err;
#ifdef GSAGE_WS_TIMING
    static long long* g_ws_dbg = nullptr;
    if (!g_ws_dbg) { GS_CUDA(cudaMalloc((void**)&g_ws_dbg, 16 * sizeof(long long))); }
    GS_CUDA(cudaMemsetAsync(g_ws_dbg, 0, 16 * sizeof(long long), s));
    U.dbg = g_ws_dbg;
#endif
    const size_t smem = (size_t)U.w_area + (size_t)U.stages * U.stage_bytes + 1024 /*align slack*/ + 512 /*barriers*/;
    GS_CHECK_ARG(smem <= (size_t)kSmemLimit, "linear_ws_umma: %zu bytes of shared memory needed", smem);
    static bool attr_set = false;
    if (!attr_set) {
        GS_CUDA(cudaFuncSetAttribute(linear_ws_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        attr_set = true;
    }
    // grid = waves x SMs, one CTA resident per SM at a time (profiling knob): > 1 hands the tiles out in smaller static shares, so an
    // SM that starts late (the next batch's sampling kernels run underneath) ends up with fewer of them, at the price of one more
    // weight load and pipeline ramp per CTA.  Two calls disagreed on the sign for the single-phase projections (-29 % / +10 %:
    // profiles/r02_persistent_waves.txt), the two-phase d = 602 projection and the pool / attention kernels were slower: default 1
    static const int env_waves = getenv("GSAGE_WS_WAVES") ? atoi(getenv("GSAGE_WS_WAVES")) : 0;
    const int waves = env_waves > 0 ? env_waves : 1;
    const int slots = sm_count() * waves;
    const int grid = U.n_tiles < slots ? U.n_tiles : slots;
    WsMaps maps;
    memset(&maps, 0, sizeof(maps));
    const int es = U.tf32 ? 4 : 2;
    for (int i = 0; i < P.n_segs; ++i) {
        const LinearSeg& g = P.seg[i];
        GS_TRY(make_map(&maps.w[i], x3 ? g.w_hi : g.w, g.O, g.d, g.ldw, g.O, es));
        if (x3) GS_TRY(make_map(&maps.wl[i], g.w_lo, g.O, g.d, g.ldw, g.O, es));
        if (!g.ids) GS_TRY(make_map(&maps.a[i], g.a, P.n, g.d, g.lda, WM, es));
        else GS_TRY(make_map(&maps.g[i], g.a, g.a_rows > 0 ? g.a_rows : 0x7FFFFFFF, g.d, g.lda, 1, es));       // rows by id
    }
    linear_ws_umma_kernel<<<grid, kWsThreads, smem, s>>>(U, maps);
#ifdef GSAGE_WS_TIMING
    if (U.n_tiles >= 8 * grid) {                                 // (big launches only)
        long long h[16];
        GS_CUDA(cudaStreamSynchronize(s));
        GS_CUDA(cudaMemcpy(h, g_ws_dbg, sizeof(h), cudaMemcpyDeviceToHost));
        const double t = (double)((U.n_tiles + grid - 1) / grid) * U.n_phases;
        int stages_per_tile = 0;
        for (int i = 0; i < P.n_segs; ++i) stages_per_tile += U.seg[i].kchunks;
        fprintf(stderr, "[ws timing] n %lld tiles/CTA %.0f phases %d x3 %d tf32 %d ring %d x %d B, %d A stages per tile | per tile, CTA 0:  epilogue warp 0: wait tfull %.0f, "
                        "ld+act+store %.0f, arrive %.0f | MMA thread: wait tempty %.0f, wait stage %.0f, issue %.0f | splitter 0: wait rows %.0f, split %.0f, loop %.0f\n",
                (long long)U.n, t, U.n_phases, U.x3, U.tf32, U.stages, U.stage_bytes, stages_per_tile, h[0] / t, h[1] / t, h[2] / t, h[4] / t, h[5] / t, h[6] / t,
                h[8] / t, h[9] / t, h[10] / t);
    }
#endif
    GS_LAUNCHED();
    return GSAGE_OK;
}

int linear_ws_umma_launch(const LinearParams& P, cudaStream_t s) { return ws_launch(P, false, s); }
int linear_ws_umma_x3_launch(const LinearParams& P, cudaStream_t s) { return ws_launch(P, true, s); }

__global__ void split_tf32_kernel(const float* __restrict__ w, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = w[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        hi[i] = h;
        lo[i] = __uint_as_float(__float_as_uint(v - h) & 0xFFFFE000u);
    }
}

int split_tf32_launch(const float* w, int64_t n, float* w_hi, float* w_lo, cudaStream_t s) {
    if (n <= 0) return GSAGE_OK;
    const int grid = (int)(ceil_div(n, 256) < 1184 ? ceil_div(n, 256) : 1184);
    split_tf32_kernel<<<grid, 256, 0, s>>>(w, n, w_hi, w_lo);
    GS_LAUNCHED();
    return GSAGE_OK;
}

}  // namespace gsage
