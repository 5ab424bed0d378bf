// linear_pool_ws_umma.cu -- the pool aggregators' per-neighbour MLP + pool on tcgen05, weights stationary.
//
// Replaces  h = relu(mlp(neibs)); h.view(N, S, H).max(dim=1)[0] | .mean(dim=1)   (nn_modules.py:224-226,240,252)
//
//   out[p, col0 + h] = reduce_{j < S} relu( A[row(p*S + j)] . W[h] + bias[h] )
//
// Same swap-AB orientation as linear_pool_umma.cu (hidden units on the M side, one hidden unit per TMEM lane, the S rows
// of a parent are S consecutive accumulator COLUMNS, the pool is a running max / sum in registers; the (N*S, H) hidden
// matrix never exists in HBM).  What changed after profiling that kernel (pokec max-pool, B = 8192: 78 % of the step,
// ~8000 cycles per 120-row tile against ~1000 cycles of MMA):
//   * W is RESIDENT in shared memory for a *phase* = as many 128-unit hidden blocks as fit (all four for d = 64; two for
//     d = 256): it used to be re-read from L2 with every tile -- 64 KB of weights for 15 KB of rows;
//   * row tiles are 64 accumulator columns wide, so four hidden blocks need 256 TMEM columns and there are TWO accumulator
//     buffers: the epilogue of tile i overlaps the MMAs of tile i+1 (it used to own all 512 columns and serialise);
//   * the epilogue is specialised at compile time for the fanouts the reference uses (S = 10, 25): parent boundaries are
//     static, one FMNMX per element, bias + relu once per parent (max_j relu(v_j + b) = relu(max_j v_j + b));
//   * four producer warps issue the tile::gather4 loads (a lone warp issues one per ~50 cycles, see linear_ws_umma.cu),
//     and the MMA thread's loop carries its ring position / descriptors by increment.
// Requires: relu, S <= 64, bf16 or fp32-as-TF32 operands, one hidden block's W (kchunks x 16 KB) <= 176 KB.
#include "linear.cuh"
#include "umma_ptx.cuh"
#include <string.h>
#include <stdlib.h>

namespace gsage {

static constexpr int QM = 128;                    // hidden units per block (UMMA M)
static constexpr int QN = 64;                     // accumulator columns (neighbour rows) per tile (UMMA N)
static constexpr int kQMaxBlocks = 4;             // hidden blocks per phase: 4 x 64 columns x 2 buffers = all of TMEM
static constexpr int kQEpiWarps = 16;             // four per TMEM lane quarter: one hidden block each
static constexpr int kQTmaWarps = 4;
static constexpr int kQThreads = 32 * (kQEpiWarps + 1 + kQTmaWarps);
static constexpr int kQWChunk = QM * 128;         // 16 KB: one hidden block x one k-chunk of W
static constexpr int kQRChunk = QN * 128;         // 8 KB: one k-chunk of a row tile
static constexpr int kQStages = 8;
static constexpr int kQSmemLimit = 227 * 1024;
static constexpr int kQIdBatch = 8;              // tiles per batch of row ids a producer warp stages in shared memory
static constexpr int kQIdBytes = kQTmaWarps * kQIdBatch * 16 * 4;

struct QParams {
    const void* a; int64_t lda; const int64_t* ids;
    const float* bias; int64_t col0;
    int d, H, S, R, PT;                           // R = PT * S rows of a tile are used
    int64_t n_rows, n_parents;
    int n_tiles, h_blocks, bpp, n_phases, kchunks, uk, tf32, dbg_skip;
    void* out; int out_bf16; int64_t ld_out;
    // backward mode: recompute the hidden rows and emit d loss / d hidden (bf16) from d loss / d pooled
    const float* dP; int64_t ld_dp; __nv_bfloat16* dhid; int64_t ld_dhid;
    float* db;                                    // += column sums of d hidden (the MLP bias gradient); may be NULL
    int* err;
    long long* dbg;                               // GSAGE_POOL_TIMING builds only: per-role cycle counters of CTA 0
};

struct QMaps { CUtensorMap w; CUtensorMap a; };

// one parent's pooled value -> HBM (relu applied here for the max pool: once per parent instead of once per element)
template <bool POOL_MAX>
__device__ __forceinline__ void q_store(const QParams& P, int64_t parent, int h, float acc, float bias) {
    const float o = POOL_MAX ? fmaxf(acc + bias, 0.0f) : acc * (1.0f / (float)P.S);
    const int64_t at = parent * P.ld_out + P.col0 + h;
    if (P.out_bf16) reinterpret_cast<__nv_bfloat16*>(P.out)[at] = __float2bfloat16_rn(o);
    else reinterpret_cast<float*>(P.out)[at] = o;
}

// backward epilogue for one parent whose S accumulator values are v[0..S): gradient of the pooled value `g` goes to the FIRST
// row that attains the max (torch.max's backward; rows drawn twice are identical, so which copy gets it does not change
// dW) and only if relu let it through; the mean pool spreads g / S over the rows whose relu is open
template <bool POOL_MAX, int S>
__device__ __forceinline__ float q_backward_parent(const QParams& P, const float* v, float bias, float g, int64_t row0, int h, bool h_ok) {
    float mx = -3.0e38f, total = 0.0f;
#pragma unroll
    for (int j = 0; j < S; ++j) mx = fmaxf(mx, v[j]);
    const bool gate = mx + bias > 0.0f;
    bool found = false;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        float o;
        if (POOL_MAX) {
            const bool hit = gate && !found && v[j] == mx;
            found = found || hit;
            o = hit ? g : 0.0f;
        } else {
            o = (v[j] + bias > 0.0f) ? g * (1.0f / (float)S) : 0.0f;
        }
        if (h_ok && row0 + j < P.n_rows) { P.dhid[(row0 + j) * P.ld_dhid + h] = __float2bfloat16_rn(o); total += o; }
    }
    return total;
}

#ifdef GSAGE_POOL_TIMING
#define QT_DECL long long qt_a = 0, qt_b = 0, qt_c = 0, qt_d = 0, qt_t = clock64()
#define QT_LAP(x) do { const long long n_ = clock64(); (x) += n_ - qt_t; qt_t = n_; } while (0)
#define QT_OUT(base, cond) do { if (P.dbg && blockIdx.x == 0 && (cond)) { P.dbg[(base)] = qt_a; P.dbg[(base) + 1] = qt_b; P.dbg[(base) + 2] = qt_c; P.dbg[(base) + 3] = qt_d; } } while (0)
#else
#define QT_DECL
#define QT_LAP(x)
#define QT_OUT(base, cond)
#endif

template <bool POOL_MAX, int S_CT, bool BWD>
__global__ void __launch_bounds__(kQThreads, 1) linear_pool_ws_umma_kernel(const QParams P, const __grid_constant__ QMaps M) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [W of a phase: bpp x kchunks x 16 KB] [ring: 8 x 8 KB row chunks] [barriers]
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* ring = smem + (size_t)P.bpp * P.kchunks * kQWChunk;
    uint64_t* bars = (uint64_t*)(ring + kQStages * kQRChunk);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kQStages + 6);
    int* ids_s = (int*)(bars + 32);                           // [producer][tile of the batch][16 row ids]
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kQStages + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kQStages + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kQStages + 2 + b); };
    const uint32_t wfull_bar = bar_base + 8u * (2 * kQStages + 4), wempty_bar = bar_base + 8u * (2 * kQStages + 5);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kQStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 32 * kQEpiWarps); }
        mbar_init(wfull_bar, 1); mbar_init(wempty_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kQEpiWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kQEpiWarps) {
        // ============ EPILOGUE: one hidden unit per thread, the pool runs along the accumulator columns ============
        const int quarter = warp & 3, slot = warp >> 2;        // TMEM lane quarter; hidden block of the phase
        const int S = S_CT > 0 ? S_CT : P.S;
        const int PT = S_CT > 0 ? QN / S_CT : P.PT;
        int it = 0;
        QT_DECL;
        for (int ph = 0; ph < P.n_phases; ++ph) {
            const int hb = ph * P.bpp + slot;
            const bool mine = slot < P.bpp && hb < P.h_blocks;
            const int h = hb * QM + quarter * 32 + lane;
            const bool h_ok = mine && h < P.H;
            const float bias = (P.bias && h_ok) ? __ldg(P.bias + h) : 0.0f;
            float bsum = 0.0f;                                 // backward: this hidden unit's bias gradient over the CTA's tiles
            for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                QT_LAP(qt_d);
                mbar_wait(tfull_bar(buf), (it >> 1) & 1, P.err);
                tc_fence_after();
                QT_LAP(qt_a);
#ifdef GSAGE_POOL_TIMING
                if (mine && !P.dbg_skip) {
#else
                if (mine) {
#endif
                    const int64_t parent0 = (int64_t)tile * PT;
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 256 + slot * QN);
                    uint32_t r0[32], r1[32];
                    tmem_ld32(taddr, r0);
                    tmem_ld32(taddr + 32, r1);
                    tmem_ld_wait();
                    QT_LAP(qt_b);
                    if (BWD) {
                        // ---- backward: d hidden[row, h] from d pooled[parent, h] (the forward's MMAs recomputed) ----
                        float v[QN];
#pragma unroll
                        for (int i = 0; i < QN; ++i) v[i] = __uint_as_float(i < 32 ? r0[i & 31] : r1[i & 31]);
                        if (S_CT > 0) {
#pragma unroll
                            for (int p = 0; p < QN / (S_CT > 0 ? S_CT : 1); ++p) {
                                const bool p_ok = parent0 + p < P.n_parents;
                                const float g = (h_ok && p_ok) ? __ldg(P.dP + (parent0 + p) * P.ld_dp + h) : 0.0f;
                                bsum += q_backward_parent<POOL_MAX, (S_CT > 0 ? S_CT : 1)>(P, v + p * S_CT, bias, g, (parent0 + p) * (int64_t)S_CT, h, h_ok && p_ok);
                            }
                        } else {
                            for (int p = 0; p < PT; ++p) {               // odd fanouts: dynamic indexing (local memory), correctness first
                                const bool p_ok = parent0 + p < P.n_parents;
                                const float g = (h_ok && p_ok) ? __ldg(P.dP + (parent0 + p) * P.ld_dp + h) : 0.0f;
                                float mx = -3.0e38f;
                                for (int j = 0; j < S; ++j) mx = fmaxf(mx, v[p * S + j]);
                                const bool gate = mx + bias > 0.0f;
                                bool found = false;
                                for (int j = 0; j < S; ++j) {
                                    const float x = v[p * S + j];
                                    float o;
                                    if (POOL_MAX) { const bool hit = gate && !found && x == mx; found = found || hit; o = hit ? g : 0.0f; }
                                    else o = (x + bias > 0.0f) ? g * (1.0f / (float)S) : 0.0f;
                                    const int64_t row = (parent0 + p) * (int64_t)S + j;
                                    if (h_ok && p_ok && row < P.n_rows) { P.dhid[row * P.ld_dhid + h] = __float2bfloat16_rn(o); bsum += o; }
                                }
                            }
                        }
                    } else
                    if (S_CT > 0) {
                        // static parent boundaries: element i of the tile belongs to parent i / S_CT
                        constexpr int SS = S_CT > 0 ? S_CT : 1, NP = QN / SS;
                        float acc[NP];
#pragma unroll
                        for (int p = 0; p < NP; ++p) {
                            acc[p] = POOL_MAX ? -3.0e38f : 0.0f;
#pragma unroll
                            for (int j = 0; j < SS; ++j) {
                                const int i = p * SS + j;
                                const float v = __uint_as_float(i < 32 ? r0[i & 31] : r1[i & 31]);
                                if (POOL_MAX) acc[p] = fmaxf(acc[p], v);
                                else acc[p] += fmaxf(v + bias, 0.0f);
                            }
                            acc[p] = POOL_MAX ? fmaxf(acc[p] + bias, 0.0f) : acc[p] * (1.0f / (float)SS);
                        }
                        // stores: one base address per tile, a stride per parent; whole tiles (all but the last) skip the bounds checks.
                        // (The first version rebuilt a 64-bit address and a bounds predicate per parent: ~110 of its ~140 instructions.)
                        if (h_ok) {
                            const int np = parent0 + NP <= P.n_parents ? NP : (int)(P.n_parents - parent0);
                            const int64_t at = parent0 * P.ld_out + P.col0 + h;
                            if (P.out_bf16) {
                                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(P.out) + at;
                                if (np == NP) {
#pragma unroll
                                    for (int p = 0; p < NP; ++p) o[p * P.ld_out] = __float2bfloat16_rn(acc[p]);
                                } else {
#pragma unroll
                                    for (int p = 0; p < NP; ++p) if (p < np) o[p * P.ld_out] = __float2bfloat16_rn(acc[p]);
                                }
                            } else {
                                float* o = reinterpret_cast<float*>(P.out) + at;
                                if (np == NP) {
#pragma unroll
                                    for (int p = 0; p < NP; ++p) o[p * P.ld_out] = acc[p];
                                } else {
#pragma unroll
                                    for (int p = 0; p < NP; ++p) if (p < np) o[p * P.ld_out] = acc[p];
                                }
                            }
                        }
                    } else {
                        float acc = POOL_MAX ? -3.0e38f : 0.0f;
                        int cnt = 0, p = 0;
#pragma unroll
                        for (int i = 0; i < QN; ++i) {
                            if (i < P.R) {
                                const float v = __uint_as_float(i < 32 ? r0[i & 31] : r1[i & 31]);
                                if (POOL_MAX) acc = fmaxf(acc, v);
                                else acc += fmaxf(v + bias, 0.0f);
                                if (++cnt == S) {
                                    if (h_ok && parent0 + p < P.n_parents) q_store<POOL_MAX>(P, parent0 + p, h, acc, bias);
                                    ++p; cnt = 0;
                                    acc = POOL_MAX ? -3.0e38f : 0.0f;
                                }
                            }
                        }
                    }
                }
                QT_LAP(qt_c);
                tc_fence_before();
                mbar_arrive(tempty_bar(buf));
            }
            if (BWD && P.db && h_ok) atomicAdd(P.db + h, bsum);
        }
        QT_OUT(0, threadIdx.x == 0);                           // wait tfull | tcgen05.ld + wait | pool + stores | arrive + loop
    } else if (warp == kQEpiWarps) {
        // ============ MMA ISSUER: D[buf][blk][hidden, row] += W[blk][kc] . rows[kc]^T, one elected thread, lean loop ============
        // (elect_one, not lane == 0: see umma_ptx.cuh -- this loop was 2600 of the kernel's 2900 cycles per tile before)
        if (elect_one()) {
            const uint32_t fmt = P.tf32 ? 2u : 1u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(QN >> 3) << 17) | ((uint32_t)(QM >> 4) << 24);
            const uint64_t desc_hi = umma_desc(0);
            const uint32_t ring16 = (smem_u32(ring) & 0x3FFFF) >> 4, w16 = (smem_u32(smem) & 0x3FFFF) >> 4;
            const int kchunks = P.kchunks;
            const bool tf32 = P.tf32 != 0;
            uint32_t stage = 0, par = 0, b16 = ring16;
            int it = 0;
            QT_DECL;
            for (int ph = 0; ph < P.n_phases; ++ph) {
                const int nb = min(P.bpp, P.h_blocks - ph * P.bpp);
                mbar_wait(wfull_bar, ph & 1, P.err);
                for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
                    const uint32_t buf = it & 1;
                    QT_LAP(qt_c);
                    mbar_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1, P.err);
                    tc_fence_after();
                    QT_LAP(qt_a);
                    const uint32_t d0 = tmem_base + buf * 256u;
                    for (int kc = 0; kc < kchunks; ++kc) {
                        mbar_wait(full_bar(stage), par, P.err);
                        tc_fence_after();
                        QT_LAP(qt_b);
                        const uint64_t bdesc = desc_hi | (uint64_t)b16;
                        uint64_t adesc = desc_hi | (uint64_t)(w16 + (uint32_t)kc * (kQWChunk >> 4));
                        const uint32_t acc = kc ? 1u : 0u;
                        if (tf32) {
                            for (int j = 0; j < nb; ++j, adesc += (uint64_t)(kchunks * (kQWChunk >> 4))) {
                                umma_tf32(d0 + j * QN, adesc, bdesc, idesc, acc);
#pragma unroll
                                for (int k = 1; k < 4; ++k) umma_tf32(d0 + j * QN, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                            }
                        } else {
                            for (int j = 0; j < nb; ++j, adesc += (uint64_t)(kchunks * (kQWChunk >> 4))) {
                                umma_bf16(d0 + j * QN, adesc, bdesc, idesc, acc);
#pragma unroll
                                for (int k = 1; k < 4; ++k) umma_bf16(d0 + j * QN, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                            }
                        }
                        umma_commit(empty_bar(stage));
                        if (++stage == kQStages) { stage = 0; par ^= 1; b16 = ring16; } else b16 += (kQRChunk >> 4);
                    }
                    umma_commit(tfull_bar(buf));
                }
                umma_commit(wempty_bar);
            }
            QT_LAP(qt_c);
            QT_OUT(4, true);                                   // wait tempty | wait full (rows landed) | issue MMAs + commits
        }
        __syncwarp();
    } else {
        // ============ TMA PRODUCERS: producer pw owns tile rows 16 pw .. 16 pw + 15 ============
        // One elected lane per warp issues (its four gather4 of a k-chunk back to back).  The row ids are the catch: a global
        // load per tile in that lane's loop is ~2000 cycles of exposed latency per tile (measured: the MMA thread waited
        // 1060 cycles per tile for rows) -- so the WHOLE warp loads the ids of kQIdBatch tiles at once (4 per lane), one batch
        // ahead, and parks them in shared memory, where the issuing lane picks up its 16 per tile.
        const int pw = warp - (kQEpiWarps + 1);
        const bool lead = pw == 0;
        const uint32_t ring_u = smem_u32(ring);
        const int uk = P.uk, kchunks = P.kchunks;
        const int my_row = 16 * pw;
        const uint32_t row_off = (uint32_t)my_row * 128u;
        int* my_ids = ids_s + pw * (kQIdBatch * 16);
        uint32_t stage = 0, par = 1, sa_u = ring_u;            // (only the elected lane's copy advances; elect.sync picks the same lane every time)
        QT_DECL;
        // lane -> (tile j of the batch, ids 4 i .. 4 i + 3 of my 16): unconditional loads from a clamped address, masked afterwards
        const int bj = lane >> 2, bi = (lane & 3) * 4;
        auto load_batch = [&](int k0, int* r) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)(k0 + bj) * gridDim.x;
            const int64_t base = tile * P.R + my_row + bi;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = tile < P.n_tiles && my_row + bi + i < P.R && base + i < P.n_rows;
                const int v = (int)__ldg(P.ids + (ok ? base + i : 0));
                r[i] = ok ? v : 0;
            }
        };
        for (int ph = 0; ph < P.n_phases; ++ph) {
            if (lead && elect_one()) {
                const int nb = min(P.bpp, P.h_blocks - ph * P.bpp);
                mbar_wait(wempty_bar, (ph & 1) ^ 1, P.err);
                mbar_arrive_expect_tx(wfull_bar, (uint32_t)(nb * kchunks * kQWChunk));
                for (int j = 0; j < nb; ++j)
                    for (int kc = 0; kc < kchunks; ++kc)
                        tma_load_2d(smem_u32(smem) + (uint32_t)(j * kchunks + kc) * kQWChunk, &M.w, kc * uk, (ph * P.bpp + j) * QM, wfull_bar);
            }
            __syncwarp();
            if (P.ids) {
                int regs[4];
                load_batch(0, regs);
                for (int k0 = 0; (int64_t)blockIdx.x + (int64_t)k0 * gridDim.x < P.n_tiles; k0 += kQIdBatch) {
                    *reinterpret_cast<int4*>(my_ids + bj * 16 + bi) = make_int4(regs[0], regs[1], regs[2], regs[3]);
                    __syncwarp();
                    load_batch(k0 + kQIdBatch, regs);
                    if (elect_one()) {
                        for (int j = 0; j < kQIdBatch; ++j) {
                            if ((int64_t)blockIdx.x + (int64_t)(k0 + j) * gridDim.x >= P.n_tiles) break;
                            int cur[16];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int4 v = *reinterpret_cast<const int4*>(my_ids + j * 16 + 4 * q);
                                cur[4 * q] = v.x; cur[4 * q + 1] = v.y; cur[4 * q + 2] = v.z; cur[4 * q + 3] = v.w;
                            }
                            for (int kc = 0, col = 0; kc < kchunks; ++kc, col += uk) {
                                QT_LAP(qt_c);
                                mbar_wait(empty_bar(stage), par, P.err);
                                QT_LAP(qt_a);
                                const uint32_t fb = full_bar(stage);
                                if (lead) mbar_arrive_expect_tx(fb, (uint32_t)kQRChunk);
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    tma_gather4(sa_u + row_off + (uint32_t)q * 512u, &M.a, col, cur[4 * q], cur[4 * q + 1], cur[4 * q + 2], cur[4 * q + 3], fb);
                                QT_LAP(qt_b);
                                if (++stage == kQStages) { stage = 0; par ^= 1; sa_u = ring_u; } else sa_u += kQRChunk;
                            }
                        }
                    }
                    __syncwarp();
                }
            } else if (lead && elect_one()) {
                for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
                    const int64_t row0 = (int64_t)tile * P.R;
                    for (int kc = 0, col = 0; kc < kchunks; ++kc, col += uk) {
                        mbar_wait(empty_bar(stage), par, P.err);
                        mbar_arrive_expect_tx(full_bar(stage), (uint32_t)kQRChunk);
                        tma_load_2d(sa_u, &M.a, col, (int)row0, full_bar(stage));
                        if (++stage == kQStages) { stage = 0; par ^= 1; sa_u = ring_u; } else sa_u += kQRChunk;
                    }
                }
            }
            __syncwarp();
        }
        QT_OUT(8, lead && lane == 0);                          // wait empty | expect_tx + gather4 issue | loop + ids from smem
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kQEpiWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host -------------------------------------------------------------------------------------------------
static int q_blocks_per_phase(const LinearParams& P) {
    const LinearSeg& s = P.seg[0];
    const int es = s.a_dtype == GSAGE_BF16 ? 2 : 4, uk = 128 / es;
    const int kchunks = (s.d + uk - 1) / uk;
    const int budget = kQSmemLimit - 2048 - kQIdBytes - kQStages * kQRChunk;
    int bpp = budget / (kchunks * kQWChunk);
    const int h_blocks = (s.O + QM - 1) / QM;
    if (bpp > kQMaxBlocks) bpp = kQMaxBlocks;
    if (bpp > h_blocks) bpp = h_blocks;
    return bpp;
}

bool linear_pool_ws_umma_eligible(const LinearParams& P) {
    if (getenv("GSAGE_NO_POOL_WS")) return false;
    if (!linear_pool_umma_eligible(P)) return false;
    if (P.act != GSAGE_ACT_RELU || P.pool_S > QN) return false;
    return q_blocks_per_phase(P) >= 1;
}

static int* g_q_err = nullptr;

static int q_launch(const LinearParams& P, const float* dP, int64_t ld_dp, void* dhid, int64_t ld_dhid, float* db, cudaStream_t s) {
    const LinearSeg& g = P.seg[0];
    const bool bwd = dhid != nullptr;
    QParams U;
    memset(&U, 0, sizeof(U));
    U.a = g.a; U.lda = g.lda; U.ids = g.ids; U.bias = g.bias; U.col0 = g.col0;
    U.d = g.d; U.H = g.O; U.S = P.pool_S;
    U.PT = QN / U.S; U.R = U.PT * U.S;
    U.n_rows = P.n; U.n_parents = P.n / P.pool_S;
    U.tf32 = g.a_dtype == GSAGE_F32 ? 1 : 0;
    U.uk = U.tf32 ? 32 : 64;
    U.kchunks = (U.d + U.uk - 1) / U.uk;
    U.n_tiles = (int)ceil_div(U.n_parents, U.PT);
    U.h_blocks = (U.H + QM - 1) / QM;
    U.bpp = q_blocks_per_phase(P);
    GS_CHECK_ARG(U.bpp >= 1, "linear_pool_ws_umma: one hidden block of W does not fit in shared memory");
    U.n_phases = (U.h_blocks + U.bpp - 1) / U.bpp;
    U.out = P.out; U.out_bf16 = P.out_dtype == GSAGE_BF16; U.ld_out = P.ld_out;
    U.dP = dP; U.ld_dp = ld_dp; U.dhid = (__nv_bfloat16*)dhid; U.ld_dhid = ld_dhid; U.db = db;
    if (!g_q_err) {
        GS_CUDA(cudaMalloc((void**)&g_q_err, sizeof(int)));
        GS_CUDA(cudaMemset(g_q_err, 0, sizeof(int)));
    }
    U.err = g_q_err;
#ifdef GSAGE_POOL_TIMING
    static long long* g_q_dbg = nullptr;
    if (!g_q_dbg) { GS_CUDA(cudaMalloc((void**)&g_q_dbg, 16 * sizeof(long long))); }
    GS_CUDA(cudaMemsetAsync(g_q_dbg, 0, 16 * sizeof(long long), s));
    U.dbg = g_q_dbg;
    U.dbg_skip = getenv("GSAGE_POOL_DBG_SKIP") ? atoi(getenv("GSAGE_POOL_DBG_SKIP")) : 0;   // 1: the epilogue only waits and arrives
#endif
    const int es = U.tf32 ? 4 : 2;
    QMaps maps;
    memset(&maps, 0, sizeof(maps));
    GS_TRY(make_map(&maps.w, g.w, g.O, g.d, g.ldw, QM, es));
    if (g.ids) GS_TRY(make_map(&maps.a, g.a, g.a_rows > 0 ? g.a_rows : 0x7FFFFFFF, g.d, g.lda, 1, es));
    else GS_TRY(make_map(&maps.a, g.a, P.n, g.d, g.lda, QN, es));
    const size_t smem = (size_t)U.bpp * U.kchunks * kQWChunk + (size_t)kQStages * kQRChunk + 1024 + 512 + kQIdBytes;
    // grid = waves x SMs, one CTA resident per SM at a time (profiling knob): > 1 hands the tiles out in smaller static shares; measured slower for this kernel
    // (profiles/r02_persistent_waves.txt), unlike the single-phase projections (linear_ws_umma.cu)
    static const int waves = getenv("GSAGE_POOL_WAVES") ? atoi(getenv("GSAGE_POOL_WAVES")) : 1;
    const int slots = sm_count() * (waves < 1 ? 1 : waves);
    const int grid = U.n_tiles < slots ? U.n_tiles : slots;
#define GS_Q_LAUNCH(MX, SCT, BW)                                                                                                   \
    do {                                                                                                                           \
        static bool attr_set = false;                                                                                              \
        if (!attr_set) {                                                                                                           \
            GS_CUDA(cudaFuncSetAttribute(linear_pool_ws_umma_kernel<MX, SCT, BW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQSmemLimit)); \
            attr_set = true;                                                                                                       \
        }                                                                                                                          \
        linear_pool_ws_umma_kernel<MX, SCT, BW><<<grid, kQThreads, smem, s>>>(U, maps);                                            \
    } while (0)
#define GS_Q_PICK(SCT)                                                                                                             \
    do {                                                                                                                           \
        if (bwd) { if (mx) GS_Q_LAUNCH(true, SCT, true); else GS_Q_LAUNCH(false, SCT, true); }                                     \
        else { if (mx) GS_Q_LAUNCH(true, SCT, false); else GS_Q_LAUNCH(false, SCT, false); }                                       \
    } while (0)
    const bool mx = P.pool_max != 0;
    if (U.S == 10) GS_Q_PICK(10);
    else if (U.S == 25) GS_Q_PICK(25);
    else GS_Q_PICK(0);
#undef GS_Q_PICK
#undef GS_Q_LAUNCH
#ifdef GSAGE_POOL_TIMING
    {
        long long h[16];
        GS_CUDA(cudaStreamSynchronize(s));
        GS_CUDA(cudaMemcpy(h, g_q_dbg, sizeof(h), cudaMemcpyDeviceToHost));
        const double t = (double)((U.n_tiles + grid - 1) / grid) * U.n_phases;
        fprintf(stderr, "[pool timing] n_tiles %d (%.0f per CTA) skip %d | per tile, CTA 0:  epilogue warp 0: wait tfull %.0f, tcgen05.ld %.0f, pool+store %.0f, arrive %.0f | "
                        "MMA thread: wait tempty %.0f, wait rows %.0f, issue %.0f | producer 0: wait empty %.0f, issue gather4 %.0f, ids+loop %.0f\n",
                U.n_tiles, t, U.dbg_skip, h[0] / t, h[1] / t, h[2] / t, h[3] / t, h[4] / t, h[5] / t, h[6] / t, h[8] / t, h[9] / t, h[10] / t);
    }
#endif
    GS_LAUNCHED();
    return GSAGE_OK;
}

int linear_pool_ws_umma_launch(const LinearParams& P, cudaStream_t s) { return q_launch(P, nullptr, 0, nullptr, 0, nullptr, s); }

// backward of the fused MLP + pool: dhid[(p*S + j), h] (bf16, ld_dhid) = d loss / d (pre-relu hidden) from dP[p, h] = d loss / d pooled;
// db[h] += sum over rows of dhid[:, h] (the caller zeroes db)
int linear_pool_ws_umma_backward_launch(const LinearParams& P, const float* dP, int64_t ld_dp, void* dhid, int64_t ld_dhid, float* db,
                                        cudaStream_t s) {
    GS_CHECK_ARG(dP && dhid && ld_dhid >= P.seg[0].O && ld_dp >= P.seg[0].O, "linear_pool_ws_umma_backward: bad arguments");
    GS_CHECK_ARG(linear_pool_ws_umma_eligible(P), "linear_pool_ws_umma_backward: operands do not qualify for the weight-stationary pool kernel");
    return q_launch(P, dP, ld_dp, dhid, ld_dhid, db, s);
}

}  // namespace gsage
