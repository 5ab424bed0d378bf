// mt19937.cu -- numpy's legacy MT19937 stream, resident on the device.
//
// Replaces the global `np.random` RandomState the reference seeds in helpers.set_seeds
// (/root/reference/helpers.py:14-18) and draws from in SparseUniformNeighborSampler.__call__
// (/root/reference/nn_modules.py:88, `np.random.choice(maxdeg, (n, S))`) and NodeProblem.iterate
// (/root/reference/problem.py:146, `np.random.permutation`).  Same words, same order, same position
// hand-off (get_state / set_state), so sampled indices are bit-exact under the same seed.
//
// HBM layout
//   ring    uint32 [cap]   UNTEMPERED words of the stream, word i at ring[i & (cap-1)], cap = 2^k.
//                          Untempered words are also the generator state: any 624 consecutive,
//                          block-aligned words are a numpy `key`.
//   cursor  int64  [2]     stream index of the next unconsumed word, double-buffered (a consuming kernel
//                          reads slot p and writes slot p^1; the host flips p per call and never reads it
//                          back unless asked -- consumption is data dependent (masked rejection) and is
//                          resolved entirely on the device).
// Host bookkeeping: lower / upper bounds on the cursor (count <= words consumed <= window) decide how
// far ahead to generate; they are re-tightened with one 8-byte read-back only when they drift apart by
// more than the ring can hold.
#include "mt19937.cuh"
#include "mt_jump.h"
#include <vector>

#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>

namespace gsage {

static constexpr int kN = 624;
static constexpr int kM = 397;

__host__ __device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}

__host__ __device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    return y;
}

// ---------------------------------------------------------------------------------------------
// generation: blocks [start, start + 624*nblocks) from the 624 words before `start`.
// One CTA walks the blocks in order; inside a block the twist has three dependency-free phases
// (227 + 227 + 170 words).  TODO(round 2): jump-ahead lanes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mt_generate_kernel(uint32_t* __restrict__ ring, uint64_t cap_mask,
                                                          int64_t start, int nblocks) {
    __shared__ uint32_t st[2][kN];
    const int tid = threadIdx.x;
    for (int i = tid; i < kN; i += 256) st[0][i] = ring[(uint64_t)(start - kN + i) & cap_mask];
    __syncthreads();
    int cur = 0;
    for (int b = 0; b < nblocks; ++b) {
        const uint32_t* o = st[cur];
        uint32_t* n = st[cur ^ 1];
        if (tid < 227) n[tid] = o[tid + kM] ^ mt_mix(o[tid], o[tid + 1]);
        __syncthreads();
        if (tid < 227) n[tid + 227] = n[tid] ^ mt_mix(o[tid + 227], o[tid + 228]);
        __syncthreads();
        if (tid < 169) n[tid + 454] = n[tid + 227] ^ mt_mix(o[tid + 454], o[tid + 455]);
        else if (tid == 169) n[623] = n[396] ^ mt_mix(o[623], n[0]);
        __syncthreads();
        const int64_t base = start + (int64_t)b * kN;
        for (int i = tid; i < kN; i += 256) ring[(uint64_t)(base + i) & cap_mask] = n[i];
        cur ^= 1;
    }
}

// ---------------------------------------------------------------------------------------------
// lane-parallel refill (jump-ahead).  Lane l generates blocks [l*K, (l+1)*K) of the refill; its start
// window is the head window H jumped ahead by l*K*624 words:  g_l(F) H = XOR_{i: g_l[i]=1} x[i .. i+624)
// where x is the stream continuing H (mt_jump.cpp).  The XOR is split over kJumpSlices CTAs per lane
// (each takes 2496 of the 19968 polynomial bits) and the partial windows are combined by the lane.
// ---------------------------------------------------------------------------------------------
static constexpr int kJumpSlices = 8;
static constexpr int kSliceWords = 78;                   // 78 x 32 = 2496 bits; 8 x 78 = 624 words
static constexpr int kSeqWords = 19968 + kN;             // words of x a slice may touch

__device__ __forceinline__ void mt_next_block(const uint32_t* o, uint32_t* n, int tid) {
    if (tid < 227) n[tid] = o[tid + kM] ^ mt_mix(o[tid], o[tid + 1]);
    __syncthreads();
    if (tid < 227) n[tid + 227] = n[tid] ^ mt_mix(o[tid + 227], o[tid + 228]);
    __syncthreads();
    if (tid < 169) n[tid + 454] = n[tid + 227] ^ mt_mix(o[tid + 454], o[tid + 455]);
    else if (tid == 169) n[623] = n[396] ^ mt_mix(o[623], n[0]);
    __syncthreads();
}

// grid (lanes-1, kJumpSlices): partial[(lane-1)*kJumpSlices + slice][624]
// x = the stream continuing the head window H: x[0..624) = H (the block before `start`), x[624 + j] = word start + j.  The same x
// serves every lane, and its 33 blocks are simply the first 33 blocks of the refill -- so they are generated ONCE, by one CTA,
// straight into the ring (mt_generate_kernel, queued right before this kernel), and every (lane, slice) CTA only loads the
// 3120 words its slice touches.  (Round 1 regenerated up to 33 blocks sequentially in each of the 248 CTAs: the jump kernel cost
// as much GPU time as generating the 5 M words themselves.)  The XOR itself takes four set bits per iteration: the loop is
// bound by the latency of the shared-memory loads, four windows in flight cost the same as one; missing bits of the last group
// point at a zero pad.
static constexpr int kJumpPrefixBlocks = kSeqWords / kN;                 // 33
static constexpr int kSliceSpan = kSliceWords * 32 + kN;                 // 3120 words of x per slice
static constexpr int kZeroPad = 1024;
__global__ void __launch_bounds__(256) mt_jump_kernel(const uint32_t* __restrict__ ring, uint64_t cap_mask, int64_t start,
                                                      const uint32_t* __restrict__ polys, uint32_t* __restrict__ partial) {
    __shared__ uint32_t xs[kSliceSpan + kZeroPad];
    __shared__ uint32_t gw[kSliceWords];
    const int tid = threadIdx.x, lane = blockIdx.x + 1, slice = blockIdx.y;
    const int64_t first = start - kN + (int64_t)slice * kSliceWords * 32;
    for (int i = tid; i < kSliceSpan; i += 256) xs[i] = ring[(uint64_t)(first + i) & cap_mask];
    for (int i = tid; i < kZeroPad; i += 256) xs[kSliceSpan + i] = 0u;
    if (tid < kSliceWords) gw[tid] = polys[(size_t)(lane - 1) * kN + slice * kSliceWords + tid];
    __syncthreads();
    uint32_t a0 = 0, a1 = 0, a2 = 0;
    const bool third = tid + 512 < kN;
    const int t2 = third ? tid + 512 : tid;                // (threads past the window's end re-read a word they already hold: a2 is dropped)
    for (int w = 0; w < kSliceWords; ++w) {
        uint32_t bits = gw[w];                             // uniform across the CTA: no divergence
        const int base = w * 32;
        while (bits) {
            int off[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (bits) { off[k] = base + __ffs(bits) - 1; bits &= bits - 1; }
                else off[k] = kSliceSpan;                  // the zero pad
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t* p = xs + off[k];
                a0 ^= p[tid];
                a1 ^= p[tid + 256];
                a2 ^= p[t2];
            }
        }
    }
    uint32_t* out = partial + ((size_t)(lane - 1) * kJumpSlices + slice) * kN;
    out[tid] = a0;
    out[tid + 256] = a1;
    if (third) out[tid + 512] = a2;
}

// grid (lanes): lane l writes stream words [start + l*K*624, start + (l+1)*K*624).  Lane 0 continues after the
// `skip0` blocks mt_generate_kernel already wrote (the jump kernel's x).
__global__ void __launch_bounds__(256) mt_generate_lanes_kernel(uint32_t* __restrict__ ring, uint64_t cap_mask,
                                                                int64_t start, int blocks_per_lane,
                                                                const uint32_t* __restrict__ partial, int skip0) {
    __shared__ uint32_t st[2][kN];
    const int tid = threadIdx.x, lane = blockIdx.x;
    const int b0 = lane == 0 ? skip0 : 0;
    for (int i = tid; i < kN; i += 256) {
        uint32_t v;
        if (lane == 0) {
            v = ring[(uint64_t)(start + (int64_t)(b0 - 1) * kN + i) & cap_mask];
        } else {
            v = 0;
            const uint32_t* p = partial + (size_t)(lane - 1) * kJumpSlices * kN + i;
#pragma unroll
            for (int s = 0; s < kJumpSlices; ++s) v ^= p[s * kN];
        }
        st[0][i] = v;
    }
    __syncthreads();
    int cur = 0;
    const int64_t base0 = start + (int64_t)lane * blocks_per_lane * kN;
    for (int b = b0; b < blocks_per_lane; ++b) {
        mt_next_block(st[cur], st[cur ^ 1], tid);
        const int64_t base = base0 + (int64_t)b * kN;
        for (int i = tid; i < kN; i += 256) ring[(uint64_t)(base + i) & cap_mask] = st[cur ^ 1][i];
        cur ^= 1;
    }
}

// ---------------------------------------------------------------------------------------------
// consumption
// ---------------------------------------------------------------------------------------------
static constexpr int kTileThreads = 256;
static constexpr int kTileRows = 8;
static constexpr int kTile = kTileThreads * kTileRows;   // 2048 words per CTA

// p == 1 path (mask == rng, incl. raw words): out[i] = temper(word[cursor + i]) & mask; cursor += count
__global__ void __launch_bounds__(256) mt_take_kernel(const uint32_t* __restrict__ ring, uint64_t cap_mask,
                                                      const int64_t* __restrict__ cursor_in,
                                                      int64_t* __restrict__ cursor_out, int64_t count, uint32_t mask,
                                                      uint32_t* __restrict__ out) {
    const int64_t cur = *cursor_in;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = i0; i < count; i += stride) out[i] = mt_temper(ring[(uint64_t)(cur + i) & cap_mask]) & mask;
    if (i0 == 0) *cursor_out = cur + count;
}

// pass 1: accepted words per tile
__global__ void __launch_bounds__(kTileThreads) mt_count_kernel(const uint32_t* __restrict__ ring, uint64_t cap_mask,
                                                                const int64_t* __restrict__ cursor_in, int64_t window,
                                                                uint32_t mask, uint32_t rng_max,
                                                                int* __restrict__ tile_count) {
    __shared__ int warp_tot[kTileThreads / 32];
    const int64_t cur = *cursor_in;
    const int64_t tile_base = (int64_t)blockIdx.x * kTile;
    int mine = 0;
#pragma unroll
    for (int j = 0; j < kTileRows; ++j) {
        const int64_t i = tile_base + j * kTileThreads + threadIdx.x;
        if (i < window) mine += ((mt_temper(ring[(uint64_t)(cur + i) & cap_mask]) & mask) <= rng_max) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
#pragma unroll
        for (int w = 0; w < kTileThreads / 32; ++w) s += warp_tot[w];
        tile_count[blockIdx.x] = s;
    }
}

// pass 2: exclusive scan of the tile counts (one CTA); flags a short window
__global__ void __launch_bounds__(1024) mt_scan_kernel(const int* __restrict__ tile_count, int64_t* __restrict__ tile_off,
                                                       int n_tiles, int64_t count, int* __restrict__ err_flag) {
    __shared__ int64_t warp_sum[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int t = base + threadIdx.x;
        const int64_t v = (t < n_tiles) ? tile_count[t] : 0;
        int64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int64_t w = warp_sum[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t y = __shfl_up_sync(0xFFFFFFFFu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            warp_sum[threadIdx.x] = w;   // inclusive
        }
        __syncthreads();
        const int64_t before = carry + ((threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0);
        if (t < n_tiles) tile_off[t] = before + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0 && carry < count) { *(volatile int*)err_flag = 1; __threadfence_system(); }   // window too short: never silently wrong
}

// pass 3: stream compaction -- the q-th accepted word (q < count) goes to out[q]; the word after the
// count-th accepted one is the new cursor
__global__ void __launch_bounds__(kTileThreads) mt_scatter_kernel(const uint32_t* __restrict__ ring, uint64_t cap_mask,
                                                                  const int64_t* __restrict__ cursor_in,
                                                                  int64_t* __restrict__ cursor_out, int64_t window,
                                                                  uint32_t mask, uint32_t rng_max, int64_t count,
                                                                  const int64_t* __restrict__ tile_off,
                                                                  uint32_t* __restrict__ out) {
    __shared__ int tot[kTileRows * (kTileThreads / 32)];     // [row][warp] -> exclusive base after the scan
    const int64_t cur = *cursor_in;
    const int64_t tile_base = (int64_t)blockIdx.x * kTile;
    const int64_t rank0 = tile_off[blockIdx.x];
    if (rank0 >= count) return;                               // whole tile is past the last needed draw
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t val[kTileRows];
    int pre[kTileRows];
    bool acc[kTileRows];
#pragma unroll
    for (int j = 0; j < kTileRows; ++j) {
        const int64_t i = tile_base + j * kTileThreads + threadIdx.x;
        uint32_t w = 0;
        bool a = false;
        if (i < window) {
            w = mt_temper(ring[(uint64_t)(cur + i) & cap_mask]) & mask;
            a = (w <= rng_max);
        }
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, a);
        val[j] = w; acc[j] = a;
        pre[j] = __popc(ballot & ((1u << lane) - 1u));
        if (lane == 0) tot[j * (kTileThreads / 32) + warp] = __popc(ballot);
    }
    __syncthreads();
    if (threadIdx.x < 32) {                                    // exclusive scan of the 64 (row, warp) totals
        int a0 = tot[2 * threadIdx.x], a1 = tot[2 * threadIdx.x + 1];
        int s = a0 + a1, x = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (threadIdx.x >= o) x += y;
        }
        const int excl = x - s;
        tot[2 * threadIdx.x] = excl;
        tot[2 * threadIdx.x + 1] = excl + a0;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kTileRows; ++j) {
        if (!acc[j]) continue;
        const int64_t q = rank0 + tot[j * (kTileThreads / 32) + warp] + pre[j];
        if (q < count) {
            out[q] = val[j];
            if (q == count - 1) *cursor_out = cur + tile_base + j * kTileThreads + threadIdx.x + 1;
        }
    }
}

// Fisher-Yates from the top (np.random.permutation): inherently sequential -- one thread walks the stream
__global__ void __launch_bounds__(256) mt_permutation_kernel(const uint32_t* __restrict__ ring, uint64_t cap_mask,
                                                             const int64_t* __restrict__ cursor_in,
                                                             int64_t* __restrict__ cursor_out, int64_t n,
                                                             int64_t limit, int64_t* __restrict__ out,
                                                             int* __restrict__ err_flag) {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) out[i] = i;
    __syncthreads();
    if (threadIdx.x != 0) return;
    int64_t c = *cursor_in;
    for (int64_t i = n - 1; i > 0; --i) {
        uint32_t m = (uint32_t)i;
        m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16;
        uint32_t w;
        do {
            if (c >= limit) { *(volatile int*)err_flag = 1; __threadfence_system(); *cursor_out = c; return; }
            w = mt_temper(ring[(uint64_t)c & cap_mask]) & m;
            ++c;
        } while (w > (uint32_t)i);
        const int64_t a = out[i], b = out[w];
        out[i] = b; out[w] = a;
    }
    *cursor_out = c;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static void host_init_genrand(uint32_t seed, uint32_t* mt) {
    mt[0] = seed;
    for (int i = 1; i < kN; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
}

static uint32_t smear(uint32_t v) {
    v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    return v;
}

int rng_ensure(gsage_rng* r, int64_t upto, cudaStream_t s);

// re-tighten the host bounds with the true cursor (8-byte read-back; syncs the stream)
static int rng_resync(gsage_rng* r, cudaStream_t s) {
    int64_t c = 0;
    GS_CUDA(cudaMemcpyAsync(&c, r->cursor + r->parity, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    GS_CUDA(cudaStreamSynchronize(s));
    r->cursor_lb = r->cursor_ub = c;
    for (auto& f : r->fb) f.pending = false;                   // older than this exact value
    return GSAGE_OK;
}

// lazily computed jump polynomials (host, ~0.1 s once per process) + device scratch for the lane refill
static int rng_init_lanes(gsage_rng* r) {
    if (r->lanes_ready) return GSAGE_OK;
    std::vector<uint32_t> polys((size_t)(r->lanes - 1) * kN);
    if (mt_jump_poly_series((uint64_t)r->lane_blocks * kN, r->lanes - 1, polys.data()) != 0) {
        set_error("rng: MT19937 characteristic polynomial computation failed");
        return GSAGE_ERR_RNG;
    }
    GS_CUDA(cudaMalloc((void**)&r->polys, sizeof(uint32_t) * polys.size()));
    GS_CUDA(cudaMemcpy(r->polys, polys.data(), sizeof(uint32_t) * polys.size(), cudaMemcpyHostToDevice));
    GS_CUDA(cudaMalloc((void**)&r->partial, sizeof(uint32_t) * (size_t)(r->lanes - 1) * kJumpSlices * kN));
    GS_CUDA(cudaFuncSetAttribute(mt_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(uint32_t) * kSeqWords)));
    r->lanes_ready = true;
    return GSAGE_OK;
}

// fold completed cursor read-backs into the host bounds (never blocks)
static void rng_poll(gsage_rng* r) {
    for (auto& f : r->fb) {
        if (!f.pending || cudaEventQuery(f.ev) != cudaSuccess) continue;
        const int64_t c = *f.host;
        // draws queued after the read-back was posted consumed >= (acc_total - acc_at) and <= (win_total - win_at) words
        r->cursor_lb = std::max(r->cursor_lb, c + (r->acc_total - f.acc_at));
        r->cursor_ub = std::min(r->cursor_ub, c + (r->win_total - f.win_at));
        f.pending = false;
    }
}

// post a read-back of the cursor as it is after everything queued on `s` so far
static void rng_post_feedback(gsage_rng* r, cudaStream_t s) {
    for (auto& f : r->fb) {
        if (f.pending || !f.host) continue;
        if (cudaMemcpyAsync(f.host, r->cursor + r->parity, sizeof(int64_t), cudaMemcpyDeviceToHost, s) != cudaSuccess) return;
        if (cudaEventRecord(f.ev, s) != cudaSuccess) return;
        f.acc_at = r->acc_total; f.win_at = r->win_total; f.pending = true;
        return;
    }
}

// Enqueue generation of stream words up to (at least) `upto` on the rng's own side stream.  The refill first waits
// for everything already queued on the caller's stream (those kernels may still read ring slots the refill is
// about to recycle), then runs concurrently with whatever the caller queues next -- the generator is independent
// of the graph data, so it overlaps the HBM-bound gather/aggregate kernels of the same step.
// `may_sync`: allowed to read the true cursor back (synchronises the caller's stream) when the host bounds are too
// loose to prove the ring has room; a prefetch passes false and simply gives up.
static int rng_generate(gsage_rng* r, int64_t upto, cudaStream_t s, bool may_sync) {
    bool fenced = false;
    rng_poll(r);
    while (upto > r->gen_end) {
        const int64_t need = ceil_div(upto - r->gen_end, kN);
        const int64_t lane_refill = (int64_t)r->lanes * r->lane_blocks;
        const bool use_lanes = r->lanes > 1 && need >= r->lane_threshold;
        int64_t blocks = use_lanes ? lane_refill : std::max<int64_t>(need, r->prefetch_blocks);
        // words from (cursor_lb - 624) must survive: the block under the cursor is the numpy `key`
        if (r->gen_end + blocks * kN - (r->cursor_lb - kN) > r->cap) {
            if (!may_sync) return GSAGE_OK;
            GS_TRY(rng_resync(r, s));
            const int64_t room = (r->cap - (r->gen_end - (r->cursor_lb - kN))) / kN;
            if (room < need) {
                set_error("rng: ring of %lld words cannot hold a look-ahead of %lld", (long long)r->cap,
                          (long long)(upto - r->cursor_lb));
                return GSAGE_ERR_RNG;
            }
            if (blocks > room) blocks = use_lanes ? -1 : room;
        }
        if (!fenced) {
            if (!r->overlap) r->side_now = s;                    // GSAGE_RNG_OVERLAP=0: refill in line on the caller's stream
            else {
                r->side_now = r->side;
                GS_CUDA(cudaEventRecord(r->ev_main, s));
                GS_CUDA(cudaStreamWaitEvent(r->side, r->ev_main, 0));
            }
            fenced = true;
        }
        if (use_lanes && blocks == lane_refill) {
            GS_TRY(rng_init_lanes(r));
            // x for the jumps = the first 33 blocks of the refill, generated once (one CTA, ~12 us), then read by every jump CTA
            mt_generate_kernel<<<1, 256, 0, r->side_now>>>(r->ring, (uint64_t)(r->cap - 1), r->gen_end, kJumpPrefixBlocks);
            GS_LAUNCHED();
            mt_jump_kernel<<<dim3(r->lanes - 1, kJumpSlices), 256, 0, r->side_now>>>(
                r->ring, (uint64_t)(r->cap - 1), r->gen_end, r->polys, r->partial);
            GS_LAUNCHED();
            mt_generate_lanes_kernel<<<r->lanes, 256, 0, r->side_now>>>(r->ring, (uint64_t)(r->cap - 1), r->gen_end, r->lane_blocks,
                                                                    r->partial, kJumpPrefixBlocks);
            GS_LAUNCHED();
        } else {
            if (blocks < 0) blocks = std::min<int64_t>(need, (r->cap - (r->gen_end - (r->cursor_lb - kN))) / kN);
            mt_generate_kernel<<<1, 256, 0, r->side_now>>>(r->ring, (uint64_t)(r->cap - 1), r->gen_end, (int)blocks);
            GS_LAUNCHED();
        }
        r->gen_end += blocks * kN;
    }
    if (fenced) GS_CUDA(cudaEventRecord(r->ev_refill, r->side_now));
    return GSAGE_OK;
}

// make sure stream words [.., upto) exist in the ring AND are visible to kernels queued on `s` from now on
int rng_ensure(gsage_rng* r, int64_t upto, cudaStream_t s) {
    GS_TRY(rng_generate(r, upto, s, true));
    if (upto > r->gen_visible) {
        GS_CUDA(cudaStreamWaitEvent(s, r->ev_refill, 0));       // ev_refill covers everything generated so far
        r->gen_visible = r->gen_end;
    }
    return GSAGE_OK;
}

// top the ring up for the next calls without making the caller's stream wait for it
static int rng_prefetch(gsage_rng* r, cudaStream_t s) {
    // far enough ahead that the refill for the NEXT batch is already done when its draws are queued (ring permitting)
    const int64_t ahead = std::min<int64_t>(2 * r->max_window + r->max_window / 2 + kN, r->cap / 2);
    return rng_generate(r, r->cursor_ub + ahead, s, false);
}

// window of raw words that holds `count` accepted draws with overwhelming probability (12 sigma)
static int64_t window_for(int64_t count, double p) {
    if (p >= 1.0) return count;
    const double mean = (double)count / p;
    const double sd = sqrt((double)count * (1.0 - p)) / p;
    return (int64_t)(mean + 12.0 * sd + 64.0);
}

// Consumers may alternate between streams.  Every consuming call ends with rng_leave (an event on its stream); the first
// call on a different stream waits for that event -- no handle of the previous stream is kept (it may be gone by then).
int rng_enter(gsage_rng* r, cudaStream_t s) {
    if (r->have_last && r->last_stream != s) GS_CUDA(cudaStreamWaitEvent(s, r->ev_switch, 0));
    r->last_stream = s;
    r->have_last = true;
    return GSAGE_OK;
}

static int rng_leave(gsage_rng* r, cudaStream_t s) {
    GS_CUDA(cudaEventRecord(r->ev_switch, s));
    return GSAGE_OK;
}

static int rng_draw(gsage_rng* r, uint32_t hi, int64_t count, uint32_t* out, cudaStream_t s) {
    const uint32_t rng_max = hi - 1u;
    if (count == 0) return GSAGE_OK;
    GS_TRY(rng_enter(r, s));
    if (rng_max == 0) {                       // numpy: rng == 0 -> zeros, no word consumed
        GS_CUDA(cudaMemsetAsync(out, 0, sizeof(uint32_t) * count, s));
        return GSAGE_OK;
    }
    const uint32_t mask = smear(rng_max);
    const double p = ((double)rng_max + 1.0) / ((double)mask + 1.0);
    const int64_t max_piece = std::max<int64_t>(1024, (int64_t)((double)(r->cap / 4) * p * 0.9));
    for (int64_t done = 0; done < count; done += max_piece) {
        const int64_t n = std::min(max_piece, count - done);
        const int64_t window = window_for(n, p);
        GS_TRY(rng_ensure(r, r->cursor_ub + window, s));
        const int64_t* cin = r->cursor + r->parity;
        int64_t* cout = r->cursor + (r->parity ^ 1);
        if (mask == rng_max) {
            const int grid = (int)std::min<int64_t>(ceil_div(n, 256 * 4), 148 * 16);
            mt_take_kernel<<<grid, 256, 0, s>>>(r->ring, (uint64_t)(r->cap - 1), cin, cout, n, mask, out + done);
            GS_LAUNCHED();
        } else {
            const int n_tiles = (int)ceil_div(window, kTile);
            if (n_tiles > r->tiles_cap) {
                set_error("rng: internal tile scratch too small (%d > %d)", n_tiles, r->tiles_cap);
                return GSAGE_ERR_RNG;
            }
            mt_count_kernel<<<n_tiles, kTileThreads, 0, s>>>(r->ring, (uint64_t)(r->cap - 1), cin, window, mask, rng_max,
                                                             r->tile_count);
            GS_LAUNCHED();
            mt_scan_kernel<<<1, 1024, 0, s>>>(r->tile_count, r->tile_off, n_tiles, n, r->err_flag);
            GS_LAUNCHED();
            mt_scatter_kernel<<<n_tiles, kTileThreads, 0, s>>>(r->ring, (uint64_t)(r->cap - 1), cin, cout, window, mask,
                                                               rng_max, n, r->tile_off, out + done);
            GS_LAUNCHED();
        }
        r->parity ^= 1;
        r->cursor_lb += n;
        r->cursor_ub += window;
        r->acc_total += n; r->win_total += window;
        r->max_window = std::max(r->max_window, window);
    }
    GS_TRY(rng_leave(r, s));
    rng_post_feedback(r, s);
    return rng_prefetch(r, s);
}

int rng_randint_internal(gsage_rng* r, uint32_t hi, int64_t count, uint32_t* out, cudaStream_t s) {
    return rng_draw(r, hi, count, out, s);
}

}  // namespace gsage

using namespace gsage;

extern "C" {

int gsage_rng_create(gsage_rng** out) {
    GS_CHECK_ARG(out, "rng_create: NULL out");
    gsage_rng* r = new gsage_rng();
    int log2cap = 25;                                           // 32 Mi words = 128 MiB of look-ahead
    if (const char* e = getenv("GSAGE_RNG_LOG2_WORDS")) log2cap = std::max(14, std::min(30, atoi(e)));
    r->cap = (int64_t)1 << log2cap;
    r->prefetch_blocks = std::max<int64_t>(1, std::min<int64_t>(64, r->cap / kN / 8));
    // lane refill = lanes x lane_blocks x 624 words (default 32 x 256 -> 5.1 M words); rings too small for it stay sequential
    r->lanes = 32; r->lane_blocks = 256; r->lane_threshold = 128;
    if (const char* e = getenv("GSAGE_RNG_OVERLAP")) r->overlap = atoi(e) != 0;
    if (const char* e = getenv("GSAGE_RNG_LANES")) r->lanes = std::max(1, std::min(148, atoi(e)));
    if (const char* e = getenv("GSAGE_RNG_LANE_BLOCKS")) r->lane_blocks = std::max(64, std::min(4096, atoi(e)));      // (>= the 33 prefix blocks)
    if ((int64_t)r->lanes * r->lane_blocks * kN > r->cap / 3) r->lanes = 1;
    r->tiles_cap = (int)(r->cap / kTile + 2);
    cudaError_t e1 = cudaMalloc((void**)&r->ring, sizeof(uint32_t) * r->cap);
    cudaError_t e2 = cudaMalloc((void**)&r->cursor, sizeof(int64_t) * 2);
    cudaError_t e3 = cudaHostAlloc((void**)&r->err_flag, sizeof(int), cudaHostAllocMapped);   // mapped pinned: polled by the host without a copy
    if (e3 == cudaSuccess) *r->err_flag = 0; else r->err_flag = nullptr;
    cudaError_t e4 = cudaMalloc((void**)&r->tile_count, sizeof(int) * r->tiles_cap);
    cudaError_t e5 = cudaMalloc((void**)&r->tile_off, sizeof(int64_t) * r->tiles_cap);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess || e5 != cudaSuccess) {
        set_error("rng_create: cudaMalloc failed");
        gsage_rng_destroy(r);
        return GSAGE_ERR_NOMEM;
    }
    int prio_lo = 0, prio_hi = 0;                               // refills are tiny next to the gather kernels they overlap:
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);       // highest priority, so their CTAs are placed as slots free up
    if (cudaStreamCreateWithPriority(&r->side, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&r->ev_main, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&r->ev_refill, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&r->ev_switch, cudaEventDisableTiming) != cudaSuccess) {
        set_error("rng_create: stream / event creation failed");
        gsage_rng_destroy(r);
        return GSAGE_ERR_CUDA;
    }
    for (auto& f : r->fb) {
        if (cudaHostAlloc((void**)&f.host, sizeof(int64_t), cudaHostAllocDefault) != cudaSuccess ||
            cudaEventCreateWithFlags(&f.ev, cudaEventDisableTiming) != cudaSuccess) {
            set_error("rng_create: pinned feedback slot allocation failed");
            gsage_rng_destroy(r);
            return GSAGE_ERR_NOMEM;
        }
    }
    *out = r;
    return gsage_rng_seed(r, 5489u, nullptr);
}

void gsage_rng_destroy(gsage_rng* r) {
    if (!r) return;
    cudaFree(r->ring); cudaFree(r->cursor); if (r->err_flag) cudaFreeHost(r->err_flag); cudaFree(r->tile_count); cudaFree(r->tile_off);
    cudaFree(r->polys); cudaFree(r->partial);
    if (r->side) { cudaStreamSynchronize(r->side); cudaStreamDestroy(r->side); }
    if (r->ev_main) cudaEventDestroy(r->ev_main);
    if (r->ev_refill) cudaEventDestroy(r->ev_refill);
    if (r->ev_switch) cudaEventDestroy(r->ev_switch);
    for (auto& f : r->fb) { if (f.ev) { cudaEventSynchronize(f.ev); cudaEventDestroy(f.ev); } if (f.host) cudaFreeHost(f.host); }
    delete r;
}

int gsage_rng_set_state(gsage_rng* r, const uint32_t* key, int pos, void* stream) {
    GS_CHECK_ARG(r && key && pos >= 0 && pos <= kN, "rng_set_state: bad arguments (pos must be in [0, 624])");
    cudaStream_t s = as_stream(stream);
    // the given key becomes stream block 0; earlier launches (either stream) may still touch the ring
    GS_CUDA(cudaStreamSynchronize(r->side));
    if (r->have_last && r->last_stream != s) GS_CUDA(cudaEventSynchronize(r->ev_switch));   // draws queued on another stream
    GS_CUDA(cudaStreamSynchronize(s));
    r->last_stream = s; r->have_last = false;
    GS_CUDA(cudaMemcpyAsync(r->ring, key, sizeof(uint32_t) * kN, cudaMemcpyHostToDevice, s));
    const int64_t c[2] = {pos, pos};
    GS_CUDA(cudaMemcpyAsync(r->cursor, c, sizeof(c), cudaMemcpyHostToDevice, s));
    GS_CUDA(cudaStreamSynchronize(s));                          // `key` / `c` are pageable host memory
    *(volatile int*)r->err_flag = 0;                            // mapped host memory; every earlier consumer has retired
    r->gen_end = kN;
    r->gen_visible = kN;
    r->max_window = 0;
    r->parity = 0;
    r->cursor_lb = r->cursor_ub = pos;
    r->origin = pos;
    for (auto& f : r->fb) { if (f.pending) cudaEventSynchronize(f.ev); f.pending = false; }
    return GSAGE_OK;
}

int gsage_rng_seed(gsage_rng* r, uint32_t seed, void* stream) {
    GS_CHECK_ARG(r, "rng_seed: NULL rng");
    uint32_t key[kN];
    host_init_genrand(seed, key);
    return gsage_rng_set_state(r, key, kN, stream);            // numpy: pos = 624 right after seeding
}

int gsage_rng_get_state(gsage_rng* r, uint32_t* key, int* pos, void* stream) {
    GS_CHECK_ARG(r && key && pos, "rng_get_state: NULL argument");
    cudaStream_t s = as_stream(stream);
    GS_TRY(rng_enter(r, s));
    GS_TRY(gsage_rng_check(r, stream));
    GS_TRY(rng_resync(r, s));
    const int64_t c = r->cursor_lb;
    // numpy reports (key of the block the last word came from, pos in 1..624); c == 0 only after set_state(pos=0)
    const int64_t blk = c > 0 ? (c - 1) / kN : 0;
    *pos = (int)(c - blk * kN);
    const uint64_t mask = (uint64_t)(r->cap - 1);
    const uint64_t lo = (uint64_t)(blk * kN) & mask;
    const int64_t first = std::min<int64_t>(kN, r->cap - (int64_t)lo);
    GS_CUDA(cudaMemcpyAsync(key, r->ring + lo, sizeof(uint32_t) * first, cudaMemcpyDeviceToHost, s));
    if (first < kN) GS_CUDA(cudaMemcpyAsync(key + first, r->ring, sizeof(uint32_t) * (kN - first), cudaMemcpyDeviceToHost, s));
    GS_CUDA(cudaStreamSynchronize(s));
    return GSAGE_OK;
}

int gsage_rng_raw(gsage_rng* r, int64_t count, uint32_t* out_dev, void* stream) {
    GS_CHECK_ARG(r && count >= 0 && (out_dev || count == 0), "rng_raw: bad arguments");
    cudaStream_t s = as_stream(stream);
    GS_TRY(rng_enter(r, s));
    const int64_t max_piece = r->cap / 4;
    for (int64_t done = 0; done < count; done += max_piece) {
        const int64_t n = std::min(max_piece, count - done);
        GS_TRY(rng_ensure(r, r->cursor_ub + n, s));
        const int grid = (int)std::min<int64_t>(ceil_div(n, 256 * 4), 148 * 16);
        mt_take_kernel<<<grid, 256, 0, s>>>(r->ring, (uint64_t)(r->cap - 1), r->cursor + r->parity,
                                            r->cursor + (r->parity ^ 1), n, 0xFFFFFFFFu, out_dev + done);
        GS_LAUNCHED();
        r->parity ^= 1;
        r->cursor_lb += n;
        r->cursor_ub += n;
        r->acc_total += n; r->win_total += n;
    }
    return rng_leave(r, s);
}

int gsage_rng_randint(gsage_rng* r, uint32_t hi, int64_t count, uint32_t* out_dev, void* stream) {
    GS_CHECK_ARG(r && hi >= 1 && count >= 0 && (out_dev || count == 0), "rng_randint: bad arguments (hi >= 1)");
    return rng_draw(r, hi, count, out_dev, as_stream(stream));
}

int gsage_rng_permutation(gsage_rng* r, int64_t n, int64_t* out_dev, void* stream) {
    GS_CHECK_ARG(r && n >= 0 && (out_dev || n == 0), "rng_permutation: bad arguments");
    if (n == 0) return GSAGE_OK;
    cudaStream_t s = as_stream(stream);
    GS_TRY(rng_enter(r, s));
    // every step needs < 2 words on average; 2n + 12 sigma + slack bounds the walk
    const int64_t window = std::min<int64_t>(2 * n + (int64_t)(12.0 * sqrt(2.0 * (double)n)) + 256, r->cap / 2);
    GS_TRY(rng_ensure(r, r->cursor_ub + window, s));
    mt_permutation_kernel<<<1, 256, 0, s>>>(r->ring, (uint64_t)(r->cap - 1), r->cursor + r->parity,
                                            r->cursor + (r->parity ^ 1), n, r->gen_end, out_dev, r->err_flag);
    GS_LAUNCHED();
    r->parity ^= 1;
    r->cursor_lb += 0;                       // n == 1 consumes nothing; the true value comes from resync
    r->cursor_ub += window;
    r->win_total += window;
    GS_TRY(rng_leave(r, s));
    return rng_resync(r, s);                  // sequential kernel anyway: re-tighten immediately
}

int gsage_rng_check(gsage_rng* r, void* stream) {
    GS_CHECK_ARG(r, "rng_check: NULL rng");
    GS_CUDA(cudaStreamSynchronize(as_stream(stream)));   // sticky flag in mapped host memory: no ordering needed against the other consumer stream
    if (*(volatile int*)r->err_flag) {
        set_error("rng: the look-ahead window held fewer accepted draws than requested (12-sigma event) -- "
                  "re-seed; results since the last check are invalid");
        return GSAGE_ERR_RNG;
    }
    return GSAGE_OK;
}

int gsage_rng_consumed(gsage_rng* r, int64_t* words, void* stream) {
    GS_CHECK_ARG(r && words, "rng_consumed: NULL argument");
    GS_TRY(rng_enter(r, as_stream(stream)));
    GS_TRY(rng_resync(r, as_stream(stream)));
    *words = r->cursor_lb - r->origin;
    return GSAGE_OK;
}

}  // extern "C"
