// peer_allreduce.cu -- the ONE collective of the path: the sum all-reduce of the parameter gradients over the ranks
// of one box, written against NVLink / NVSwitch peer memory instead of calling NCCL.
//
// The gradient bucket is ~0.6-0.9 MB: at that size an all-reduce is pure latency, and NCCL's ring / tree
// protocols cost ~0.1 ms at 8 ranks -- round 1 measured that tail as the whole loss of scaling of the train step (efficiency 0.95 at
// N = 8).  Here every rank's bucket lives in symmetric memory (the same allocation mapped into every process of the box), and
// ONE kernel per rank does a one-shot all-reduce: it waits until every peer has entered the kernel (flags in the peers' buffers,
// release / acquire at system scope), then reads all `world` buckets directly over NVLink -- 16-byte loads, peers visited in a
// rank-rotated order so that the eight readers do not hit the same GPU at the same time -- sums them with each rank's
// local/global batch weight, writes the reduced gradient into the rank's own (private) optimiser buffer, and accumulates the
// squared norm that clip_grad_norm needs (models.py:102) on the way: collective + norm in one pass over the data.  A second flag
// round at the end keeps a rank from leaving -- and its next backward from overwriting the bucket -- while a peer still reads it.
//
// Layout of a rank's symmetric buffer (fp32 words): [ n gradient words | scale | pad to 32 words | start flags: 16 u32 | end
// flags: 16 u32 ].  Flags carry the call's epoch (1, 2, ...; every rank makes the same sequence of calls).
#include "common.cuh"

namespace gsage {

static constexpr int kPeerMax = 16;

struct PeerArgs {
    float* buf[kPeerMax];
    int world, rank;
    int64_t n;                 // gradient words (multiple of 4)
    int64_t flag_word;         // word offset of the start flags
    uint32_t epoch;
    float scale;
    float* out; float* sumsq;
    unsigned int* done;        // local counter of finished blocks
    int* err;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {      // plain (coherent) 16-byte load: the data was written by another GPU
    float4 v;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
// bounded spin: a rank that never shows up must not hang the box
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t epoch, int* err) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(p) - epoch) < 0) {
        if (clock64() - t0 > 20000000000LL) { if (err) atomicExch(err, 1); __trap(); }       // ~10 s
        __nanosleep(64);
    }
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(const PeerArgs A) {
    __shared__ float sc[kPeerMax];
    __shared__ float ws[8];
    __shared__ int is_last;
    float* mine = A.buf[A.rank];
    uint32_t* my_flags = reinterpret_cast<uint32_t*>(mine) + A.flag_word;
    const int t = threadIdx.x;
    // ---- entry: publish my weight, tell every peer I am here, wait until every peer is ----
    if (blockIdx.x == 0) {
        if (t == 0) { mine[A.n] = A.scale; __threadfence_system(); }
        __syncthreads();
        if (t < A.world) st_release_sys(reinterpret_cast<uint32_t*>(A.buf[t]) + A.flag_word + A.rank, A.epoch);
    }
    if (t < A.world) {
        wait_flag(my_flags + t, A.epoch, A.err);
        sc[t] = (t == A.rank) ? A.scale : *reinterpret_cast<volatile float*>(A.buf[t] + A.n);
    }
    __syncthreads();
    // ---- one-shot reduce: out[i] = sum_p scale_p * buf_p[i] ----
    float ss = 0.0f;
    const int64_t n4 = A.n >> 2, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + t; i < n4; i += stride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int k = 0; k < A.world; ++k) {
            int p = A.rank + k; if (p >= A.world) p -= A.world;
            const float4 v = ld_peer_v4(A.buf[p] + 4 * i);
            const float s = sc[p];
            acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
        }
        *reinterpret_cast<float4*>(A.out + 4 * i) = acc;
        ss += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
    }
    if (A.sumsq) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
        if ((t & 31) == 0) ws[t >> 5] = ss;
        __syncthreads();
        if (t == 0) {
            float tot = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += ws[w];
            atomicAdd(A.sumsq, tot);
        }
    }
    // ---- exit: the last block of this rank tells every peer "I have read your bucket" and waits for the same from them ----
    __syncthreads();
    if (t == 0) {
        __threadfence();
        is_last = atomicAdd(A.done, 1u) == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (is_last) {
        if (t < A.world) {
            st_release_sys(reinterpret_cast<uint32_t*>(A.buf[t]) + A.flag_word + kPeerMax + A.rank, A.epoch);
            wait_flag(my_flags + kPeerMax + t, A.epoch, A.err);
        }
        if (t == 0) *A.done = 0;
    }
}

static int* g_peer_err = nullptr;
static unsigned int* g_peer_done = nullptr;

}  // namespace gsage

using namespace gsage;

extern "C" {

int64_t gsage_peer_allreduce_words(int64_t n) { return (n + 3) / 4 * 4 + 32 + 2 * kPeerMax; }

int gsage_peer_allreduce(const uint64_t* peer_ptrs_host, int world, int rank, int64_t n, uint32_t epoch, float scale, float* out_dev,
                         float* sumsq_dev, void* stream) {
    GS_CHECK_ARG(peer_ptrs_host && out_dev && world >= 1 && world <= kPeerMax && rank >= 0 && rank < world && n > 0 && n % 4 == 0 && epoch > 0,
                 "peer_allreduce: bad arguments (world <= %d, n a multiple of 4, epoch >= 1)", kPeerMax);
    cudaStream_t s = as_stream(stream);
    if (!g_peer_err) {
        GS_CUDA(cudaMalloc((void**)&g_peer_err, sizeof(int)));
        GS_CUDA(cudaMemset(g_peer_err, 0, sizeof(int)));
        GS_CUDA(cudaMalloc((void**)&g_peer_done, sizeof(unsigned int)));
        GS_CUDA(cudaMemset(g_peer_done, 0, sizeof(unsigned int)));
    }
    PeerArgs A;
    for (int p = 0; p < kPeerMax; ++p) A.buf[p] = p < world ? reinterpret_cast<float*>(peer_ptrs_host[p]) : nullptr;
    A.world = world; A.rank = rank; A.n = n; A.flag_word = n + 32; A.epoch = epoch; A.scale = scale;
    A.out = out_dev; A.sumsq = sumsq_dev; A.done = g_peer_done; A.err = g_peer_err;
    if (sumsq_dev) GS_CUDA(cudaMemsetAsync(sumsq_dev, 0, sizeof(float), s));
    // every block must be resident at once (the last one waits for the peers): at most two per SM
    const int grid = (int)std::min<int64_t>(ceil_div(n / 4, 256), 2 * (int64_t)sm_count());
    peer_allreduce_kernel<<<grid, 256, 0, s>>>(A);
    GS_LAUNCHED();
    return GSAGE_OK;
}

}  // extern "C"
