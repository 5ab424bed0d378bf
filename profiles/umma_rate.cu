// umma_rate.cu -- micro-benchmark: how many cycles does one tcgen05.mma (cta_group::1, SS operands, SWIZZLE_128B K-major) occupy?
//
// One elected thread per CTA issues `iters` x 4 MMAs (K = 16 bf16 each, one 64-element k-chunk) back to back into one accumulator and
// commits once; cycles from the first issue to the commit's arrival / MMAs = the sustained per-instruction cost.  One CTA per SM
// (148), so the number is what a persistent kernel sees, not a single SM running alone.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/build/umma_rate profiles/umma_rate.cu && profiles/build/umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../pytorch_graphsage_b200/csrc/umma_ptx.cuh"
using namespace gsage;

template <bool TF32>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int N, int iters, int n_a, int commit_every, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (n_a * 16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;   // zero operands: no NaN slow paths
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 0 && elect_one()) {
        const uint32_t fmt = TF32 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t desc_hi = umma_desc(0);
        const uint32_t a16 = (smem_u32(smem) & 0x3FFFF) >> 4, b16 = a16 + (uint32_t)(n_a * 16384 >> 4);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            // A walks over n_a different 16 KB blocks (the pool kernel: four W blocks against one tile of rows), B stays
            const uint64_t adesc = desc_hi | (uint64_t)(a16 + (uint32_t)(i % n_a) * 1024u), bdesc = desc_hi | (uint64_t)b16;
            const uint32_t d = tmem + (uint32_t)((i % n_a) * N) % 512u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (TF32) umma_tf32(d, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                else umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
            }
            if (commit_every > 0 && (i + 1) % commit_every == 0) umma_commit(smem_u32(&bar2));   // (nobody waits on it: the cost of the commit itself)
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0, nullptr);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <bool TF32>
static void run(int N, int n_a, long long* d_cycles, int commit_every = 0) {
    const int iters = 4000;
    const size_t smem = (size_t)n_a * 16384 + 32768 + 1024;
    cudaFuncSetAttribute(umma_rate_kernel<TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_rate_kernel<TF32><<<148, 128, smem>>>(N, iters, n_a, commit_every, d_cycles);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
    long long c[148];
    cudaMemcpy(c, d_cycles, sizeof(c), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = c[i] > mx ? c[i] : mx;
    const double per = (double)mx / (iters * 4.0);
    const double macs = 128.0 * N * (TF32 ? 8 : 16);
    if (commit_every) printf("[tcgen05.commit after every %d MMAs] ", commit_every * 4);
    printf("%s M=128 N=%-3d K=%-2d, %d A blocks: %6.1f cycles per tcgen05.mma = %5.0f MAC/clk/SM = %6.0f TFLOP/s on 148 SMs at 1.93 GHz\n", TF32 ? "tf32" : "bf16", N,
           TF32 ? 8 : 16, n_a, per, macs / per, macs / per * 2 * 148 * 1.93e9 / 1e12);
}

int main() {
    long long* d_cycles;
    cudaMalloc(&d_cycles, 148 * sizeof(long long));
    for (int n_a = 1; n_a <= 4; n_a *= 4)
        for (int N = 32; N <= 256; N *= 2) run<false>(N, n_a, d_cycles);
    for (int N = 64; N <= 256; N *= 2) run<true>(N, 4, d_cycles);
    for (int ce = 1; ce <= 4; ce *= 2) { run<false>(64, 4, d_cycles, ce); run<false>(128, 4, d_cycles, ce); }
    return 0;
}
