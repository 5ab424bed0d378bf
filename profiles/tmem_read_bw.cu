// tmem_read_bw.cu -- micro-benchmark: how fast can the warps of one SM read TMEM (tcgen05.ld)?
//
// The pool kernel (linear_pool_ws_umma.cu) reads every fp32 accumulator element exactly once (the max / mean over a parent's S
// columns happens in registers), so its floor is the TMEM read rate, not the tensor pipe.  This prints bytes / clock / SM for
// 4 .. 16 reading warps and two instruction shapes, with and without a dependent-use wait after every load.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/build/tmem_read_bw profiles/tmem_read_bw.cu && profiles/build/tmem_read_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* r);

template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}

template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}

// mode 0: two loads, one wait (what the pool epilogue does); mode 1: keep one load in flight while "using" the previous one
template <int X, int MODE>
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    uint32_t a[X], b[X], acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (MODE == 0) {
        for (int i = 0; i < iters; ++i) {
            ld<X>(base + ((i & 1) ? 256u : 0u), a);
            ld<X>(base + ((i & 1) ? 256u : 0u) + X, b);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int k = 0; k < X; ++k) acc = max(acc, max(a[k], b[k]));
        }
    } else {
        ld<X>(base, a);
        for (int i = 0; i < iters; ++i) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            ld<X>(base + X, b);
#pragma unroll
            for (int k = 0; k < X; ++k) acc = max(acc, a[k]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            ld<X>(base + ((i & 1) ? 0u : 256u), a);
#pragma unroll
            for (int k = 0; k < X; ++k) acc = max(acc, b[k]);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int X, int MODE>
static void run(int warps, int iters, long long* d_cycles, uint32_t* d_sink) {
    tmem_read_kernel<X, MODE><<<148, 32 * warps>>>(iters, d_cycles, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
    long long c[148];
    cudaMemcpy(c, d_cycles, sizeof(c), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = c[i] > mx ? c[i] : mx;
    const double bytes = (double)iters * 2.0 * X * 32.0 * 4.0 * warps;
    printf("x%-2d mode %d  %2d warps: %9lld cycles for %d iterations -> %7.1f B/clk/SM  (%.0f cycles per 64-column x 16-warp tile)\n", X, MODE, warps, mx,
           iters, bytes / (double)mx, 131072.0 / (bytes / (double)mx));
}

int main() {
    long long* d_cycles;
    uint32_t* d_sink;
    cudaMalloc(&d_cycles, 148 * sizeof(long long));
    cudaMalloc(&d_sink, 2048);
    const int iters = 4000;
    for (int warps = 4; warps <= 16; warps *= 2) {
        run<32, 0>(warps, iters, d_cycles, d_sink);
        run<32, 1>(warps, iters, d_cycles, d_sink);
        run<16, 0>(warps, iters, d_cycles, d_sink);
        run<16, 1>(warps, iters, d_cycles, d_sink);
    }
    return 0;
}
