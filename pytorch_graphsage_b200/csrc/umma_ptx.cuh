// umma_ptx.cuh -- inline-PTX building blocks shared by the tcgen05 kernels (linear_umma.cu, linear_pool_umma.cu):
// mbarrier, TMA (tile + tile::gather4), tcgen05.mma / commit / ld, UMMA smem descriptors, tensor-map encoding.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace gsage {

// ---- PTX helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a CONVERGED warp (all 32 lanes must reach this together).  Guard single-thread issue loops (tcgen05.mma, TMA,
// tcgen05.commit) with this instead of `lane == 0`: those instructions take uniform-register operands, and inside an
// `if (lane == 0)` region ptxas cannot prove that a single lane is active, so it wraps EVERY one of them in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop over the active lanes (~60-160 cycles of dependent issue per instruction, measured:
// profiles/r02_pool_phase_cycles.txt).  Under elect.sync it emits them back to back.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xFFFFFFFF;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a pipeline bug must not hang the box
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {       // ~2 s at 1.9 GHz
            if (err) atomicExch(err, 1);
            __trap();
        }
    }
}
// the same for a waiter that is far AHEAD of what it waits for (a producer waiting for a free buffer): back off between polls,
// so that the poll loop does not take issue slots from the warps doing the work (ncu on the attention kernel: 32 M polls, 90 % of
// all executed instructions, from twelve producer lanes)
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, int* err) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        __nanosleep(256);
        if (clock64() - t0 > 4000000000LL) {
            if (err) atomicExch(err, 1);
            __trap();
        }
    }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the thread's arrival on the mbarrier is performed by the hardware once every cp.async it has issued so far has
// landed (.noinc: it counts as one of the expected arrivals) -- the completion mechanism of CUTLASS' sm100
// cp.async->UMMA mainloop: no proxy fence, no blocking wait in the producer
__device__ __forceinline__ void cp_async_arrive_on(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// one 2-D tile (box of the descriptor) starting at (col, row) -> swizzled smem tile; bytes are credited to `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int col, int row, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(col), "r"(row), "r"(bar) : "memory");
}
// four rows of the table, picked by id, 64 columns each -> four consecutive 128-byte rows of the swizzled tile
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
// pull `bytes` (multiple of 16) starting at p into L2: used to fetch WHOLE rows of the next tile from DRAM in one
// contiguous burst each, so the 128-byte-per-row tile loads that follow hit L2 instead of re-opening DRAM pages
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// LSU-path L2 prefetch of one 128-byte line (fire and forget: no register, no scoreboard)
__device__ __forceinline__ void prefetch_l2_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// `nthreads` threads (this one is `t`) pull `rows` whole rows of an operand tile into L2: row r of the tile is
// a[ids ? ids[row0 + r] : row0 + r] (pitch bytes apart), `bytes` long.  Consecutive threads take consecutive lines of the
// same row, so a row's DRAM page is opened once.  The TMA unit keeps only a few KB of requests in flight per SM
// (measured: ~16 KB per microsecond-latency round trip), so the tile loads of the projection kernels run at DRAM
// *latency* unless the lines are already in L2 when TMA asks for them -- this is what puts them there.
__device__ __forceinline__ void prefetch_tile_rows(const void* a, int64_t pitch, const int64_t* ids, int64_t row0, int64_t n,
                                                   int rows, int bytes, int t, int nthreads) {
    const int lines = (bytes + 127) / 128 + 1;             // + 1: rows are not 128-byte aligned, the last line covers the tail
    for (int i = t; i < rows * lines; i += nthreads) {
        const int r = i / lines, l = i - r * lines;
        const int64_t row = row0 + r;
        if (row >= n) continue;
        const int64_t src = ids ? __ldg(ids + row) : row;
        int off = l * 128;
        if (off >= bytes) off = bytes - 1;
        prefetch_l2_line((const char*)a + src * pitch + off);
    }
}
// pacing of the prefetch warps: the MMA warp publishes how many tiles it has issued; a prefetch warp waits (sleeping)
// until `want` tiles have been issued.  Approximate on purpose (issue, not completion) -- it only has to keep the
// prefetched lines within L2's reach of their use.
__device__ __forceinline__ void progress_publish(volatile int* p, int v) { *p = v; }
__device__ __forceinline__ void progress_wait(volatile int* p, int want) {
    while (*p < want) __nanosleep(200);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows 128 B apart, 8-row atoms 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, 16-byte units
    d |= (uint64_t)1 << 16;                                // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// ---- host: TMA descriptors ---------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}

// bf16 (rows, cols) row-major with `ld` elements between rows; box = 64 columns x box_rows, 128-byte swizzle
static int make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int es) {
    PFN_cuTensorMapEncodeTiled_v12000 enc = tensor_map_encoder();
    GS_CHECK_ARG(enc, "linear_umma: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    const cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(m, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GS_CHECK_ARG(r == CUDA_SUCCESS, "linear_umma: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return GSAGE_OK;
}


// one 32-byte store (a whole DRAM sector per thread; st.global.v8.b32 needs sm_100+)
__device__ __forceinline__ void st_global_v8(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g, uint32_t h) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h) : "memory");
}

// The projection kernels' epilogue for one thread: 32 accumulator columns of ONE output row (TMEM lane = row) -> + bias -> ACT -> HBM.
// A thread owns a row segment, so every warp-wide store touches 32 different lines and the LSU takes them one by one: the
// stores, not the MMAs, set the pace of the weight-stationary kernel whenever d <= 256 (cycle counters:
// profiles/r02_linear_ws_phase_cycles.txt).  32-byte stores halve the number of requests (bf16, d = 64: 186 -> 124 us); staging
// the chunk through shared memory to store along the rows was slower on every shape (the tensor core's operand reads already
// use the shared-memory bandwidth: 8 KB per 67-cycle MMA).
template <int ACT>
__device__ __forceinline__ void epilogue_store32(const uint32_t* r, const float* bias, int valid, void* out, int out_bf16) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j < valid) v[j] += __ldg(bias + j);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (ACT == GSAGE_ACT_RELU) v[j] = fmaxf(v[j], 0.0f);
        if (ACT == GSAGE_ACT_TANH) v[j] = tanhf(v[j]);
    }
    const bool vec = valid == 32 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const bool vec32 = vec && ((reinterpret_cast<uintptr_t>(out) & 31) == 0);
    if (out_bf16) {
        if (vec32) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
                st_global_v8(reinterpret_cast<char*>(out) + 32 * q, pack_bf16(v[16 * q], v[16 * q + 1]), pack_bf16(v[16 * q + 2], v[16 * q + 3]),
                             pack_bf16(v[16 * q + 4], v[16 * q + 5]), pack_bf16(v[16 * q + 6], v[16 * q + 7]), pack_bf16(v[16 * q + 8], v[16 * q + 9]),
                             pack_bf16(v[16 * q + 10], v[16 * q + 11]), pack_bf16(v[16 * q + 12], v[16 * q + 13]), pack_bf16(v[16 * q + 14], v[16 * q + 15]));
        } else if (vec) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                reinterpret_cast<uint4*>(out)[q] = make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                                                              pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < valid) reinterpret_cast<__nv_bfloat16*>(out)[j] = __float2bfloat16_rn(v[j]);
        }
    } else {
        if (vec32) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                st_global_v8(reinterpret_cast<char*>(out) + 32 * q, __float_as_uint(v[8 * q]), __float_as_uint(v[8 * q + 1]), __float_as_uint(v[8 * q + 2]),
                             __float_as_uint(v[8 * q + 3]), __float_as_uint(v[8 * q + 4]), __float_as_uint(v[8 * q + 5]), __float_as_uint(v[8 * q + 6]),
                             __float_as_uint(v[8 * q + 7]));
        } else if (vec) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                reinterpret_cast<float4*>(out)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < valid) reinterpret_cast<float*>(out)[j] = v[j];
        }
    }
}

}  // namespace gsage
