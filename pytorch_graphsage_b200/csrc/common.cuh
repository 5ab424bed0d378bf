// common.cuh -- shared host/device helpers for libgsage_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <string>

#include "../../include/gsage_b200.h"

namespace gsage {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define GS_CHECK_ARG(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            ::gsage::set_error(__VA_ARGS__);                      \
            return GSAGE_ERR_INVALID;                             \
        }                                                         \
    } while (0)

#define GS_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t err__ = (call);                                                                \
        if (err__ != cudaSuccess) {                                                                \
            ::gsage::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                               cudaGetErrorString(err__));                                         \
            return GSAGE_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

// every kernel launch goes through this so bench.py can report gpu_launches honestly
#define GS_LAUNCHED()                                                                              \
    do {                                                                                           \
        ::gsage::g_launches.fetch_add(1, std::memory_order_relaxed);                               \
        cudaError_t err__ = cudaGetLastError();                                                    \
        if (err__ != cudaSuccess) {                                                                \
            ::gsage::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,            \
                               cudaGetErrorString(err__));                                         \
            return GSAGE_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

#define GS_TRY(call)                  \
    do {                              \
        int st__ = (call);            \
        if (st__ != GSAGE_OK) return st__; \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t dtype_size(int dt) { return dt == GSAGE_BF16 ? 2 : 4; }

int sm_count();   // cached cudaDevAttrMultiProcessorCount of the current device

// ---- device helpers -----------------------------------------------------------------------------
#ifdef __CUDACC__

// 16-byte read-only load that does not pollute L1 (rows of the feature table are touched once per CTA)
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// L2 eviction-priority policies (createpolicy): streaming table rows should leave L2 first, the small reduced-row
// buffer that the projection kernel re-reads right away should stay
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 ldg_nc_v4_hint(const void* p, uint64_t policy) {
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
        : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ void stg_v4_hint(void* p, const uint4& v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy) : "memory");
}

__device__ __forceinline__ void st_shared_v4(uint32_t smem_addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <typename T> struct ElemTraits;
template <> struct ElemTraits<float> {
    static constexpr int kPerVec = 4;    // elements per 16-byte vector
    __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
        f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
        f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    }
    __device__ static __forceinline__ uint4 pack(const float* f) {
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
    __device__ static __forceinline__ float load(const float* p) { return *p; }
    __device__ static __forceinline__ void store(float* p, float v) { *p = v; }
};
template <> struct ElemTraits<__nv_bfloat16> {
    static constexpr int kPerVec = 8;
    __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
        f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
        f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
    }
    __device__ static __forceinline__ uint4 pack(const float* f) {
        return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
    }
    __device__ static __forceinline__ float load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    __device__ static __forceinline__ void store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == GSAGE_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == GSAGE_ACT_TANH) return tanhf(v);
    return v;
}

#endif  // __CUDACC__

}  // namespace gsage
