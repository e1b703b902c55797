// common.cuh -- error plumbing and small device helpers shared by the libgmr kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "gmr.h"

namespace gmr {

void set_error(const char* fmt, ...);

#define GMR_CHECK_CUDA(expr)                                                                      \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            gmr::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return GMR_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define GMR_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            gmr::set_error(__VA_ARGS__);  \
            return GMR_ERR_INVALID;       \
        }                                 \
    } while (0)

#define GMR_LAUNCH_CHECK()                                                                        \
    do {                                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            gmr::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return GMR_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

int sm_count();

// ---- streaming / read-only loads -------------------------------------------------------------
// L2 eviction priorities go through createpolicy + .L2::cache_hint (the bare .L2::evict_* qualifiers
// exist only for 256-bit loads on sm_100a).

__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// CSR arrays are read exactly once: keep them out of L1 and mark them first to leave L2, so that
// the gathered embedding rows (which ARE reused) keep the cache.
__device__ __forceinline__ int ld_stream_s32(const int* p, uint64_t pol)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_stream_f32(const float* p, uint64_t pol)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
// gathered rows: read-only path, allocate in L1, evict_last in L2 (they are the reused operand)
__device__ __forceinline__ float4 ld_gather_f4(const float4* p, uint64_t pol)
{
    float4 v;
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_gather_f1(const float* p, uint64_t pol)
{
    float v;
    asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}

}  // namespace gmr
