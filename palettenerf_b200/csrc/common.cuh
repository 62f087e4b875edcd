// common.cuh — shared helpers for the pnerf_b200 kernels (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pnerf_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "pnerf_b200 kernels are written for sm_100a (B200) only"
#endif

namespace pnerf {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_last_cuda_error(cudaError_t e, const char* where);

inline int check_launch(const char* where) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_last_cuda_error(e, where);
        return PNERF_ERR_CUDA;
    }
    return PNERF_OK;
}

template <typename T>
__host__ __device__ __forceinline__ T ceil_div(T a, T b) {
    return (a + b - 1) / b;
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

// streaming (read-once / write-once) accesses: keep them out of L1 so the occupancy bitfield and the
// hash tables keep the cache.
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace pnerf

#define PNERF_REQUIRE(cond)                        \
    do {                                           \
        if (!(cond)) return PNERF_ERR_INVALID_ARG; \
    } while (0)
