// field_cache.cu — the derived fp16 state the inference kernels read, rebuilt from the fp32 parameters in TWO launches:
//   pnerf_field_cache_tables  the hash tables: fp32 masters [n, 2] of the density and palette grids -> ONE interleaved fp16
//                             table [n][2 grids][2] (an 8-byte entry serves both grids' gather), the semantic grid -> its own
//                             fp16 table. An HBM stream: 8 + 8 (+ 8) bytes read, 8 (+ 4) written per entry.
//   pnerf_field_cache_gather  every small tensor (MLP weight images in mma.sync and tcgen05 order, head bias, palette):
//                             out[i] = *src[i] through a table of source ADDRESSES built once by the host (the parameters are
//                             ~15 separate tensors; the address table is rebuilt only when one of them moves).
//   pnerf_field_cache_merge   the two layers of the tcgen05 weight image that are products of reference layers without an
//                             activation between them (field_tc.cuh layer table): fp32 products of the fp32 parameters, rounded
//                             to fp16 once, written in the image's [k-chunk][n][8 halfs] order. 78 k MACs, one small launch.
// Before: 3 strided torch copies + cat + 2 index + ~6 small copies = ~12 launches, 93 us per view; the reference has no
// counterpart (its MLPs read the fp32 parameters through autocast on every call, palette/network.py:156-280).
#include <cuda_fp16.h>

#include "common.cuh"
#include "field_tc.cuh"

namespace pnerf {

__global__ void __launch_bounds__(256) k_cache_tables(const float2* __restrict__ t_sigma, const float2* __restrict__ t_palette,
                                                      const float2* __restrict__ t_clip, uint32_t n, uint2* __restrict__ pair,
                                                      uint32_t* __restrict__ clip) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float2 a = __ldg(t_sigma + i), b = __ldg(t_palette + i);
        const __half2 ha = __floats2half2_rn(a.x, a.y), hb = __floats2half2_rn(b.x, b.y);
        pair[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&ha), *reinterpret_cast<const uint32_t*>(&hb));
        if (t_clip) {
            const float2 c = __ldg(t_clip + i);
            const __half2 hc = __floats2half2_rn(c.x, c.y);
            clip[i] = *reinterpret_cast<const uint32_t*>(&hc);
        }
    }
}

// src [n16 + n32] device addresses of fp32 scalars (0 = the value 0); the first n16 go to out16 as fp16, the rest to out32 as
// fp32, those at or behind clamp_from clamped to [0, 1] (the palette: basis_color.clamp(0, 1), palette/renderer.py:322)
__global__ void __launch_bounds__(256) k_cache_gather(const unsigned long long* __restrict__ src, uint32_t n16, uint32_t n32,
                                                      uint32_t clamp_from, __half* __restrict__ out16, float* __restrict__ out32) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16 + n32) return;
    const unsigned long long a = src[i];
    float v = a ? *reinterpret_cast<const float*>(a) : 0.f;
    if (i < n16) {
        out16[i] = __float2half_rn(v);
    } else {
        if (i - n16 >= clamp_from) v = fminf(fmaxf(v, 0.f), 1.f);
        out32[i - n16] = v;
    }
}

// one thread per element of TD0 (64 x 64) and TB1 (32 x 64) of the tcgen05 weight image
__global__ void __launch_bounds__(256) k_cache_merge(const float* __restrict__ s1 /*[16,64]*/, const float* __restrict__ d0 /*[64,15]*/,
                                                     const float* __restrict__ b1 /*[15,64]*/, const float* __restrict__ orw /*[13,15]*/,
                                                     const float* __restrict__ omw /*[4,15]*/, __half* __restrict__ image) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 64 * 64) {
        const int n = i >> 6, k = i & 63;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 15; j++) acc = fmaf(__ldg(d0 + n * 15 + j), __ldg(s1 + (1 + j) * 64 + k), acc);
        image[tc_woff(TD0) / 2 + ((k >> 3) * 64 + n) * 8 + (k & 7)] = __float2half_rn(acc);
    } else if (i < 64 * 64 + 32 * 64) {
        const int e = i - 64 * 64, n = e >> 6, k = e & 63;
        float acc = 0.f;
        if (n < 17) {
            const float* w = n < 13 ? orw + n * 15 : omw + (n - 13) * 15;
#pragma unroll
            for (int j = 0; j < 15; j++) acc = fmaf(__ldg(w + j), __ldg(b1 + j * 64 + k), acc);
        }
        image[tc_woff(TB1) / 2 + ((k >> 3) * 32 + n) * 8 + (k & 7)] = __float2half_rn(acc);
    }
}

}  // namespace pnerf

using namespace pnerf;

extern "C" {

int pnerf_field_cache_tables(const float* table_sigma, const float* table_palette, const float* table_clip, uint32_t n_entries,
                             void* pair, void* clip, void* stream) {
    if (!table_sigma || !table_palette || !pair || (table_clip && !clip)) return PNERF_ERR_INVALID_ARG;
    if (n_entries == 0) return PNERF_OK;
    const uint32_t grid = min(ceil_div(n_entries, 256u), 148u * 16u);   // grid-stride: 2 waves of 8 CTAs per SM
    k_cache_tables<<<grid, 256, 0, (cudaStream_t)stream>>>((const float2*)table_sigma, (const float2*)table_palette,
                                                          (const float2*)table_clip, n_entries, (uint2*)pair, (uint32_t*)clip);
    return check_launch("pnerf_field_cache_tables");
}

int pnerf_field_cache_gather(const uint64_t* src, uint32_t n16, uint32_t n32, uint32_t clamp_from, void* out16, float* out32,
                             void* stream) {
    if (!src || (n16 && !out16) || (n32 && !out32)) return PNERF_ERR_INVALID_ARG;
    if (n16 + n32 == 0) return PNERF_OK;
    k_cache_gather<<<ceil_div(n16 + n32, 256u), 256, 0, (cudaStream_t)stream>>>((const unsigned long long*)src, n16, n32,
                                                                                 clamp_from, (__half*)out16, out32);
    return check_launch("pnerf_field_cache_gather");
}

int pnerf_field_cache_merge(const float* sigma1_w, const float* diff0_w, const float* basis1_w, const float* offrad_w,
                            const float* omega_w, void* wimage_tc, void* stream) {
    if (!sigma1_w || !diff0_w || !basis1_w || !offrad_w || !omega_w || !wimage_tc) return PNERF_ERR_INVALID_ARG;
    k_cache_merge<<<(64 * 64 + 32 * 64) / 256, 256, 0, (cudaStream_t)stream>>>(sigma1_w, diff0_w, basis1_w, offrad_w, omega_w,
                                                                              (__half*)wimage_tc);
    return check_launch("pnerf_field_cache_merge");
}

}  // extern "C"
