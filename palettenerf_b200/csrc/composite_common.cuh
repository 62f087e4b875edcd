// composite_common.cuh — warp scans and the per-chunk transmittance bookkeeping shared by the training compositors
// (composite.cu, composite_palette.cu).
#pragma once
#include "common.cuh"

namespace pnerf {

__device__ __forceinline__ float warp_scan_mul(float v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v *= u;
    }
    return v;
}
__device__ __forceinline__ float warp_scan_add(float v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v += u;
    }
    return v;
}

// Per-chunk transmittance bookkeeping shared by all training kernels.
// in : alpha (0 for lanes past the end), valid, running T (warp-uniform)
// out: T_before / T_after for this lane's sample, `last` = lane index of the terminating sample in this chunk
//      (32 if the ray does not terminate here). Updates T_carry to the value after lane 31.
struct ChunkT {
    float T_before, T_after;
    uint32_t last;
};
__device__ __forceinline__ ChunkT chunk_transmittance(float alpha, bool valid, float& T_carry, float T_thresh,
                                                      uint32_t lane) {
    ChunkT c;
    const float om = 1.0f - alpha;
    const float p_incl = warp_scan_mul(om, lane);
    float p_excl = __shfl_up_sync(0xffffffffu, p_incl, 1);
    if (lane == 0) p_excl = 1.0f;
    c.T_before = T_carry * p_excl;
    c.T_after = T_carry * p_incl;
    const uint32_t term = __ballot_sync(0xffffffffu, valid && (c.T_after < T_thresh));
    c.last = term ? (uint32_t)(__ffs(term) - 1) : 32u;
    T_carry = __shfl_sync(0xffffffffu, c.T_after, 31);
    return c;
}

}  // namespace pnerf
