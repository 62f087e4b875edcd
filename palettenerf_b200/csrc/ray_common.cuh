// ray_common.cuh — per-ray helpers shared by the near/far kernel (raymarch.cu) and the ray generator (rays.cu).
#pragma once
#include "common.cuh"

namespace pnerf {

// ray / AABB slab test (ref: raymarching.cu:111-147): a miss yields near = far = FLT_MAX, a hit clamps near to min_near.
__device__ __forceinline__ void slab_near_far(float ox, float oy, float oz, float dx, float dy, float dz,
                                              const float* __restrict__ aabb, float min_near, float& near_out,
                                              float& far_out) {
    const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
    const float flt_max = 3.402823466e+38f;
    float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx;
    if (near > far) { float s = near; near = far; far = s; }
    float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
    if (near_y > far_y) { float s = near_y; near_y = far_y; far_y = s; }
    bool miss = (near > far_y) || (near_y > far);
    if (!miss) {
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
        if (near_z > far_z) { float s = near_z; near_z = far_z; far_z = s; }
        miss = (near > far_z) || (near_z > far);
        if (!miss) {
            if (near_z > near) near = near_z;
            if (far_z < far) far = far_z;
            if (near < min_near) near = min_near;
        }
    }
    near_out = miss ? flt_max : near;
    far_out = miss ? flt_max : far;
}

}  // namespace pnerf
