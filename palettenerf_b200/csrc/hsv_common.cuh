// hsv_common.cuh — RGB <-> HSV of one colour, shared by the stand-alone kernels (palette_ops.cu) and the RegionEdit epilogue
// of the fused renderer. H in degrees [0,360), S and V in [0,100] (ref: palette/src/palette.cu:58-85, 101-132).
#pragma once
#include "common.cuh"

namespace pnerf {

__device__ __forceinline__ bool near_eq(float a, float b) { return fabsf(a - b) < 1e-9f; }

__device__ __forceinline__ void rgb_to_hsv_dev(float r, float g, float b, float& h, float& s, float& v) {
    const float cmax = fmaxf(fmaxf(r, g), b), cmin = fminf(fminf(r, g), b);
    const float diff = cmax - cmin;
    if (near_eq(diff, 0.f)) h = 0.f;
    else if (near_eq(cmax, r)) h = fmodf(60 * ((g - b) / diff) + 360, 360.f);
    else if (near_eq(cmax, g)) h = fmodf(60 * ((b - r) / diff) + 120, 360.f);
    else h = fmodf(60 * ((r - g) / diff) + 240, 360.f);
    s = near_eq(cmax, 0.f) ? 0.f : (diff / cmax) * 100;
    v = cmax * 100;
}

__device__ __forceinline__ void hsv_to_rgb_dev(float h, float s, float v, float& r, float& g, float& b) {
    const float c = s / 100 * v / 100;
    const float x = c * (1 - fabsf(fmodf(h / 60, 2.f) - 1));
    const float m = v / 100 - c;
    r = 0; g = 0; b = 0;
    if (h >= 0 && h < 60) { r = c; g = x; }
    else if (h >= 60 && h < 120) { r = x; g = c; }
    else if (h >= 120 && h < 180) { g = c; b = x; }
    else if (h >= 180 && h < 240) { g = x; b = c; }
    else if (h >= 240 && h < 300) { r = x; b = c; }
    else { r = c; b = x; }
    r += m; g += m; b += m;
}

}  // namespace pnerf
