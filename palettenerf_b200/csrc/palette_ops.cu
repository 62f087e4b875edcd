// palette_ops.cu — palette backend ops for B200 (sm_100a): RGB <-> HSV kernels and the host-side weighted RGB
// histogram. Replaces palette/src/palette.cu:46-149 and palette/src/bindings.cpp:40-91 of the reference.
// H in degrees [0,360), S and V in [0,100] (the reference's convention, not OpenCV's).
#include <cmath>

#include "common.cuh"

namespace pnerf {

__device__ __forceinline__ bool near_eq(float a, float b) { return fabsf(a - b) < 1e-9f; }

__global__ void __launch_bounds__(256) k_rgb_to_hsv(uint32_t n, const float* __restrict__ input,
                                                    float* __restrict__ output) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float r = input[(size_t)i * 3], g = input[(size_t)i * 3 + 1], b = input[(size_t)i * 3 + 2];
    const float cmax = fmaxf(fmaxf(r, g), b), cmin = fminf(fminf(r, g), b);
    const float diff = cmax - cmin;
    float h;
    if (near_eq(diff, 0.f)) h = 0.f;
    else if (near_eq(cmax, r)) h = fmodf(60 * ((g - b) / diff) + 360, 360.f);
    else if (near_eq(cmax, g)) h = fmodf(60 * ((b - r) / diff) + 120, 360.f);
    else h = fmodf(60 * ((r - g) / diff) + 240, 360.f);
    const float s = near_eq(cmax, 0.f) ? 0.f : (diff / cmax) * 100;
    output[(size_t)i * 3] = h;
    output[(size_t)i * 3 + 1] = s;
    output[(size_t)i * 3 + 2] = cmax * 100;
}

__global__ void __launch_bounds__(256) k_hsv_to_rgb(uint32_t n, const float* __restrict__ input,
                                                    float* __restrict__ output) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float h = input[(size_t)i * 3], s = input[(size_t)i * 3 + 1], v = input[(size_t)i * 3 + 2];
    const float c = s / 100 * v / 100;
    const float x = c * (1 - fabsf(fmodf(h / 60, 2.f) - 1));
    const float m = v / 100 - c;
    float r = 0, g = 0, b = 0;
    if (h >= 0 && h < 60) { r = c; g = x; }
    else if (h >= 60 && h < 120) { r = x; g = c; }
    else if (h >= 120 && h < 180) { g = c; b = x; }
    else if (h >= 180 && h < 240) { g = x; b = c; }
    else if (h >= 240 && h < 300) { r = x; b = c; }
    else { r = c; b = x; }
    output[(size_t)i * 3] = r + m;
    output[(size_t)i * 3 + 1] = g + m;
    output[(size_t)i * 3 + 2] = b + m;
}

}  // namespace pnerf

using namespace pnerf;

extern "C" {

int pnerf_rgb_to_hsv(uint32_t n, const float* input, float* output, void* stream) {
    if (n == 0) return PNERF_OK;
    PNERF_REQUIRE(input && output);
    k_rgb_to_hsv<<<ceil_div(n, 256u), 256, 0, (cudaStream_t)stream>>>(n, input, output);
    return check_launch("rgb_to_hsv");
}

int pnerf_hsv_to_rgb(uint32_t n, const float* input, float* output, void* stream) {
    if (n == 0) return PNERF_OK;
    PNERF_REQUIRE(input && output);
    k_hsv_to_rgb<<<ceil_div(n, 256u), 256, 0, (cudaStream_t)stream>>>(n, input, output);
    return check_launch("hsv_to_rgb");
}

// Host-only (the reference's is too): weighted histogram over a (2^b)^3 RGB cube, r most significant.
int pnerf_compute_rgb_histogram(const float* colors_rgb, const float* weights, uint64_t n, int bits_per_channel,
                                double* bin_weights, float* bin_centers_rgb) {
    PNERF_REQUIRE(colors_rgb && weights && bin_weights && bin_centers_rgb);
    PNERF_REQUIRE(bits_per_channel >= 1 && bits_per_channel <= 8);
    const int bpc = bits_per_channel;
    const uint32_t side = 1u << bpc, num_bins = 1u << (3 * bpc);
    for (uint32_t i = 0; i < num_bins; i++) bin_weights[i] = 0.0;
    for (uint64_t i = 0; i < n; i++) {
        uint32_t bin = 0;
        for (int ch = 0; ch < 3; ch++) {
            const float c = std::fmax(0.0f, std::fmin(0.999f, colors_rgb[i * 3 + ch]));
            bin = (bin << bpc) + (uint32_t)(c * (float)side);
        }
        bin_weights[bin] += (double)weights[i];
    }
    for (uint32_t bin = 0; bin < num_bins; bin++) {
        uint32_t code = bin;
        for (int ch = 2; ch >= 0; ch--) {
            bin_centers_rgb[bin * 3 + ch] = ((float)(code & (side - 1)) + 0.5f) / (float)side;
            code >>= bpc;
        }
    }
    return PNERF_OK;
}

}  // extern "C"
