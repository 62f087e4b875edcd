// palette_ops.cu — palette backend ops for B200 (sm_100a): RGB <-> HSV kernels and the host-side weighted RGB
// histogram. Replaces palette/src/palette.cu:46-149 and palette/src/bindings.cpp:40-91 of the reference.
// H in degrees [0,360), S and V in [0,100] (the reference's convention, not OpenCV's).
#include <cmath>

#include "hsv_common.cuh"

namespace pnerf {

__global__ void __launch_bounds__(256) k_rgb_to_hsv(uint32_t n, const float* __restrict__ input,
                                                    float* __restrict__ output) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float h, sat, v;
    rgb_to_hsv_dev(input[(size_t)i * 3], input[(size_t)i * 3 + 1], input[(size_t)i * 3 + 2], h, sat, v);
    output[(size_t)i * 3] = h;
    output[(size_t)i * 3 + 1] = sat;
    output[(size_t)i * 3 + 2] = v;
}

__global__ void __launch_bounds__(256) k_hsv_to_rgb(uint32_t n, const float* __restrict__ input,
                                                    float* __restrict__ output) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r, g, b;
    hsv_to_rgb_dev(input[(size_t)i * 3], input[(size_t)i * 3 + 1], input[(size_t)i * 3 + 2], r, g, b);
    output[(size_t)i * 3] = r;
    output[(size_t)i * 3 + 1] = g;
    output[(size_t)i * 3 + 2] = b;
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
