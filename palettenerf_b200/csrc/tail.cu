// tail.cu — the per-ray epilogue of run_cuda (ref: palette/renderer.py:399-429 training, :525-551 inference;
// nerf/renderer.py:335-343): depth normalisation and background mixing,
//     depth_n = clamp(depth - near, min = 0) / (far - near)
//     image   = image_raw  + (1 - weights_sum) * bg_color
//     direct  = direct_raw + (1 - weights_sum) * bg_color
// In torch that is ~10 elementwise launches on [N]-sized tensors forward and ~12 backward (each 2-8 us): one launch
// here, and one for the backward (d image_raw = d image, d direct_raw = d direct, d weights_sum = -sum_c bg_c (d image_c
// + d direct_c)). `direct_raw` may be three columns of a wider row-major tensor (the training branch's channel
// composite): it is read in place and its gradient is written as a full [N, stride] tensor (zero elsewhere), so autograd
// adds it to the loss's gradient of that tensor with one launch instead of slicing.
#include "common.cuh"

namespace pnerf {

__global__ void __launch_bounds__(256) k_tail_fwd(uint32_t N, const float* __restrict__ depth, const float* __restrict__ nears,
                                                  const float* __restrict__ fars, const float* __restrict__ image,
                                                  const float* __restrict__ ws, const float* __restrict__ direct,
                                                  uint32_t direct_stride, const float* __restrict__ bg, uint32_t bg_stride,
                                                  float* __restrict__ depth_n, float* __restrict__ image_out,
                                                  float* __restrict__ direct_out) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float w = 1.0f - ws[n];
    if (depth_n) depth_n[n] = fmaxf(depth[n] - nears[n], 0.0f) / (fars[n] - nears[n]);
    const float* b = bg + (size_t)n * bg_stride;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float wb = w * b[c];
        image_out[(size_t)n * 3 + c] = image[(size_t)n * 3 + c] + wb;
        if (direct_out) direct_out[(size_t)n * 3 + c] = direct[(size_t)n * direct_stride + c] + wb;
    }
}

// g_ws [N]; g_direct_full [N, stride] = zeros with g_direct in columns [col, col + 3) (NULL: no direct branch)
__global__ void __launch_bounds__(256) k_tail_bwd(uint32_t N, const float* __restrict__ g_image, const float* __restrict__ g_direct,
                                                  const float* __restrict__ bg, uint32_t bg_stride, uint32_t stride,
                                                  uint32_t col, float* __restrict__ g_ws, float* __restrict__ g_direct_full) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* b = bg + (size_t)n * bg_stride;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float g = g_image ? g_image[(size_t)n * 3 + c] : 0.f;
        if (g_direct) g += g_direct[(size_t)n * 3 + c];
        acc = fmaf(b[c], g, acc);
    }
    g_ws[n] = -acc;
    if (g_direct_full) {
        float* row = g_direct_full + (size_t)n * stride;
        for (uint32_t c = 0; c < stride; c++) row[c] = (c >= col && c < col + 3 && g_direct) ? g_direct[(size_t)n * 3 + (c - col)] : 0.f;
    }
}

}  // namespace pnerf

using namespace pnerf;

extern "C" {

int pnerf_render_tail_forward(uint32_t N, const float* depth, const float* nears, const float* fars, const float* image,
                              const float* weights_sum, const float* direct, uint32_t direct_stride, const float* bg,
                              uint32_t bg_stride, float* depth_n, float* image_out, float* direct_out, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(image && weights_sum && bg && image_out && (bg_stride == 0 || bg_stride == 3));
    PNERF_REQUIRE(depth_n == nullptr || (depth && nears && fars));
    PNERF_REQUIRE(direct_out == nullptr || (direct && direct_stride >= 3));
    k_tail_fwd<<<ceil_div(N, 256u), 256, 0, (cudaStream_t)stream>>>(N, depth, nears, fars, image, weights_sum, direct,
                                                                    direct_stride, bg, bg_stride, depth_n, image_out, direct_out);
    return check_launch("render_tail_forward");
}

int pnerf_render_tail_backward(uint32_t N, const float* g_image, const float* g_direct, const float* bg, uint32_t bg_stride,
                               uint32_t stride, uint32_t col, float* g_weights_sum, float* g_direct_full, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(bg && g_weights_sum && (bg_stride == 0 || bg_stride == 3));
    PNERF_REQUIRE(g_direct_full == nullptr || (stride >= 3 && col + 3 <= stride));
    k_tail_bwd<<<ceil_div(N, 256u), 256, 0, (cudaStream_t)stream>>>(N, g_image, g_direct, bg, bg_stride, stride, col,
                                                                    g_weights_sum, g_direct_full);
    return check_launch("render_tail_backward");
}

}  // extern "C"
