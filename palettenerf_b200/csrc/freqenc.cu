// freqenc.cu — NeRF sinusoidal encoder for B200 (sm_100a). Replaces freqencoder/src/freqencoder.cu:30-129.
// Output row = [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] per input dim block of D, cos written as
// sin(. + pi/2) with the fast __sinf like the reference. One thread per (point, input dim) computes all
// 2*deg outputs from one load (the reference launches one thread per *output* element and re-reads x).
#include "common.cuh"

namespace pnerf {

__global__ void __launch_bounds__(256) k_freq_fwd(const float* __restrict__ inputs, uint32_t B, uint32_t D, uint32_t deg,
                                                  uint32_t C, float* __restrict__ outputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float x = inputs[t];
    float* o = outputs + (size_t)b * C;
    o[d] = x;
    const float half_pi = 3.141592653589793f / 2;
    for (uint32_t f = 0; f < deg; f++) {
        const float xs = scalbnf(x, (int)f);
        o[D + (2 * f) * D + d] = __sinf(xs + 0.0f);
        o[D + (2 * f + 1) * D + d] = __sinf(xs + half_pi);
    }
}

// ref: freqencoder.cu:63-94 — uses the saved outputs: d/dx sin = cos, d/dx cos = -sin, scaled by 2^f
__global__ void __launch_bounds__(256) k_freq_bwd(const float* __restrict__ grad, const float* __restrict__ outputs,
                                                  uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                                  float* __restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float* g = grad + (size_t)b * C;
    const float* o = outputs + (size_t)b * C;
    float result = g[d];
    for (uint32_t f = 0; f < deg; f++) {
        const uint32_t is = D + (2 * f) * D + d, ic = is + D;
        result += scalbnf(1.0f, (int)f) * (g[is] * o[ic] - g[ic] * o[is]);
    }
    grad_inputs[t] = result;
}

}  // namespace pnerf

using namespace pnerf;

extern "C" {

int pnerf_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* outputs,
                              void* stream) {
    if (B == 0 || D == 0) return PNERF_OK;
    PNERF_REQUIRE(inputs && outputs);
    PNERF_REQUIRE(C == D + 2 * D * deg);
    k_freq_fwd<<<ceil_div(B * D, 256u), 256, 0, (cudaStream_t)stream>>>(inputs, B, D, deg, C, outputs);
    return check_launch("freq_encode_forward");
}

int pnerf_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg,
                               uint32_t C, float* grad_inputs, void* stream) {
    if (B == 0 || D == 0) return PNERF_OK;
    PNERF_REQUIRE(grad && outputs && grad_inputs);
    PNERF_REQUIRE(C == D + 2 * D * deg);
    k_freq_bwd<<<ceil_div(B * D, 256u), 256, 0, (cudaStream_t)stream>>>(grad, outputs, B, D, deg, C, grad_inputs);
    return check_launch("freq_encode_backward");
}

}  // extern "C"
