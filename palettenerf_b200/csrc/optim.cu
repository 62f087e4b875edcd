// optim.cu — Adam step for all tensors of a parameter group in ONE launch (second half).
//
// Replaces torch.optim.Adam as the reference uses it (ref: palette/utils.py:719-724 / nerf/utils.py:895-900:
// scaler.step(optimizer) with optim.Adam(params, betas=(0.9, 0.99), eps=1e-15); the hash tables are 12.66 M fp32
// entries each, 50.6 MB of parameters + 2 x 50.6 MB of moments per trained grid). The update is a pure stream:
// read p, g, m, v (16 B/element), write p, m, v (12 B/element) — 28 B per element against HBM. The GradScaler contract
// of torch's fused optimizers is kept: `grad_scale` (device scalar, gradients are divided by it on the fly — no separate
// unscale pass over the gradients) and `found_inf` (device scalar: non-zero -> the whole step is skipped, steps are not
// advanced), so the kernel is CUDA-graph capturable and needs no host synchronisation.
//
// Arithmetic = torch's (aten/src/ATen/native/cuda/fused_adam_utils.cuh, non-amsgrad, maximize = false):
//   g = grad / grad_scale (+ weight_decay * p);  m = lerp(m, g, 1 - beta1);  v = beta2 * v + (1 - beta2) g^2
//   p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps),   t = step + 1
#include "common.cuh"

namespace pnerf {

constexpr int kAdamMaxTensors = PNERF_ADAM_MAX_TENSORS;
constexpr int kAdamThreads = 256;
constexpr int kAdamPerBlock = kAdamThreads * 8;     // elements per CTA (two float4 per thread)

struct AdamArgs {
    pnerf_adam_tensor t[kAdamMaxTensors];
    uint32_t first_block[kAdamMaxTensors + 1];       // CTA range of tensor i: [first_block[i], first_block[i+1])
    uint32_t count;
    float lr, beta1, beta2, eps, weight_decay;
    const float* lr_dev;                              // optional device learning rate (overrides lr)
    const float* grad_scale; const float* found_inf;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float inv_scale, float wd, float beta1,
                                         float beta2, float step_size, float rsqrt_bc2, float eps) {
    g *= inv_scale;
    if (wd != 0.f) g = fmaf(wd, p, g);
    m = m + (1.0f - beta1) * (g - m);                                   // std::lerp(m, g, 1 - beta1), weight < 0.5 branch
    v = beta2 * v + (1.0f - beta2) * g * g;
    const float denom = sqrtf(v) * rsqrt_bc2 + eps;
    p -= step_size * m / denom;
}

__global__ void __launch_bounds__(kAdamThreads) k_adam(const __grid_constant__ AdamArgs a) {
    if (a.found_inf && __ldg(a.found_inf) != 0.f) return;              // GradScaler: skip the step
    uint32_t ti = 0;
#pragma unroll 1
    while (ti + 1 < a.count && blockIdx.x >= a.first_block[ti + 1]) ti++;
    const pnerf_adam_tensor& T = a.t[ti];
    // bias corrections 1 - beta^t = -expm1(t log beta): fp32 expm1f / logf keep the relative error of the correction at a
    // few 1e-7 even for t = 1 (where 1 - 0.99 cancels), without the double-precision pow() torch uses (a long serial
    // prologue in front of every warp's first load; measured: 97 -> 92 us for the 12.7 M-element step)
    const float t = __ldg(T.step) + 1.0f;
    const float lr = a.lr_dev ? __ldg(a.lr_dev) : a.lr;
    const float step_size = lr / (-expm1f(t * logf(a.beta1)));
    const float rsqrt_bc2 = 1.0f / sqrtf(-expm1f(t * logf(a.beta2)));
    const float inv_scale = a.grad_scale ? 1.0f / __ldg(a.grad_scale) : 1.0f;
    const uint64_t base = (uint64_t)(blockIdx.x - a.first_block[ti]) * kAdamPerBlock;
    const uint64_t n = T.n;
    const bool vec = ((reinterpret_cast<uintptr_t>(T.p) | reinterpret_cast<uintptr_t>(T.g) | reinterpret_cast<uintptr_t>(T.m) |
                       reinterpret_cast<uintptr_t>(T.v)) & 15) == 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const uint64_t i = base + ((uint64_t)k * kAdamThreads + threadIdx.x) * 4;
        if (i >= n) break;
        if (vec && i + 4 <= n) {
            float4 p = *reinterpret_cast<const float4*>(T.p + i);
            const float4 g = ld_stream4(reinterpret_cast<const float4*>(T.g + i));
            float4 m = *reinterpret_cast<const float4*>(T.m + i);
            float4 v = *reinterpret_cast<const float4*>(T.v + i);
            adam_one(p.x, g.x, m.x, v.x, inv_scale, a.weight_decay, a.beta1, a.beta2, step_size, rsqrt_bc2, a.eps);
            adam_one(p.y, g.y, m.y, v.y, inv_scale, a.weight_decay, a.beta1, a.beta2, step_size, rsqrt_bc2, a.eps);
            adam_one(p.z, g.z, m.z, v.z, inv_scale, a.weight_decay, a.beta1, a.beta2, step_size, rsqrt_bc2, a.eps);
            adam_one(p.w, g.w, m.w, v.w, inv_scale, a.weight_decay, a.beta1, a.beta2, step_size, rsqrt_bc2, a.eps);
            *reinterpret_cast<float4*>(T.p + i) = p;
            *reinterpret_cast<float4*>(T.m + i) = m;
            *reinterpret_cast<float4*>(T.v + i) = v;
            if (T.mirror) {
                char* mp = reinterpret_cast<char*>(T.mirror) + (i >> 1) * T.mirror_stride;
                *reinterpret_cast<__half2*>(mp) = __floats2half2_rn(p.x, p.y);
                *reinterpret_cast<__half2*>(mp + T.mirror_stride) = __floats2half2_rn(p.z, p.w);
            }
        } else {
            for (uint64_t j = i; j < n && j < i + 4; j++) {
                float p = T.p[j], m = T.m[j], v = T.v[j];
                adam_one(p, T.g[j], m, v, inv_scale, a.weight_decay, a.beta1, a.beta2, step_size, rsqrt_bc2, a.eps);
                T.p[j] = p; T.m[j] = m; T.v[j] = v;
                if (T.mirror) reinterpret_cast<__half*>(reinterpret_cast<char*>(T.mirror) + (j >> 1) * T.mirror_stride)[j & 1] = __float2half_rn(p);
            }
        }
    }
}

// step += 1 for every tensor of the launch unless the step was skipped (runs after k_adam on the same stream)
__global__ void k_adam_bump(const __grid_constant__ AdamArgs a) {
    if (a.found_inf && __ldg(a.found_inf) != 0.f) return;
    if (threadIdx.x < a.count) *const_cast<float*>(a.t[threadIdx.x].step) += 1.0f;
}

// found_inf := 1 if any gradient of the launch's tensors is not finite (GradScaler's check before the step; replaces
// torch._amp_foreach_non_finite_check_and_unscale_ with inv_scale = 1, which is what GradScaler.step runs for an optimizer that
// unscales on the fly: a multi-tensor kernel at 2.2 TB/s). One streaming pass over g: 8 elements per thread, one store per
// warp that saw a non-finite value. The caller zeroes found_inf.
__global__ void __launch_bounds__(kAdamThreads) k_found_inf(const __grid_constant__ AdamArgs a, float* __restrict__ found_inf) {
    uint32_t ti = 0;
#pragma unroll 1
    while (ti + 1 < a.count && blockIdx.x >= a.first_block[ti + 1]) ti++;
    const pnerf_adam_tensor& T = a.t[ti];
    const uint64_t base = (uint64_t)(blockIdx.x - a.first_block[ti]) * kAdamPerBlock;
    const uint64_t n = T.n;
    const bool vec = (reinterpret_cast<uintptr_t>(T.g) & 15) == 0;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const uint64_t i = base + ((uint64_t)k * kAdamThreads + threadIdx.x) * 4;
        if (i >= n) break;
        if (vec && i + 4 <= n) {
            const float4 g = *reinterpret_cast<const float4*>(T.g + i);     // (stays in L2 for the Adam kernel that follows)
            // x - x is 0 for finite x and NaN for +-inf / NaN: one add per element, the sum is NaN if any element is not finite
            const float s = (g.x - g.x) + (g.y - g.y) + (g.z - g.z) + (g.w - g.w);
            bad |= !(s == 0.f);
        } else {
            for (uint64_t j = i; j < n && j < i + 4; j++) { const float g = T.g[j]; bad |= !((g - g) == 0.f); }
        }
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *found_inf = 1.0f;
}

}  // namespace pnerf

using namespace pnerf;

extern "C" {

int pnerf_adam_step(const pnerf_adam_tensor* tensors, uint32_t count, float lr, const float* lr_dev, float beta1, float beta2,
                    float eps, float weight_decay, const float* grad_scale, const float* found_inf, void* stream) {
    if (count == 0) return PNERF_OK;
    PNERF_REQUIRE(tensors != nullptr && count <= (uint32_t)kAdamMaxTensors);
    PNERF_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f);
    AdamArgs a;
    uint64_t blocks = 0;
    for (uint32_t i = 0; i < count; i++) {
        PNERF_REQUIRE(tensors[i].p && tensors[i].g && tensors[i].m && tensors[i].v && tensors[i].step);
        if (tensors[i].mirror)
            PNERF_REQUIRE((tensors[i].n & 1) == 0 && (reinterpret_cast<uintptr_t>(tensors[i].mirror) & 3) == 0 &&
                          tensors[i].mirror_stride >= 4 && (tensors[i].mirror_stride & 3) == 0);
        a.t[i] = tensors[i];
        a.first_block[i] = (uint32_t)blocks;
        blocks += ceil_div<uint64_t>(tensors[i].n, kAdamPerBlock);
        if (blocks > 0x7fffffffull) return PNERF_ERR_UNSUPPORTED;
    }
    a.first_block[count] = (uint32_t)blocks;
    a.count = count; a.lr = lr; a.lr_dev = lr_dev; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
    a.grad_scale = grad_scale; a.found_inf = found_inf;
    cudaStream_t s = (cudaStream_t)stream;
    if (blocks) k_adam<<<(uint32_t)blocks, kAdamThreads, 0, s>>>(a);
    k_adam_bump<<<1, kAdamMaxTensors, 0, s>>>(a);
    return check_launch("adam_step");
}

int pnerf_found_inf(const pnerf_adam_tensor* tensors, uint32_t count, float* found_inf, void* stream) {
    if (count == 0) return PNERF_OK;
    PNERF_REQUIRE(tensors != nullptr && found_inf != nullptr && count <= (uint32_t)kAdamMaxTensors);
    AdamArgs a;
    uint64_t blocks = 0;
    for (uint32_t i = 0; i < count; i++) {
        PNERF_REQUIRE(tensors[i].g != nullptr);
        a.t[i] = tensors[i];
        a.first_block[i] = (uint32_t)blocks;
        blocks += ceil_div<uint64_t>(tensors[i].n, kAdamPerBlock);
        if (blocks > 0x7fffffffull) return PNERF_ERR_UNSUPPORTED;
    }
    a.first_block[count] = (uint32_t)blocks;
    a.count = count; a.lr = 0.f; a.lr_dev = nullptr; a.beta1 = a.beta2 = a.eps = a.weight_decay = 0.f;
    a.grad_scale = nullptr; a.found_inf = nullptr;
    if (blocks) k_found_inf<<<(uint32_t)blocks, kAdamThreads, 0, (cudaStream_t)stream>>>(a, found_inf);
    return check_launch("found_inf");
}

}  // extern "C"
