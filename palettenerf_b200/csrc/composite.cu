// composite.cu — front-to-back volume compositing for B200 (sm_100a).
//
// Replaces raymarching/src/raymarching.cu:504-844 (training forward/backward, rgb and n-channel "flex" variants)
// and :1025-1205 (inference, in-place) of the reference.
//
// Training kernels are warp-per-ray instead of thread-per-ray: lanes stride the ray's samples (coalesced loads
// of sigma / deltas / rgb), transmittance comes from a warp product-scan carried across 32-sample chunks, early
// termination is a ballot (the terminating sample is still accumulated, as raymarching.cu:561-564 does), and the
// n-channel variant turns lanes into channels for the accumulation so each sample row is one coalesced load and
// no per-thread temp[128] array exists. Summation order therefore differs from the reference's serial loop by
// fp32 reassociation only (tolerance stated in tests/test_composite.py).
#include "common.cuh"
#include "composite_common.cuh"

namespace pnerf {

// ------------------------------------------------------------------------------------------------
// composite_rays_train forward (ref: raymarching.cu:504-580)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_comp_train_fwd(const float* __restrict__ sigmas,
                                                        const float* __restrict__ rgbs,
                                                        const float* __restrict__ deltas,
                                                        const int32_t* __restrict__ rays, uint32_t M, uint32_t N,
                                                        float T_thresh, float* __restrict__ weights_sum,
                                                        float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1];
    const uint32_t num_steps = (uint32_t)rays[n * 3 + 2];

    float r = 0, g = 0, b = 0, ws = 0, d = 0;
    if (num_steps != 0 && offset + num_steps <= M) {
        float T = 1.0f, t_carry = 0.0f;
        for (uint32_t base = 0; base < num_steps; base += 32) {
            const uint32_t k = base + lane;
            const bool valid = k < num_steps;
            float alpha = 0.f, rdt = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
            if (valid) {
                const size_t s = (size_t)offset + k;
                const float2 dl = reinterpret_cast<const float2*>(deltas)[s];
                alpha = 1.0f - __expf(-sigmas[s] * dl.x);
                rdt = dl.y;
                c0 = rgbs[s * 3 + 0]; c1 = rgbs[s * 3 + 1]; c2 = rgbs[s * 3 + 2];
            }
            const ChunkT ct = chunk_transmittance(alpha, valid, T, T_thresh, lane);
            const float t_incl = t_carry + warp_scan_add(rdt, lane);
            t_carry = __shfl_sync(0xffffffffu, t_incl, 31);
            if (valid && lane <= ct.last) {
                const float w = alpha * ct.T_before;
                r += w * c0; g += w * c1; b += w * c2;
                d += w * t_incl;
                ws += w;
            }
            if (ct.last < 32u) break;
        }
        r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); ws = warp_sum(ws); d = warp_sum(d);
    }
    if (lane == 0) {
        weights_sum[index] = ws;
        depth[index] = d;
        image[index * 3 + 0] = r;
        image[index * 3 + 1] = g;
        image[index * 3 + 2] = b;
    }
}

// ------------------------------------------------------------------------------------------------
// composite_rays_train backward (ref: raymarching.cu:681-761)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_comp_train_bwd(
    const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_image,
    const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
    const int32_t* __restrict__ rays, const float* __restrict__ weights_sum, const float* __restrict__ image,
    uint32_t M, uint32_t N, float T_thresh, float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1];
    const uint32_t num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) return;

    const float g0 = grad_image[index * 3 + 0], g1 = grad_image[index * 3 + 1], g2 = grad_image[index * 3 + 2];
    const float gws = grad_weights_sum[index];
    const float r_final = image[index * 3 + 0], g_final = image[index * 3 + 1], b_final = image[index * 3 + 2];
    const float ws_term = gws * (1 - weights_sum[index]);

    float T = 1.0f, r_carry = 0.f, g_carry = 0.f, b_carry = 0.f;
    for (uint32_t base = 0; base < num_steps; base += 32) {
        const uint32_t k = base + lane;
        const bool valid = k < num_steps;
        const size_t s = (size_t)offset + k;
        float alpha = 0.f, dt = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (valid) {
            dt = deltas[s * 2];
            alpha = 1.0f - __expf(-sigmas[s] * dt);
            c0 = rgbs[s * 3 + 0]; c1 = rgbs[s * 3 + 1]; c2 = rgbs[s * 3 + 2];
        }
        const ChunkT ct = chunk_transmittance(alpha, valid, T, T_thresh, lane);
        const float w = alpha * ct.T_before;
        const float r_acc = r_carry + warp_scan_add(w * c0, lane);
        const float g_acc = g_carry + warp_scan_add(w * c1, lane);
        const float b_acc = b_carry + warp_scan_add(w * c2, lane);
        r_carry = __shfl_sync(0xffffffffu, r_acc, 31);
        g_carry = __shfl_sync(0xffffffffu, g_acc, 31);
        b_carry = __shfl_sync(0xffffffffu, b_acc, 31);
        if (valid && lane <= ct.last) {
            grad_rgbs[s * 3 + 0] = g0 * w;
            grad_rgbs[s * 3 + 1] = g1 * w;
            grad_rgbs[s * 3 + 2] = g2 * w;
            grad_sigmas[s] = dt * (g0 * (ct.T_after * c0 - (r_final - r_acc)) +
                                   g1 * (ct.T_after * c1 - (g_final - g_acc)) +
                                   g2 * (ct.T_after * c2 - (b_final - b_acc)) + ws_term);
        }
        if (ct.last < 32u) break;
    }
}

// ------------------------------------------------------------------------------------------------
// n-channel ("flex") training forward / backward (ref: raymarching.cu:583-645, 764-819)
// KC = ceil(n_channel / 32) accumulators per lane.
// ------------------------------------------------------------------------------------------------
template <int KC>
__global__ void __launch_bounds__(256) k_comp_flex_train_fwd(const float* __restrict__ sigmas,
                                                             const float* __restrict__ input,
                                                             const float* __restrict__ deltas,
                                                             const int32_t* __restrict__ rays, uint32_t M, uint32_t N,
                                                             uint32_t nc, float T_thresh, float* __restrict__ output) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1];
    const uint32_t num_steps = (uint32_t)rays[n * 3 + 2];
    float acc[KC];
#pragma unroll
    for (int i = 0; i < KC; i++) acc[i] = 0.f;

    // NB: the reference drops rays with offset + num_steps >= M here (raymarching.cu:601), not > M.
    if (num_steps != 0 && offset + num_steps < M) {
        float T = 1.0f;
        for (uint32_t base = 0; base < num_steps; base += 32) {
            const uint32_t k = base + lane;
            const bool valid = k < num_steps;
            float alpha = 0.f;
            if (valid) {
                const size_t s = (size_t)offset + k;
                alpha = 1.0f - __expf(-sigmas[s] * deltas[s * 2]);
            }
            const ChunkT ct = chunk_transmittance(alpha, valid, T, T_thresh, lane);
            const float w = alpha * ct.T_before;
            const uint32_t cnt = min(min(32u, num_steps - base), ct.last + 1u);
            const float* row = input + ((size_t)offset + base) * nc;
#pragma unroll 4
            for (uint32_t j = 0; j < cnt; j++) {
                const float wj = __shfl_sync(0xffffffffu, w, j);
#pragma unroll
                for (int i = 0; i < KC; i++) {
                    const uint32_t c = lane + 32u * i;
                    if (c < nc) acc[i] += wj * row[(size_t)j * nc + c];
                }
            }
            if (ct.last < 32u) break;
        }
    }
#pragma unroll
    for (int i = 0; i < KC; i++) {
        const uint32_t c = lane + 32u * i;
        if (c < nc) output[(size_t)index * nc + c] = acc[i];
    }
}

template <int KC>
__global__ void __launch_bounds__(256) k_comp_flex_train_bwd(const float* __restrict__ grad_output,
                                                             const float* __restrict__ sigmas,
                                                             const float* __restrict__ deltas,
                                                             const int32_t* __restrict__ rays, uint32_t M, uint32_t N,
                                                             uint32_t nc, float T_thresh,
                                                             float* __restrict__ grad_input) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1];
    const uint32_t num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps >= M) return;
    float g[KC];
#pragma unroll
    for (int i = 0; i < KC; i++) {
        const uint32_t c = lane + 32u * i;
        g[i] = (c < nc) ? grad_output[(size_t)index * nc + c] : 0.f;
    }
    float T = 1.0f;
    for (uint32_t base = 0; base < num_steps; base += 32) {
        const uint32_t k = base + lane;
        const bool valid = k < num_steps;
        float alpha = 0.f;
        if (valid) {
            const size_t s = (size_t)offset + k;
            alpha = 1.0f - __expf(-sigmas[s] * deltas[s * 2]);
        }
        const ChunkT ct = chunk_transmittance(alpha, valid, T, T_thresh, lane);
        const float w = alpha * ct.T_before;
        // the reference breaks *before* writing the terminating sample's gradient (raymarching.cu:806-811)
        const uint32_t cnt = min(min(32u, num_steps - base), ct.last);
        float* row = grad_input + ((size_t)offset + base) * nc;
        for (uint32_t j = 0; j < cnt; j++) {
            const float wj = __shfl_sync(0xffffffffu, w, j);
#pragma unroll
            for (int i = 0; i < KC; i++) {
                const uint32_t c = lane + 32u * i;
                if (c < nc) row[(size_t)j * nc + c] = g[i] * wj;
            }
        }
        if (ct.last < 32u) break;
    }
}

// ------------------------------------------------------------------------------------------------
// inference compositing (ref: raymarching.cu:1025-1111): thread per alive ray, n_step <= 8 samples, same serial
// arithmetic as the reference (bit-identical accumulators); this is the compatibility path — the fused render
// kernel keeps these accumulators in registers instead.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_comp_rays(uint32_t n_alive, uint32_t n_step, float T_thresh,
                                                   int32_t* __restrict__ rays_alive, float* __restrict__ rays_t,
                                                   const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                   const float* __restrict__ deltas, float* __restrict__ weights_sum,
                                                   float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const size_t s0 = (size_t)n * n_step;
    float t = rays_t[index];
    float ws = weights_sum[index], d = depth[index];
    float r = image[index * 3 + 0], g = image[index * 3 + 1], b = image[index * 3 + 2];
    uint32_t step = 0;
    while (step < n_step) {
        const size_t s = s0 + step;
        const float dt = deltas[s * 2];
        if (dt == 0) break;
        const float alpha = 1.0f - __expf(-sigmas[s] * dt);
        const float T = 1 - ws;
        const float w = alpha * T;
        ws += w;
        t += deltas[s * 2 + 1];
        d += w * t;
        r += w * rgbs[s * 3 + 0];
        g += w * rgbs[s * 3 + 1];
        b += w * rgbs[s * 3 + 2];
        if (T < T_thresh) break;
        step++;
    }
    if (step < n_step) rays_alive[n] = -1;
    else rays_t[index] = t;
    weights_sum[index] = ws;
    depth[index] = d;
    image[index * 3 + 0] = r;
    image[index * 3 + 1] = g;
    image[index * 3 + 2] = b;
}

// n-channel inference accumulate (ref: raymarching.cu:1114-1185): thread per (ray, channel); the weights are
// recomputed per channel thread (<= 8 __expf) so that the input rows are read coalesced and no temp[128] exists.
__global__ void __launch_bounds__(256) k_comp_rays_flex(uint32_t n_alive, uint32_t n_step, uint32_t nc,
                                                        float T_thresh, const int32_t* __restrict__ rays_alive,
                                                        const float* __restrict__ sigmas,
                                                        const float* __restrict__ input,
                                                        const float* __restrict__ deltas,
                                                        const float* __restrict__ weights_sum,
                                                        float* __restrict__ output) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = (uint32_t)(tid / nc), c = (uint32_t)(tid % nc);
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const size_t s0 = (size_t)n * n_step;
    float ws = weights_sum[index];
    float acc = output[(size_t)index * nc + c];
    for (uint32_t step = 0; step < n_step; step++) {
        const size_t s = s0 + step;
        const float dt = deltas[s * 2];
        if (dt == 0) break;
        const float alpha = 1.0f - __expf(-sigmas[s] * dt);
        const float T = 1 - ws;
        const float w = alpha * T;
        ws += w;
        acc += w * input[s * nc + c];
        if (T < T_thresh) break;
    }
    output[(size_t)index * nc + c] = acc;
}

}  // namespace pnerf

using namespace pnerf;

extern "C" {

int pnerf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                       uint32_t M, uint32_t N, float T_thresh, float* weights_sum, float* depth,
                                       float* image, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(sigmas && rgbs && deltas && rays && weights_sum && depth && image);
    k_comp_train_fwd<<<ceil_div(N, 8u), 256, 0, (cudaStream_t)stream>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh,
                                                                        weights_sum, depth, image);
    return check_launch("composite_rays_train_forward");
}

int pnerf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                        const float* rgbs, const float* deltas, const int32_t* rays,
                                        const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                        float T_thresh, float* grad_sigmas, float* grad_rgbs, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(grad_weights_sum && grad_image && sigmas && rgbs && deltas && rays && weights_sum && image &&
                  grad_sigmas && grad_rgbs);
    k_comp_train_bwd<<<ceil_div(N, 8u), 256, 0, (cudaStream_t)stream>>>(grad_weights_sum, grad_image, sigmas, rgbs,
                                                                        deltas, rays, weights_sum, image, M, N,
                                                                        T_thresh, grad_sigmas, grad_rgbs);
    return check_launch("composite_rays_train_backward");
}

int pnerf_composite_rays_flex_train_forward(const float* sigmas, const float* input, const float* deltas,
                                            const int32_t* rays, uint32_t M, uint32_t N, uint32_t n_channel,
                                            float T_thresh, float* output, void* stream) {
    if (N == 0 || n_channel == 0) return PNERF_OK;
    PNERF_REQUIRE(sigmas && input && deltas && rays && output);
    if (n_channel > 128) return PNERF_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t grid = ceil_div(N, 8u);
    switch ((n_channel + 31) / 32) {
        case 1: k_comp_flex_train_fwd<1><<<grid, 256, 0, s>>>(sigmas, input, deltas, rays, M, N, n_channel, T_thresh, output); break;
        case 2: k_comp_flex_train_fwd<2><<<grid, 256, 0, s>>>(sigmas, input, deltas, rays, M, N, n_channel, T_thresh, output); break;
        case 3: k_comp_flex_train_fwd<3><<<grid, 256, 0, s>>>(sigmas, input, deltas, rays, M, N, n_channel, T_thresh, output); break;
        default: k_comp_flex_train_fwd<4><<<grid, 256, 0, s>>>(sigmas, input, deltas, rays, M, N, n_channel, T_thresh, output); break;
    }
    return check_launch("composite_rays_flex_train_forward");
}

int pnerf_composite_rays_flex_train_backward(const float* grad_output, const float* sigmas, const float* input,
                                             const float* deltas, const int32_t* rays, const float* output, uint32_t M,
                                             uint32_t N, uint32_t n_channel, float T_thresh, float* grad_input,
                                             void* stream) {
    if (N == 0 || n_channel == 0) return PNERF_OK;
    (void)input; (void)output;  // the reference kernel takes but never reads them (raymarching.cu:764-819)
    PNERF_REQUIRE(grad_output && sigmas && deltas && rays && grad_input);
    if (n_channel > 128) return PNERF_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t grid = ceil_div(N, 8u);
    switch ((n_channel + 31) / 32) {
        case 1: k_comp_flex_train_bwd<1><<<grid, 256, 0, s>>>(grad_output, sigmas, deltas, rays, M, N, n_channel, T_thresh, grad_input); break;
        case 2: k_comp_flex_train_bwd<2><<<grid, 256, 0, s>>>(grad_output, sigmas, deltas, rays, M, N, n_channel, T_thresh, grad_input); break;
        case 3: k_comp_flex_train_bwd<3><<<grid, 256, 0, s>>>(grad_output, sigmas, deltas, rays, M, N, n_channel, T_thresh, grad_input); break;
        default: k_comp_flex_train_bwd<4><<<grid, 256, 0, s>>>(grad_output, sigmas, deltas, rays, M, N, n_channel, T_thresh, grad_input); break;
    }
    return check_launch("composite_rays_flex_train_backward");
}

int pnerf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                         const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum, float* depth,
                         float* image, void* stream) {
    if (n_alive == 0) return PNERF_OK;
    PNERF_REQUIRE(rays_alive && rays_t && sigmas && rgbs && deltas && weights_sum && depth && image);
    k_comp_rays<<<ceil_div(n_alive, 128u), 128, 0, (cudaStream_t)stream>>>(n_alive, n_step, T_thresh, rays_alive, rays_t,
                                                                          sigmas, rgbs, deltas, weights_sum, depth,
                                                                          image);
    return check_launch("composite_rays");
}

int pnerf_composite_rays_flex(uint32_t n_alive, uint32_t n_step, uint32_t n_channel, float T_thresh,
                              const int32_t* rays_alive, const float* rays_t, const float* sigmas, const float* input,
                              const float* deltas, const float* weights_sum, float* output, void* stream) {
    if (n_alive == 0 || n_channel == 0) return PNERF_OK;
    (void)rays_t;
    PNERF_REQUIRE(rays_alive && sigmas && input && deltas && weights_sum && output);
    if (n_channel > 128) return PNERF_ERR_UNSUPPORTED;
    const uint64_t threads = (uint64_t)n_alive * n_channel;
    k_comp_rays_flex<<<(uint32_t)ceil_div<uint64_t>(threads, 256), 256, 0, (cudaStream_t)stream>>>(
        n_alive, n_step, n_channel, T_thresh, rays_alive, sigmas, input, deltas, weights_sum, output);
    return check_launch("composite_rays_flex");
}

}  // extern "C"
