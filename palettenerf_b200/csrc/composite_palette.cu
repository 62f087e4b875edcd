// composite_palette.cu — ONE-pass compositor of the palette training step (B200, sm_100a).
//
// The reference composites the training samples twice per step: composite_rays_train on (sigma, rgb) and
// composite_rays_flex_train on the 13+clip+Nb auxiliary channels (palette/renderer.py:354, 387-397), each re-reading
// sigma / deltas and re-deriving the transmittance, and its backward needs two zero-filled [M, ...] gradient buffers
// (raymarching/raymarching.py:283-284, 335). Here both happen in one warp-per-ray pass:
//   forward : lanes stride the ray's samples; every lane accumulates w_k * (rgb_k, t_k, 1, flex_k[0..NF)) for its own
//             samples (NF + 5 independent FMAs per sample, all loads of a chunk in flight at once), one warp reduction
//             per channel at the END of the ray instead of one shuffle per sample;
//   backward: grad_rgb_k = g_image * w_k, grad_flex_k = g_maps * w_k with the reference's termination rules (the
//             terminating sample gets an rgb gradient but no flex gradient, raymarching.cu:561-564 vs :806-811), and
//             zeros for every later sample of the ray — so the gradient buffers need NO memset. sigma is a constant of
//             the palette stage (palette/renderer.py:334-335): no sigma gradient is produced.
// Arithmetic per sample is the reference's (alpha = 1 - __expf(-sigma dt), w = alpha T, T *= 1 - alpha); sums are
// reassociated (lane-strided partial sums), covered by the tolerance in tests/test_fused_train_gpu.py.
#include "common.cuh"
#include "composite_common.cuh"

namespace pnerf {

template <int NF>
__global__ void __launch_bounds__(256) k_pal_comp_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                      const float* __restrict__ flex, const float* __restrict__ deltas,
                                                      const int32_t* __restrict__ rays, uint32_t M, uint32_t N, float T_thresh,
                                                      float* __restrict__ weights_sum, float* __restrict__ depth,
                                                      float* __restrict__ image, float* __restrict__ maps) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1];
    const uint32_t num_steps = (uint32_t)rays[n * 3 + 2];
    float r = 0, g = 0, b = 0, ws = 0, d = 0;
    float acc[NF];
#pragma unroll
    for (int c = 0; c < NF; c++) acc[c] = 0.f;
    const bool ok_rgb = num_steps != 0 && offset + num_steps <= M;    // ref: raymarching.cu:524
    const bool ok_flex = num_steps != 0 && offset + num_steps < M;    // ref: raymarching.cu:601 (sic: >= drops the ray)
    if (ok_rgb) {
        float T = 1.0f, t_carry = 0.0f;
        for (uint32_t base = 0; base < num_steps; base += 32) {
            const uint32_t k = base + lane;
            const bool valid = k < num_steps;
            const size_t s = (size_t)offset + k;
            float alpha = 0.f, rdt = 0.f;
            if (valid) {
                const float2 dl = reinterpret_cast<const float2*>(deltas)[s];
                alpha = 1.0f - __expf(-sigmas[s] * dl.x);
                rdt = dl.y;
            }
            const ChunkT ct = chunk_transmittance(alpha, valid, T, T_thresh, lane);
            const float t_incl = t_carry + warp_scan_add(rdt, lane);
            t_carry = __shfl_sync(0xffffffffu, t_incl, 31);
            if (valid && lane <= ct.last) {
                const float w = alpha * ct.T_before;
                r += w * rgbs[s * 3 + 0]; g += w * rgbs[s * 3 + 1]; b += w * rgbs[s * 3 + 2];
                d += w * t_incl;
                ws += w;
                if (ok_flex) {
                    const float* row = flex + s * NF;
#pragma unroll
                    for (int c = 0; c < NF; c++) acc[c] += w * row[c];
                }
            }
            if (ct.last < 32u) break;
        }
        r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); ws = warp_sum(ws); d = warp_sum(d);
    }
#pragma unroll
    for (int c = 0; c < NF; c++) {
        const float tot = warp_sum(acc[c]);
        if (lane == (uint32_t)(c & 31)) maps[(size_t)index * NF + c] = tot;
    }
    if (lane == 0) {
        weights_sum[index] = ws;
        depth[index] = d;
        image[index * 3 + 0] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

template <int NF>
__global__ void __launch_bounds__(256) k_pal_comp_bwd(const float* __restrict__ grad_image, const float* __restrict__ grad_maps,
                                                      const float* __restrict__ sigmas, const float* __restrict__ deltas,
                                                      const int32_t* __restrict__ rays, uint32_t M, uint32_t N, float T_thresh,
                                                      float* __restrict__ grad_rgbs, float* __restrict__ grad_flex) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1];
    const uint32_t num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset >= M) return;
    const bool ok_rgb = offset + num_steps <= M, ok_flex = offset + num_steps < M;
    const uint32_t n_write = min(num_steps, M - offset);     // rows of this ray that exist in the buffers
    const float g0 = grad_image[index * 3 + 0], g1 = grad_image[index * 3 + 1], g2 = grad_image[index * 3 + 2];
    // lanes = channels for the flex rows: every sample row (NF floats) is written with NF/32 coalesced stores
    constexpr int KC = (NF + 31) / 32;
    float gm[KC];
#pragma unroll
    for (int i = 0; i < KC; i++) {
        const uint32_t c = lane + 32u * i;
        gm[i] = (c < (uint32_t)NF) ? __ldg(grad_maps + (size_t)index * NF + c) : 0.f;
    }
    float T = 1.0f;
    bool done = !ok_rgb;                                      // warp-uniform: the ray has terminated (or was dropped)
    for (uint32_t base = 0; base < n_write; base += 32) {
        const uint32_t k = base + lane;
        const bool valid = k < n_write;
        const size_t s = (size_t)offset + k;
        float w = 0.f;
        uint32_t last = 32u;
        if (!done) {
            float alpha = 0.f;
            if (valid) alpha = 1.0f - __expf(-sigmas[s] * deltas[s * 2]);
            const ChunkT ct = chunk_transmittance(alpha, valid, T, T_thresh, lane);
            w = alpha * ct.T_before;
            last = ct.last;
        }
        const float wr = (!done && lane <= last) ? w : 0.f;              // terminating sample included
        const float wf = (!done && ok_flex && lane < last) ? w : 0.f;    // terminating sample excluded (ref :806-811)
        if (valid) { grad_rgbs[s * 3 + 0] = g0 * wr; grad_rgbs[s * 3 + 1] = g1 * wr; grad_rgbs[s * 3 + 2] = g2 * wr; }
        const uint32_t cnt = min(32u, n_write - base);
        float* rows = grad_flex + ((size_t)offset + base) * NF;
#pragma unroll 4
        for (uint32_t j = 0; j < cnt; j++) {
            const float wj = __shfl_sync(0xffffffffu, wf, j);
#pragma unroll
            for (int i = 0; i < KC; i++) {
                const uint32_t c = lane + 32u * i;
                if (c < (uint32_t)NF) rows[(size_t)j * NF + c] = gm[i] * wj;
            }
        }
        if (last < 32u) done = true;
    }
}

}  // namespace pnerf

using namespace pnerf;

extern "C" {

int pnerf_palette_composite_train_forward(const float* sigmas, const float* rgbs, const float* flex, const float* deltas,
                                          const int32_t* rays, uint32_t M, uint32_t N, uint32_t nflex, float T_thresh,
                                          float* weights_sum, float* depth, float* image, float* maps, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(sigmas && rgbs && flex && deltas && rays && weights_sum && depth && image && maps);
    if (nflex != 33) return PNERF_ERR_UNSUPPORTED;            // 13 + clip_dim 16 + 4 bases: the reference's default
    k_pal_comp_fwd<33><<<ceil_div(N, 8u), 256, 0, (cudaStream_t)stream>>>(sigmas, rgbs, flex, deltas, rays, M, N, T_thresh,
                                                                         weights_sum, depth, image, maps);
    return check_launch("palette_composite_train_forward");
}

int pnerf_palette_composite_train_backward(const float* grad_image, const float* grad_maps, const float* sigmas,
                                           const float* deltas, const int32_t* rays, uint32_t M, uint32_t N, uint32_t nflex,
                                           float T_thresh, float* grad_rgbs, float* grad_flex, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(grad_image && grad_maps && sigmas && deltas && rays && grad_rgbs && grad_flex);
    if (nflex != 33) return PNERF_ERR_UNSUPPORTED;
    k_pal_comp_bwd<33><<<ceil_div(N, 8u), 256, 0, (cudaStream_t)stream>>>(grad_image, grad_maps, sigmas, deltas, rays, M, N,
                                                                         T_thresh, grad_rgbs, grad_flex);
    return check_launch("palette_composite_train_backward");
}

}  // extern "C"
