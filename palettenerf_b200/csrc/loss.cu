// loss.cu — the per-ray losses of the palette training step in ONE launch (value, its terms and every gradient).
//
// Replaces the tensor program of PaletteTrainer.train_step after model.render (ref: palette/utils.py:486-567):
//   loss = mean_rays( mean_c (image - gt)^2 )                                   (:486)
//        + mean( (direct_rgb - gt)^2 )                                          (:487)
//        + [pred_clip] mean( (clip_feat - gt_clip)^2 )                          (:490-492)
//        + l_sparsity * mean(omega_sparsity) + l_offsets * mean(offsets_norm) + l_view_dep * mean(view_dep_norm)   (:521-529, 546-553)
//        + l_smooth * mean(smooth_norm)                                         (:538-542, 555)
//        + l_weight * mean( (gt_weights - basis_acc)^2 )                        (:532-536, 564)
//        + l_palette * mean_b( sum_c (basis_color - basis_color_origin)^2 )     (:544, 561)
// In torch this is ~16 elementwise / reduction kernels forward and ~25 backward, each 2-3 us of launch-bound work on
// [N]-sized maps (N = 4096 rays): ~14 % of the captured training step. Here one pass over the inputs (64 rays per
// CTA, coalesced) writes the gradient of every input (for an upstream gradient of 1) and per-CTA partial sums; a
// second, tiny launch adds the partials in a fixed order (deterministic, no atomics, no zero-initialised scratch).
// The autograd Function scales the gradients by the incoming gradient (the GradScaler's scale) in its backward.
//
// The regulariser / feature / blending-weight maps are COLUMNS of the renderer's channel-composite output
// (`maps` [N, stride]); the gradient is written as one [N, stride] tensor (zero in the columns no term reads), so
// autograd does not assemble it from per-slice zero-fill + copy + add kernels.
#include "common.cuh"

namespace pnerf {

constexpr int kLossTerms = 10;     // total, rgb, direct, clip, sparsity, offsets, view_dep, smooth, weight, palette
constexpr int kLossThreads = 256;
constexpr int kLossRays = 64;      // rays per CTA
constexpr int kMaxStride = 256;

struct LossArgs {
    const float* image; const float* direct_rgb; const float* gt_rgb;      // [N,3]
    const float* maps; uint32_t stride;                                      // [N, stride]
    int col_sparsity, col_offsets, col_view_dep, col_smooth, col_clip, col_basis;   // -1: term absent
    uint32_t clip_dim, num_basis;
    const float* gt_clip; const float* gt_weights;                           // [N, clip_dim], [N, num_basis]
    const float* basis_color; const float* basis_color_origin;               // [num_basis, 3] or NULL
    float l_sparsity, l_offsets, l_view_dep, l_smooth, l_weight, l_palette;
    uint32_t N;
    float* terms; float* per_ray;                                            // [10], [N] (per-ray rgb error; may be NULL)
    float* g_image; float* g_direct; float* g_maps; float* g_basis_color;
    float* partials;                                                         // [blocks, kLossTerms] scratch
};

// pass 1: CTA b owns rays [64 b, 64 b + 64): its slice of image / direct_rgb (192 floats each) and of maps (64 x stride
// floats, contiguous) is read with coalesced accesses, one element per thread and step; the gradient of every element
// is written on the spot; the CTA's partial sums of the nine terms go to partials[b] (fixed-order tree: deterministic).
__global__ void __launch_bounds__(kLossThreads) k_palette_loss(LossArgs a) {
    __shared__ uint8_t kind[kMaxStride];       // term a column feeds: 0 none, 3 clip, 4 sparsity, 5 offsets, 6 view_dep, 7 smooth, 8 weight
    __shared__ float red[kLossTerms][kLossThreads / 32];
    for (uint32_t c = threadIdx.x; c < a.stride; c += kLossThreads) {
        const int ci = (int)c;
        uint8_t k = 0;
        if (ci == a.col_sparsity) k = 4;
        else if (ci == a.col_offsets) k = 5;
        else if (ci == a.col_view_dep) k = 6;
        else if (ci == a.col_smooth) k = 7;
        else if (a.col_clip >= 0 && ci >= a.col_clip && ci < a.col_clip + (int)a.clip_dim) k = 3;
        else if (a.col_basis >= 0 && ci >= a.col_basis && ci < a.col_basis + (int)a.num_basis) k = 8;
        kind[c] = k;
    }
    __syncthreads();
    float acc[kLossTerms];
#pragma unroll
    for (int i = 0; i < kLossTerms; i++) acc[i] = 0.f;
    const float invN = 1.0f / (float)a.N;
    const float k_rgb = 2.0f / (3.0f * (float)a.N);
    const float k_clip = a.col_clip >= 0 ? 2.0f / ((float)a.N * (float)a.clip_dim) : 0.f;
    const float k_w = a.col_basis >= 0 ? 2.0f * a.l_weight / ((float)a.N * (float)a.num_basis) : 0.f;
    const uint32_t n0 = blockIdx.x * kLossRays, rays = min((uint32_t)kLossRays, a.N - n0);
    // rgb terms: element e = (ray, channel)
    for (uint32_t e = threadIdx.x; e < rays * 3; e += kLossThreads) {
        const size_t i = (size_t)n0 * 3 + e;
        const float gt = a.gt_rgb[i];
        const float di = a.image[i] - gt, dd = a.direct_rgb[i] - gt;
        acc[1] += di * di; acc[2] += dd * dd;
        a.g_image[i] = k_rgb * di;
        a.g_direct[i] = k_rgb * dd;
    }
    if (a.per_ray) {
        for (uint32_t r = threadIdx.x; r < rays; r += kLossThreads) {
            const size_t i = (size_t)(n0 + r) * 3;
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < 3; c++) { const float d = a.image[i + c] - a.gt_rgb[i + c]; e += d * d; }
            a.per_ray[n0 + r] = e * (1.0f / 3.0f);
        }
    }
    // channel terms: element e = (ray, column) of the CTA's contiguous slice of maps
    for (uint32_t e = threadIdx.x; e < rays * a.stride; e += kLossThreads) {
        const uint32_t r = e / a.stride, c = e - r * a.stride;
        const size_t i = (size_t)n0 * a.stride + e;
        const uint8_t k = kind[c];
        float g = 0.f;
        if (k) {
            const float v = a.maps[i];
            if (k == 4) { acc[4] += v; g = a.l_sparsity * invN; }
            else if (k == 5) { acc[5] += v; g = a.l_offsets * invN; }
            else if (k == 6) { acc[6] += v; g = a.l_view_dep * invN; }
            else if (k == 7) { acc[7] += v; g = a.l_smooth * invN; }
            else if (k == 3) {
                const float d = v - a.gt_clip[(size_t)(n0 + r) * a.clip_dim + (c - a.col_clip)];
                acc[3] += d * d; g = k_clip * d;
            } else {
                const float d = v - a.gt_weights[(size_t)(n0 + r) * a.num_basis + (c - a.col_basis)];
                acc[8] += d * d; g = k_w * d;
            }
        }
        a.g_maps[i] = g;
    }
    if (blockIdx.x == 0 && a.basis_color && threadIdx.x < a.num_basis * 3) {
        const float d = a.basis_color[threadIdx.x] - a.basis_color_origin[threadIdx.x];
        acc[9] = d * d;
        a.g_basis_color[threadIdx.x] = 2.0f * a.l_palette * d / (float)a.num_basis;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 1; i < kLossTerms; i++) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[i][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLossTerms && threadIdx.x >= 1) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kLossThreads / 32; w++) v += red[threadIdx.x][w];
        a.partials[(size_t)blockIdx.x * kLossTerms + threadIdx.x] = v;
    }
}

// pass 2: one warp per term sums the CTA partials in a fixed order and applies the normalisation / lambda
__global__ void __launch_bounds__(32 * kLossTerms) k_palette_loss_finish(LossArgs a, uint32_t blocks) {
    __shared__ float t[kLossTerms];
    const int lane = threadIdx.x & 31, term = threadIdx.x >> 5;
    float v = 0.f;
    if (term >= 1)
        for (uint32_t b = lane; b < blocks; b += 32) v += a.partials[(size_t)b * kLossTerms + term];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) t[term] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        const float N = (float)a.N, invN = 1.0f / N;
        t[1] = t[1] / (3.0f * N);
        t[2] = t[2] / (3.0f * N);
        t[3] = a.col_clip >= 0 ? t[3] / (N * (float)a.clip_dim) : 0.f;
        t[4] = a.col_sparsity >= 0 ? a.l_sparsity * t[4] * invN : 0.f;
        t[5] = a.col_offsets >= 0 ? a.l_offsets * t[5] * invN : 0.f;
        t[6] = a.col_view_dep >= 0 ? a.l_view_dep * t[6] * invN : 0.f;
        t[7] = a.col_smooth >= 0 ? a.l_smooth * t[7] * invN : 0.f;
        t[8] = a.col_basis >= 0 ? a.l_weight * t[8] / (N * (float)a.num_basis) : 0.f;
        t[9] = a.basis_color ? a.l_palette * t[9] / (float)a.num_basis : 0.f;
        float total = 0.f;
#pragma unroll
        for (int i = 1; i < kLossTerms; i++) { a.terms[i] = t[i]; total += t[i]; }
        a.terms[0] = total;
    }
}

// out[i] *= *s for up to four buffers in one launch (the backward of the loss: upstream gradient = the loss scale)
struct ScaleArgs { float* p[4]; uint32_t n[4]; const float* s; };
__global__ void __launch_bounds__(256) k_scale_buffers(ScaleArgs a) {
    const float s = __ldg(a.s);
    const uint32_t stride = gridDim.x * blockDim.x;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        if (!a.p[b]) continue;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n[b]; i += stride) a.p[b][i] *= s;
    }
}


// ------------------------------------------------------------------------------------------------
// Smooth-loss channel of the training field (ref: palette/renderer.py:360-381). Per sample, from the field's channel row
// `ch` and the row `cj` of the SAME field evaluated at a jittered position:
//     gate   = exp(-|x - x_j|^2 / bound^2 / s_xyz - |diffuse - diffuse_j|^2 / s_color - |clip - clip_j| / s_clip)   (a constant)
//     smooth = gate * (sum_b (omega_j - omega)^2 + sum_c (clip_j - clip)^2)
// written into column 3 of `ch` in place (the field leaves that column zero and its backward never reads it). The torch
// expressions for this are ~30 elementwise launches over the full static sample capacity forward and backward (3.7 ms per
// step at 4096 rays); these two kernels touch the valid rows only.
// Row layout: 3 smooth, 10-12 diffuse, 13 .. 13+cd clip, 13+cd .. +nb omega.
// ------------------------------------------------------------------------------------------------
struct SmoothArgs {
    float* ch; const float* cj; const float* x; const float* xj;
    const int* count; uint32_t M, nflex, cd, nb, pred_clip;
    float r_xyz, r_color, r_clip;        // 1 / (bound^2 s_xyz), 1 / s_color, 1 / s_clip (0: term off)
    float* gate;
    float* g_ch; float* g_cj;            // backward: d loss / d ch (in: incoming gradient, out: with the smooth column folded in)
};

__global__ void __launch_bounds__(256) k_smooth_fwd(SmoothArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = a.count ? min((uint32_t)max(*a.count, 0), a.M) : a.M;
    if (i >= n) return;
    float* c = a.ch + (size_t)i * a.nflex;
    const float* j = a.cj + (size_t)i * a.nflex;
    float k = 0.f, d2 = 0.f;
#pragma unroll
    for (int q = 0; q < 3; q++) { const float e = a.x[(size_t)i * 3 + q] - a.xj[(size_t)i * 3 + q]; d2 += e * e; }
    k += d2 * a.r_xyz;
    d2 = 0.f;
#pragma unroll
    for (int q = 0; q < 3; q++) { const float e = c[10 + q] - j[10 + q]; d2 += e * e; }
    k += d2 * a.r_color;
    float sc = 0.f;
    if (a.pred_clip) {
        for (uint32_t q = 0; q < a.cd; q++) { const float e = j[13 + q] - c[13 + q]; sc += e * e; }
        if (a.r_clip > 0.f) k += sqrtf(sc) * a.r_clip;
    }
    const float g = __expf(-k);
    float so = 0.f;
    const uint32_t c0 = 13 + a.cd;
    for (uint32_t q = 0; q < a.nb; q++) { const float e = j[c0 + q] - c[c0 + q]; so += e * e; }
    a.gate[i] = g;
    c[3] = (so + sc) * g;
}

__global__ void __launch_bounds__(256) k_smooth_bwd(SmoothArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = a.count ? min((uint32_t)max(*a.count, 0), a.M) : a.M;
    if (i >= n) return;
    const float* c = a.ch + (size_t)i * a.nflex;
    const float* j = a.cj + (size_t)i * a.nflex;
    float* gc = a.g_ch + (size_t)i * a.nflex;
    float* gj = a.g_cj + (size_t)i * a.nflex;
    const float s = 2.f * gc[3] * a.gate[i];
    const uint32_t c0 = 13 + a.cd;
    for (uint32_t q = 0; q < a.nflex; q++) {
        const bool term = (q >= c0 && q < c0 + a.nb) || (a.pred_clip && q >= 13 && q < 13 + a.cd);
        const float d = term ? s * (j[q] - c[q]) : 0.f;
        gj[q] = d;
        if (term) gc[q] -= d;
    }
    gc[3] = 0.f;
}

}  // namespace pnerf

using namespace pnerf;

extern "C" {

uint32_t pnerf_palette_loss_partials(uint32_t N) { return ceil_div(N, (uint32_t)kLossRays) * (uint32_t)kLossTerms; }

int pnerf_palette_loss(const pnerf_palette_loss_args* p, void* stream) {
    PNERF_REQUIRE(p != nullptr);
    if (p->N == 0) return PNERF_ERR_INVALID_ARG;      // a mean over zero rays is undefined (torch returns nan)
    PNERF_REQUIRE(p->image && p->direct_rgb && p->gt_rgb && p->maps && p->terms && p->g_image && p->g_direct && p->g_maps);
    PNERF_REQUIRE(p->stride >= 1 && p->stride <= (uint32_t)kMaxStride && p->partials);
    const int cols[4] = {p->col_sparsity, p->col_offsets, p->col_view_dep, p->col_smooth};
    for (int c : cols) PNERF_REQUIRE(c < (int)p->stride);
    if (p->col_clip >= 0) PNERF_REQUIRE(p->gt_clip && p->clip_dim >= 1 && p->col_clip + p->clip_dim <= p->stride);
    if (p->col_basis >= 0) PNERF_REQUIRE(p->gt_weights && p->num_basis >= 1 && p->col_basis + p->num_basis <= p->stride);
    if (p->basis_color) PNERF_REQUIRE(p->basis_color_origin && p->g_basis_color && p->num_basis >= 1 && p->num_basis * 3 <= (uint32_t)kLossThreads);
    const uint32_t blocks = ceil_div(p->N, (uint32_t)kLossRays);
    LossArgs a;
    a.image = p->image; a.direct_rgb = p->direct_rgb; a.gt_rgb = p->gt_rgb; a.maps = p->maps; a.stride = p->stride;
    a.col_sparsity = p->col_sparsity; a.col_offsets = p->col_offsets; a.col_view_dep = p->col_view_dep;
    a.col_smooth = p->col_smooth; a.col_clip = p->col_clip; a.col_basis = p->col_basis;
    a.clip_dim = p->clip_dim; a.num_basis = p->num_basis; a.gt_clip = p->gt_clip; a.gt_weights = p->gt_weights;
    a.basis_color = p->basis_color; a.basis_color_origin = p->basis_color_origin;
    a.l_sparsity = p->lambda_sparsity; a.l_offsets = p->lambda_offsets; a.l_view_dep = p->lambda_view_dep;
    a.l_smooth = p->lambda_smooth; a.l_weight = p->lambda_weight; a.l_palette = p->lambda_palette;
    a.N = p->N; a.terms = p->terms; a.per_ray = p->per_ray;
    a.g_image = p->g_image; a.g_direct = p->g_direct; a.g_maps = p->g_maps; a.g_basis_color = p->g_basis_color;
    a.partials = p->partials;
    k_palette_loss<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(a);
    k_palette_loss_finish<<<1, 32 * kLossTerms, 0, (cudaStream_t)stream>>>(a, blocks);
    return check_launch("palette_loss");
}

int pnerf_scale_buffers(float* b0, uint32_t n0, float* b1, uint32_t n1, float* b2, uint32_t n2, float* b3, uint32_t n3,
                        const float* scale, void* stream) {
    PNERF_REQUIRE(scale != nullptr);
    ScaleArgs a;
    a.p[0] = n0 ? b0 : nullptr; a.p[1] = n1 ? b1 : nullptr; a.p[2] = n2 ? b2 : nullptr; a.p[3] = n3 ? b3 : nullptr;
    a.n[0] = n0; a.n[1] = n1; a.n[2] = n2; a.n[3] = n3; a.s = scale;
    const uint32_t most = max(max(n0, n1), max(n2, n3));
    if (most == 0) return PNERF_OK;
    const uint32_t blocks = min(ceil_div(most, 256u), 4u * (uint32_t)kNumSMs);
    k_scale_buffers<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("scale_buffers");
}

/* smooth-loss channel (palette/renderer.py:360-381): channels [M, nflex] (column 3 written in place), channels_j = the field at
 * the jittered positions, xyzs / xyzs_j [M,3], count (optional device int: valid rows), gate [M] saved for the backward */
int pnerf_palette_smooth_forward(float* channels, const float* channels_j, const float* xyzs, const float* xyzs_j, uint32_t M,
                                 const int32_t* count, uint32_t nflex, uint32_t clip_dim, uint32_t num_basis, uint32_t pred_clip,
                                 float bound, float sigma_xyz, float sigma_color, float sigma_clip, float* gate, void* stream) {
    PNERF_REQUIRE(channels && channels_j && xyzs && xyzs_j && gate);
    PNERF_REQUIRE(nflex >= 13 + clip_dim + num_basis && bound > 0.f && sigma_xyz > 0.f && sigma_color > 0.f);
    if (M == 0) return PNERF_OK;
    SmoothArgs a = {};
    a.ch = channels; a.cj = channels_j; a.x = xyzs; a.xj = xyzs_j; a.count = count; a.M = M; a.nflex = nflex; a.cd = clip_dim;
    a.nb = num_basis; a.pred_clip = pred_clip; a.r_xyz = 1.f / (bound * bound) / sigma_xyz; a.r_color = 1.f / sigma_color;
    a.r_clip = sigma_clip > 0.f ? 1.f / sigma_clip : 0.f; a.gate = gate;
    k_smooth_fwd<<<ceil_div(M, 256u), 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("palette_smooth_forward");
}

/* g_channels [M, nflex]: in = gradient of the loss w.r.t. the channel rows (column 3 = d loss / d smooth), out = the same with
 * the smooth term folded into the omega / clip columns and column 3 cleared; g_channels_j [M, nflex] out (valid rows only) */
int pnerf_palette_smooth_backward(float* g_channels, float* g_channels_j, const float* channels, const float* channels_j,
                                  const float* gate, uint32_t M, const int32_t* count, uint32_t nflex, uint32_t clip_dim,
                                  uint32_t num_basis, uint32_t pred_clip, void* stream) {
    PNERF_REQUIRE(g_channels && g_channels_j && channels && channels_j && gate && nflex >= 13 + clip_dim + num_basis);
    if (M == 0) return PNERF_OK;
    SmoothArgs a = {};
    a.ch = const_cast<float*>(channels); a.cj = channels_j; a.count = count; a.M = M; a.nflex = nflex; a.cd = clip_dim;
    a.nb = num_basis; a.pred_clip = pred_clip; a.gate = const_cast<float*>(gate); a.g_ch = g_channels; a.g_cj = g_channels_j;
    k_smooth_bwd<<<ceil_div(M, 256u), 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("palette_smooth_backward");
}

}  // extern "C"
