// nerf_train.cu — fused forward / backward kernels of the STAGE-1 field (NeRFNetwork) for TRAINING (B200, sm_100a).
//
// Replaces, for `sigmas, rgbs = self(xyzs, dirs)` of NeRFRenderer.run_cuda's training branch (nerf/renderer.py:289-298 ->
// nerf/network.py:78-124), what the reference runs as one hash-grid kernel, one SH kernel, 5 cuBLAS GEMMs and a dozen
// elementwise / cat / slice kernels forward plus their autograd graph backward:
//
//   k_nerf_train_fwd   hash grid (fp16 table) -> sigma net 32-64-16 -> sigma = density_scale * exp(h0), geo = h[1:16];
//                      SH(4) ++ geo -> colour net 31-64-64-3 -> sigmoid. 32 samples per warp on mma.sync.m16n8k16 with the
//                      weights resident in shared memory; every layer input is SAVED as the fp16 A fragments the warp
//                      already holds (train_common.cuh: 512-byte units, coalesced).
//   k_nerf_train_bwd   d rgb -> sigmoid' -> colour net backwards through the transposed weights -> d geo; d sigma ->
//                      trunc_exp' (exp(clamp(h0, -15, 15)), activation.py:14-17) -> column 0 of d h; sigma net backwards ->
//                      d(grid features) [M,32] fp32 for the run-length hash-grid backward (gridenc.cu). Emits the
//                      pre-activation gradient of every layer in fragment order (ybuf).
//   weight gradients   the split-K tensor-core kernel of the palette field (fused_train.cu::k_field_wgrad) on this field's
//                      job table.
// Unlike the palette stage, geometry IS trained here: gradients reach the sigma net and the density hash grid.
// Static-capacity mode as in fused_train.cu: with `m_dev` the kernels take the sample count from device memory.
#include "composite_common.cuh"
#include "train_common.cuh"

namespace pnerf {

// saved units per half-tile (16 samples)
enum NXSlot { NXS0 = 0, NXS1 = 2, NXV0 = 6, NXV1 = 8, NXV2 = 12, kNUX = 16 };
enum NYSlot { NYS0 = 0, NYS1 = 4, NYV0 = 5, NYV1 = 9, NYV2 = 13, kNUY = 14 };

// forward blob: B fragments [NT][KS][32] x uint2 per layer
enum NLayer { NL_S0, NL_S1, NL_V0, NL_V1, NL_V2, kNumNLayers };
__host__ __device__ constexpr int nl_ks(int l) { return l == NL_S0 ? 2 : l == NL_V0 ? 2 : 4; }
__host__ __device__ constexpr int nl_nt(int l) { return l == NL_S0 ? 8 : l == NL_S1 ? 2 : l == NL_V0 ? 8 : l == NL_V1 ? 8 : 1; }
__host__ __device__ constexpr int nl_off(int l) {
    int o = 0;
    for (int i = 0; i < l; i++) o += nl_ks(i) * nl_nt(i) * 32;
    return o;
}
constexpr int kNWUnits = nl_off(kNumNLayers);

// transposed blob (dX = dY W): V2^T 64x16, V1^T 64x64, (V0[:, 16:32])^T 16x64, S1^T 64x16, S0^T 32x64
enum NTLayer { NT_V2, NT_V1, NT_V0, NT_S1, NT_S0, kNumNTLayers };
__host__ __device__ constexpr int ntl_ks(int l) { return l == NT_V2 ? 1 : l == NT_S1 ? 1 : 4; }
__host__ __device__ constexpr int ntl_nt(int l) { return l == NT_V0 ? 2 : l == NT_S0 ? 4 : 8; }
__host__ __device__ constexpr int ntl_off(int l) {
    int o = 0;
    for (int i = 0; i < l; i++) o += ntl_ks(i) * ntl_nt(i) * 32;
    return o;
}
constexpr int kNTUnits = ntl_off(kNumNTLayers);

// packed fp32 weight-gradient buffer: S0 64x32, S1 16x64, V0 64x32, V1 64x64, V2 16x64
constexpr int kNDwS0 = 0, kNDwS1 = 2048, kNDwV0 = 3072, kNDwV1 = 5120, kNDwV2 = 9216, kNDwFloats = 10240;

constexpr int kNerfFwdWarps = 8;
constexpr float kExp15 = 3269017.372472110f, kExpM15 = 3.059023205018258e-07f;

struct NerfFwdScratch {
    __half feat[32][kFeatStride];
    float out[32][12];          // [0] logit, [1..3] rgb, [4..11] carried geo fragments (uint32)
};

__global__ void __launch_bounds__(kNerfFwdWarps * 32, 2)
k_nerf_train_fwd(const float* __restrict__ xyzs, const float* __restrict__ dirs, uint32_t M, pnerf_nerf_train f,
                 uint32_t* __restrict__ xbuf, float* __restrict__ sigma, float* __restrict__ rgb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TrainSmem* sm = reinterpret_cast<TrainSmem*>(smem_raw);
    uint2* wts = reinterpret_cast<uint2*>(smem_raw + ((sizeof(TrainSmem) + 15) & ~(size_t)15));
    NerfFwdScratch* scratch = reinterpret_cast<NerfFwdScratch*>(wts + kNWUnits);
    if (threadIdx.x < f.L) make_level(sm->lp[threadIdx.x], threadIdx.x, f.offsets, f.S, f.H, 3, 0, false);
    const int slow = __syncthreads_or(threadIdx.x < f.L && sm->lp[threadIdx.x].mask == 0u);
    if (threadIdx.x == 0) sm->fast_wrap = slow ? 0u : 1u;
    {
        const uint4* src = reinterpret_cast<const uint4*>(f.wfwd);
        uint4* dst = reinterpret_cast<uint4*>(wts);
        for (int i = threadIdx.x; i < kNWUnits / 2; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    NerfFwdScratch& ws = scratch[wid];
    if (f.m_dev) M = min(M, (uint32_t)__ldg(f.m_dev));
    const uint32_t n_tiles = ceil_div(M, 32u);
    const bool fast = sm->fast_wrap && f.L == 16;

    for (uint32_t tile = blockIdx.x * kNerfFwdWarps + wid; tile < n_tiles; tile += gridDim.x * kNerfFwdWarps) {
        const uint32_t s = tile * 32 + lane;
        const bool active = s < M;
        float x = 0, y = 0, z = 0, dx = 0, dy = 0, dz = 1;
        if (active) {
            x = xyzs[(size_t)s * 3]; y = xyzs[(size_t)s * 3 + 1]; z = xyzs[(size_t)s * 3 + 2];
            dx = dirs[(size_t)s * 3]; dy = dirs[(size_t)s * 3 + 1]; dz = dirs[(size_t)s * 3 + 2];
        }
        const float u = (x + f.bound) / (2 * f.bound), v = (y + f.bound) / (2 * f.bound), w = (z + f.bound) / (2 * f.bound);
        const bool in_range = active && !((u < 0 || u > 1) || (v < 0 || v > 1) || (w < 0 || w > 1));
        uint32_t* xtile = xbuf + (size_t)tile * 2 * kNUX * 128;
        uint32_t* carry = reinterpret_cast<uint32_t*>(&ws.out[lane][4]);   // [t][4]: this lane's geo fragment of half-tile t

        // ---- phase 1: density grid -> sigma net -> (logit, geo) ----
        if (fast) {
            __half (*rows)[kFeatStride] = ws.feat;      // 80-byte rows: a 4-level batch is one aligned 16-byte store
            auto st = [rows](int, int smp, int l0, const uint32_t (&wd)[4]) {
                *reinterpret_cast<uint4*>(&rows[smp][2 * l0]) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            };
            gather_coop<1, 4>(f.table, sm->lp, u, v, w, in_range, lane, st);
        } else {
            gather_features((const __half*)f.table, sm->lp, f.L, u, v, w, in_range, ws.feat[lane]);
        }
        __syncwarp();
#pragma unroll 1
        for (int t = 0; t < 2; t++) {
            uint32_t* xb = xtile + t * kNUX * 128;
            uint32_t a2[2][4];
            ldmatrix_a(a2[0], &ws.feat[0][0], 16 * t, 0, lane);
            ldmatrix_a(a2[1], &ws.feat[0][0], 16 * t, 16, lane);
            st_unit(xb, NXS0, a2[0], lane);
            st_unit(xb, NXS0 + 1, a2[1], lane);
            float c8[8][4];
            mma_layer<2, 8>(wts + nl_off(NL_S0), a2, c8, lane);
            uint32_t a4[4][4];
            chain<8, ACT_RELU>(c8, a4);
#pragma unroll
            for (int j = 0; j < 4; j++) st_unit(xb, NXS1 + j, a4[j], lane);
            float c2[2][4];
            mma_layer<4, 2>(wts + nl_off(NL_S1), a4, c2, lane);
            if ((lane & 3) == 0) {
                ws.out[16 * t + (lane >> 2)][0] = c2[0][0];
                ws.out[16 * t + (lane >> 2) + 8][0] = c2[0][2];
            }
            uint32_t geo[1][4];
            chain<2, ACT_NONE>(c2, geo);   // column 0 (the logit) rides along; the colour net's weights for it are zero
#pragma unroll
            for (int i = 0; i < 4; i++) carry[t * 4 + i] = geo[0][i];
        }
        __syncwarp();

        // ---- phase 2: SH(4) ++ geo -> colour net -> sigmoid ----
        {
            float sh[16];
            sh_eval<4, false>(dx, dy, dz, sh, nullptr, nullptr, nullptr);
#pragma unroll
            for (int i = 0; i < 8; i++) reinterpret_cast<__half2*>(ws.feat[lane])[i] = __floats2half2_rn(sh[2 * i], sh[2 * i + 1]);
        }
        __syncwarp();
#pragma unroll 1
        for (int t = 0; t < 2; t++) {
            uint32_t* xb = xtile + t * kNUX * 128;
            uint32_t a2[2][4];
            ldmatrix_a(a2[0], &ws.feat[0][0], 16 * t, 0, lane);
#pragma unroll
            for (int i = 0; i < 4; i++) a2[1][i] = carry[t * 4 + i];
            st_unit(xb, NXV0, a2[0], lane);
            st_unit(xb, NXV0 + 1, a2[1], lane);
            float c8[8][4];
            mma_layer<2, 8>(wts + nl_off(NL_V0), a2, c8, lane);
            uint32_t a4[4][4];
            chain<8, ACT_RELU>(c8, a4);
#pragma unroll
            for (int j = 0; j < 4; j++) st_unit(xb, NXV1 + j, a4[j], lane);
            mma_layer<4, 8>(wts + nl_off(NL_V1), a4, c8, lane);
            chain<8, ACT_RELU>(c8, a4);
#pragma unroll
            for (int j = 0; j < 4; j++) st_unit(xb, NXV2 + j, a4[j], lane);
            float c1[1][4];
            mma_layer<4, 1>(wts + nl_off(NL_V2), a4, c1, lane);
            const int r = lane >> 2, q = lane & 3;
            // exact sigmoid (the backward uses rgb (1 - rgb) of the stored value)
            if (2 * q < 3) { ws.out[16 * t + r][1 + 2 * q] = 1.0f / (1.0f + __expf(-c1[0][0])); ws.out[16 * t + r + 8][1 + 2 * q] = 1.0f / (1.0f + __expf(-c1[0][2])); }
            if (2 * q + 1 < 3) { ws.out[16 * t + r][2 + 2 * q] = 1.0f / (1.0f + __expf(-c1[0][1])); ws.out[16 * t + r + 8][2 + 2 * q] = 1.0f / (1.0f + __expf(-c1[0][3])); }
        }
        __syncwarp();
        if (active) sigma[s] = f.density_scale * __expf(ws.out[lane][0]);
        {
            const uint32_t rows = min(32u, M - tile * 32);
            float* ro = rgb + (size_t)tile * 96;
            for (uint32_t i = lane; i < rows * 3; i += 32) { const uint32_t r = i / 3; ro[i] = ws.out[r][1 + i - r * 3]; }
        }
        __syncwarp();
    }
}

// =====================================================================================================================
// backward (data gradients)
// =====================================================================================================================
constexpr int kNDStride = 24;      // halfs per row of the d(rgb pre-activation) staging tile (48 B: conflict-free ldmatrix rows)
struct NerfBwdScratch {
    __half d[32][kNDStride];
    float dlogit[32];
};

__global__ void __launch_bounds__(kTrainWarps * 32, 1)
k_nerf_train_bwd(uint32_t M, pnerf_nerf_train f, const uint32_t* __restrict__ xbuf, uint32_t* __restrict__ ybuf,
                 const float* __restrict__ grad_sigma, const float* __restrict__ grad_rgb, const float* __restrict__ sigma,
                 const float* __restrict__ rgb, float* __restrict__ d_enc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2* wt = reinterpret_cast<uint2*>(smem_raw);
    NerfBwdScratch* scratch = reinterpret_cast<NerfBwdScratch*>(wt + kNTUnits);
    {
        const uint4* src = reinterpret_cast<const uint4*>(f.wbwd);
        uint4* dst = reinterpret_cast<uint4*>(wt);
        for (int i = threadIdx.x; i < kNTUnits / 2; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    NerfBwdScratch& bs = scratch[wid];
    if (f.m_dev) M = min(M, (uint32_t)__ldg(f.m_dev));
    const uint32_t n_tiles = ceil_div(M, 32u);
    const float inv_ds = f.density_scale != 0.f ? 1.0f / f.density_scale : 0.f;

    for (uint32_t tile = blockIdx.x * kTrainWarps + wid; tile < n_tiles; tile += gridDim.x * kTrainWarps) {
        const uint32_t s = tile * 32 + lane;
        const bool active = s < M;
        const uint32_t* xtile = xbuf + (size_t)tile * 2 * kNUX * 128;
        uint32_t* ytile = ybuf + (size_t)tile * 2 * kNUY * 128;
        {
            __half* drow = bs.d[lane];
            float dl = 0.f;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float dv = 0.f;
                if (active) { const float o = rgb[(size_t)s * 3 + c]; dv = grad_rgb[(size_t)s * 3 + c] * o * (1.0f - o); }
                drow[c] = __float2half_rn(dv);
            }
#pragma unroll
            for (int c = 3; c < 16; c++) drow[c] = __float2half_rn(0.f);
            if (active) {
                // sigma_out = density_scale * exp(h0); d sigma_out / d h0 = density_scale * exp(clamp(h0, -15, 15))
                const float e = fminf(fmaxf(sigma[s] * inv_ds, kExpM15), kExp15);
                dl = grad_sigma[s] * f.density_scale * e;
            }
            bs.dlogit[lane] = dl;
        }
        __syncwarp();
#pragma unroll 1
        for (int t = 0; t < 2; t++) {
            const uint32_t* xb = xtile + t * kNUX * 128;
            uint32_t* yb = ytile + t * kNUY * 128;
            const uint32_t s0 = tile * 32 + 16 * t;
            float c8[8][4];
            uint32_t a4[4][4];
            uint32_t av[1][4];
            ldmatrix_a_s<kNDStride>(av[0], &bs.d[0][0], 16 * t, 0, lane);
            st_unit(yb, NYV2, av[0], lane);
            mma_layer<1, 8>(wt + ntl_off(NT_V2), av, c8, lane);
            deriv_pack<DRV_RELU>(c8, xb, NXV2, yb, NYV1, a4, lane);
            mma_layer<4, 8>(wt + ntl_off(NT_V1), a4, c8, lane);
            deriv_pack<DRV_RELU>(c8, xb, NXV1, yb, NYV0, a4, lane);
            float c2[2][4];
            mma_layer<4, 2>(wt + ntl_off(NT_V0), a4, c2, lane);        // d h [16]: column 0 gets nothing from the colour net
            if ((lane & 3) == 0) {
                c2[0][0] += bs.dlogit[16 * t + (lane >> 2)];
                c2[0][2] += bs.dlogit[16 * t + (lane >> 2) + 8];
            }
            uint32_t a1[1][4];
            chain<2, ACT_NONE>(c2, a1);
            st_unit(yb, NYS1, a1[0], lane);
            mma_layer<1, 8>(wt + ntl_off(NT_S1), a1, c8, lane);
            deriv_pack<DRV_RELU>(c8, xb, NXS1, yb, NYS0, a4, lane);
            float c4[4][4];
            mma_layer<4, 4>(wt + ntl_off(NT_S0), a4, c4, lane);
            store_denc(d_enc, s0, M, c4, lane);
        }
        __syncwarp();
    }
}

// =====================================================================================================================
// one-pass compositor of the stage-1 training step
// =====================================================================================================================
// The reference composites the training samples twice (nerf/renderer.py:301-327): composite_rays_train(sigmas, rgbs) and
// composite_rays_train(sigmas, rgb_norm) with rgb_norm_i = |gt(ray) - rgb_i|^2 built by spread_ray_to_sample + four
// elementwise kernels over [M,3]; its backward runs both compositors again plus the elementwise chain and needs four
// zero-filled [M, ...] gradient buffers. Here: ONE warp-per-ray pass forward (rgb, depth, weights_sum and the error channel,
// the error formed in registers from the ray's target colour) and ONE backward that writes d sigma (both compositors'
// contributions) and d rgb (image path + error path) for every sample of every ray that fits — zeros behind the
// terminating sample, so no buffer is memset. Per-sample arithmetic and termination rules: raymarching.cu:504-580, 681-761.
__global__ void __launch_bounds__(256)
k_nerf_comp_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                const int32_t* __restrict__ rays, const float* __restrict__ gt, uint32_t M, uint32_t N, float T_thresh,
                float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image, float* __restrict__ err_map) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1];
    const uint32_t num_steps = (uint32_t)rays[n * 3 + 2];
    float r = 0, g = 0, b = 0, ws = 0, d = 0, e = 0;
    if (num_steps != 0 && offset + num_steps <= M) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
        if (gt) { t0 = gt[index * 3]; t1 = gt[index * 3 + 1]; t2 = gt[index * 3 + 2]; }
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
                if (gt) e += w * ((t0 - c0) * (t0 - c0) + (t1 - c1) * (t1 - c1) + (t2 - c2) * (t2 - c2));
            }
            if (ct.last < 32u) break;
        }
        r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); ws = warp_sum(ws); d = warp_sum(d); e = warp_sum(e);
    }
    if (lane == 0) {
        weights_sum[index] = ws;
        depth[index] = d;
        image[index * 3 + 0] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
        err_map[index] = e;
    }
}

__global__ void __launch_bounds__(256)
k_nerf_comp_bwd(const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_image, const float* __restrict__ grad_err,
                const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                const int32_t* __restrict__ rays, const float* __restrict__ gt, const float* __restrict__ weights_sum,
                const float* __restrict__ image, const float* __restrict__ err_map, uint32_t M, uint32_t N, float T_thresh,
                float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1];
    const uint32_t num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) return;

    const float g0 = grad_image[index * 3 + 0], g1 = grad_image[index * 3 + 1], g2 = grad_image[index * 3 + 2];
    const float ge = (grad_err && gt) ? grad_err[index] : 0.f;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    if (gt) { t0 = gt[index * 3]; t1 = gt[index * 3 + 1]; t2 = gt[index * 3 + 2]; }
    const float r_final = image[index * 3 + 0], g_final = image[index * 3 + 1], b_final = image[index * 3 + 2];
    const float e_final = err_map[index];
    const float ws_term = (grad_weights_sum ? grad_weights_sum[index] : 0.f) * (1 - weights_sum[index]);

    float T = 1.0f, r_carry = 0.f, g_carry = 0.f, b_carry = 0.f, e_carry = 0.f;
    uint32_t base = 0;
    for (; base < num_steps; base += 32) {
        const uint32_t k = base + lane;
        const bool valid = k < num_steps;
        const size_t s = (size_t)offset + k;
        float alpha = 0.f, dt = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (valid) {
            dt = deltas[s * 2];
            alpha = 1.0f - __expf(-sigmas[s] * dt);
            c0 = rgbs[s * 3 + 0]; c1 = rgbs[s * 3 + 1]; c2 = rgbs[s * 3 + 2];
        }
        const float err = (t0 - c0) * (t0 - c0) + (t1 - c1) * (t1 - c1) + (t2 - c2) * (t2 - c2);
        const ChunkT ct = chunk_transmittance(alpha, valid, T, T_thresh, lane);
        const float w = alpha * ct.T_before;
        const float r_acc = r_carry + warp_scan_add(w * c0, lane);
        const float g_acc = g_carry + warp_scan_add(w * c1, lane);
        const float b_acc = b_carry + warp_scan_add(w * c2, lane);
        const float e_acc = e_carry + warp_scan_add(w * err, lane);
        r_carry = __shfl_sync(0xffffffffu, r_acc, 31);
        g_carry = __shfl_sync(0xffffffffu, g_acc, 31);
        b_carry = __shfl_sync(0xffffffffu, b_acc, 31);
        e_carry = __shfl_sync(0xffffffffu, e_acc, 31);
        if (valid) {
            const bool live = lane <= ct.last;
            const float gw = live ? w : 0.f;
            grad_rgbs[s * 3 + 0] = gw * (g0 + 2.0f * ge * (c0 - t0));
            grad_rgbs[s * 3 + 1] = gw * (g1 + 2.0f * ge * (c1 - t1));
            grad_rgbs[s * 3 + 2] = gw * (g2 + 2.0f * ge * (c2 - t2));
            grad_sigmas[s] = live ? dt * (g0 * (ct.T_after * c0 - (r_final - r_acc)) + g1 * (ct.T_after * c1 - (g_final - g_acc)) +
                                          g2 * (ct.T_after * c2 - (b_final - b_acc)) +
                                          ge * (ct.T_after * err - (e_final - e_acc)) + ws_term)
                                  : 0.f;
        }
        if (ct.last < 32u) { base += 32; break; }
    }
    for (uint32_t k = base + lane; k < num_steps; k += 32) {   // behind the terminating chunk: no gradient
        const size_t s = (size_t)offset + k;
        grad_rgbs[s * 3 + 0] = 0.f; grad_rgbs[s * 3 + 1] = 0.f; grad_rgbs[s * 3 + 2] = 0.f;
        grad_sigmas[s] = 0.f;
    }
}

}  // namespace pnerf

using namespace pnerf;

extern "C" {

uint64_t pnerf_nerf_train_xbuf_bytes(uint32_t M) { return (uint64_t)ceil_div(M, 32u) * 2 * kNUX * 512; }
uint64_t pnerf_nerf_train_ybuf_bytes(uint32_t M) { return (uint64_t)ceil_div(M, 32u) * 2 * kNUY * 512; }
uint32_t pnerf_nerf_train_dw_floats(void) { return kNDwFloats; }
uint32_t pnerf_nerf_train_wfwd_units(void) { return kNWUnits; }
uint32_t pnerf_nerf_train_wbwd_units(void) { return kNTUnits; }

static int nerf_args_ok(const pnerf_nerf_train* p) {
    if (!p || !p->table || !p->offsets || !p->wfwd || !p->wbwd) return PNERF_ERR_INVALID_ARG;
    if (p->L != 16) return PNERF_ERR_UNSUPPORTED;
    return PNERF_OK;
}

int pnerf_nerf_train_forward(const float* xyzs, const float* dirs, uint32_t M, const pnerf_nerf_train* p, void* xbuf, float* sigma,
                             float* rgb, void* stream) {
    if (M == 0) return PNERF_OK;
    PNERF_REQUIRE(xyzs && dirs && xbuf && sigma && rgb);
    if (int st = nerf_args_ok(p)) return st;
    const size_t smem = ((sizeof(TrainSmem) + 15) & ~(size_t)15) + (size_t)kNWUnits * sizeof(uint2) + sizeof(NerfFwdScratch) * kNerfFwdWarps;
    const uint32_t grid = min(ceil_div(ceil_div(M, 32u), (uint32_t)kNerfFwdWarps), 2u * (uint32_t)kNumSMs);
    static bool attr_done = false;   // once per process (keeps cudaFuncSetAttribute out of graph capture)
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_nerf_train_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_cuda_error(e, "nerf_train_forward attr"); return PNERF_ERR_CUDA; }
        attr_done = true;
    }
    k_nerf_train_fwd<<<grid, kNerfFwdWarps * 32, smem, (cudaStream_t)stream>>>(xyzs, dirs, M, *p, (uint32_t*)xbuf, sigma, rgb);
    return check_launch("nerf_train_forward");
}

int pnerf_nerf_train_backward(uint32_t M, const pnerf_nerf_train* p, const void* xbuf, void* ybuf, const float* grad_sigma,
                              const float* grad_rgb, const float* sigma, const float* rgb, float* d_enc, void* stream) {
    if (M == 0) return PNERF_OK;
    PNERF_REQUIRE(xbuf && ybuf && grad_sigma && grad_rgb && sigma && rgb && d_enc);
    if (int st = nerf_args_ok(p)) return st;
    const size_t smem = (size_t)kNTUnits * sizeof(uint2) + sizeof(NerfBwdScratch) * kTrainWarps;
    const uint32_t grid = min(ceil_div(ceil_div(M, 32u), (uint32_t)kTrainWarps), 2u * (uint32_t)kNumSMs);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_nerf_train_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_cuda_error(e, "nerf_train_backward attr"); return PNERF_ERR_CUDA; }
        attr_done = true;
    }
    k_nerf_train_bwd<<<grid, kTrainWarps * 32, smem, (cudaStream_t)stream>>>(M, *p, (const uint32_t*)xbuf, (uint32_t*)ybuf, grad_sigma,
                                                                            grad_rgb, sigma, rgb, d_enc);
    return check_launch("nerf_train_backward");
}

int pnerf_nerf_train_wgrad(uint32_t M, const void* xbuf, const void* ybuf, float* dwbuf, const int32_t* m_dev, void* stream) {
    if (M == 0) return PNERF_OK;
    PNERF_REQUIRE(xbuf && ybuf && dwbuf);
    WJobs J;
    J.n = 0;
    add_job_at(J, NYS0, 4, NXS0, 2, 32, kNDwS0);
    add_job_at(J, NYS1, 1, NXS1, 4, 64, kNDwS1);
    add_job_at(J, NYV0, 4, NXV0, 2, 32, kNDwV0);
    add_job_at(J, NYV1, 4, NXV1, 4, 64, kNDwV1);
    add_job_at(J, NYV2, 1, NXV2, 4, 64, kNDwV2);
    return launch_field_wgrad((const uint32_t*)xbuf, (const uint32_t*)ybuf, M, m_dev, kNUX, kNUY, J, dwbuf, (cudaStream_t)stream,
                              "nerf_train_wgrad");
}

int pnerf_nerf_composite_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                       const float* gt, uint32_t M, uint32_t N, float T_thresh, float* weights_sum, float* depth,
                                       float* image, float* err_map, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(sigmas && rgbs && deltas && rays && weights_sum && depth && image && err_map);
    k_nerf_comp_fwd<<<ceil_div(N, 8u), 256, 0, (cudaStream_t)stream>>>(sigmas, rgbs, deltas, rays, gt, M, N, T_thresh, weights_sum,
                                                                      depth, image, err_map);
    return check_launch("nerf_composite_train_forward");
}

int pnerf_nerf_composite_train_backward(const float* grad_weights_sum, const float* grad_image, const float* grad_err,
                                        const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                        const float* gt, const float* weights_sum, const float* image, const float* err_map,
                                        uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas, float* grad_rgbs, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(grad_image && sigmas && rgbs && deltas && rays && weights_sum && image && err_map && grad_sigmas && grad_rgbs);
    k_nerf_comp_bwd<<<ceil_div(N, 8u), 256, 0, (cudaStream_t)stream>>>(grad_weights_sum, grad_image, grad_err, sigmas, rgbs, deltas,
                                                                      rays, gt, weights_sum, image, err_map, M, N, T_thresh,
                                                                      grad_sigmas, grad_rgbs);
    return check_launch("nerf_composite_train_backward");
}

}  // extern "C"
