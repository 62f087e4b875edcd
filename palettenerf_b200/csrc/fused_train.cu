// fused_train.cu — fused forward / backward / weight-gradient kernels of the palette field for TRAINING (B200, sm_100a).
//
// Replaces, for the palette stage (palette/renderer.py:322-429), what the reference runs as ~14 cuBLAS GEMMs and ~60
// elementwise / cat / detach kernels forward plus their autograd graph backward (palette/network.py:156-280):
//
//   k_field_train_fwd   hash grids -> sigma / diffuse / view-dependent / basis nets -> heads -> training blend and the
//                       per-sample regulariser channels, 32 samples per warp on mma.sync.m16n8k16, weights resident in
//                       shared memory. Writes sigma [M], rgb [M,3], flex [M, 13+clip+Nb] and SAVES every layer input as
//                       the fp16 A fragments it already holds in registers: xbuf[half-tile][unit][4 regs][32 lanes]
//                       (unit = one m16 x k16 block) — 128-byte coalesced stores, no shared-memory transposes.
//   k_field_train_bwd   data gradients: the blend / regulariser / softplus-normalise / sigmoid derivatives are evaluated
//                       by the lane that owns the sample, staged as fp16 rows, and chained backwards through the
//                       TRANSPOSED weights with the same register-chained MMA scheme; activation derivatives come from
//                       the saved fragments (C-fragment layout == A-fragment layout). Emits the pre-activation gradient
//                       of every layer in fragment order (ybuf) and d(loss)/d(grid features) [M,32] fp32 for the
//                       palette (and semantic) hash grid, which go to the run-length hash-grid backward (gridenc.cu).
//   k_field_wgrad       all weight gradients dW = dY^T X as a tensor-core split-K over the samples: a warp owns one
//                       layer and a chunk of half-tiles, transposes the saved fragments with movmatrix (both operands
//                       are exactly the transposes of what was stored), prefetches the next half-tile while the MMAs of
//                       the current one run, and adds its partial sum once with fp32 red.global.add. Every saved
//                       fragment is read exactly once.
// Static-capacity mode: when `m_dev` is set the kernels take the sample count from device memory (the march kernel's
// counter), so a training step needs no host synchronisation and can be captured in a CUDA graph.
//
// The reference's detach() placements are honoured: sigma and geo features are constants in this stage (no sigma-net or
// sigma-grid gradient; palette/network.py:168, palette/renderer.py:334-335), diffuse enters the basis net detached
// (network.py:257) and view_dep enters rgb detached (renderer.py:351); view_dep / diffuse are trained through
// direct_rgb and the regularisers. The smooth-loss branch (renderer.py:360-381) stays on the unfused path.
#include "train_common.cuh"

namespace pnerf {

// ---- saved-activation units per half-tile (16 samples) ------------------------------------------------------------
enum XSlot { XD0 = 0, XD1 = 1, XD2 = 5, XV0 = 9, XV1 = 11, XV2 = 15, XB0 = 19, XB1 = 22, XH = 26, XC0 = 27, XC1 = 29 };
constexpr int kUXNoClip = 27, kUXClip = 33;
enum YSlot { YD0 = 0, YD1 = 4, YD2 = 8, YV0 = 9, YV1 = 13, YV2 = 17, YB0 = 18, YB1 = 22, YH = 23, YC0 = 25, YC1 = 29 };
constexpr int kUYNoClip = 25, kUYClip = 30;

// ---- transposed-weight blob (dX = dY W): per layer [NT][KS][32] x uint2, same fragment order as the forward blob ----
enum TLayer { T_H, T_B1, T_B0, T_D2, T_D1, T_V2, T_V1, T_C1, T_C0, kNumTLayers };
__host__ __device__ constexpr int tl_ks(int l) { return l == T_H ? 2 : l == T_B1 ? 1 : l == T_B0 ? 4 : l == T_D2 ? 1 : l == T_D1 ? 4 : l == T_V2 ? 1 : l == T_V1 ? 4 : l == T_C1 ? 1 : 4; }
__host__ __device__ constexpr int tl_nt(int l) { return l == T_H ? 2 : l == T_B1 ? 8 : l == T_B0 ? 4 : l == T_D2 ? 8 : l == T_D1 ? 8 : l == T_V2 ? 8 : l == T_V1 ? 8 : l == T_C1 ? 8 : 4; }
__host__ __device__ constexpr int tl_off(int l) {
    int o = 0;
    for (int i = 0; i < l; i++) o += tl_ks(i) * tl_nt(i) * 32;
    return o;
}
constexpr int kTUnitsNoClip = tl_off(T_C1);
constexpr int kTUnitsClip = tl_off(kNumTLayers);

// ---- packed fp32 weight-gradient buffer: per layer [N_pad][K_pad] row-major ------------------------------------------
enum DwLayer { DW_D0, DW_D1, DW_D2, DW_V0, DW_V1, DW_V2, DW_B0, DW_B1, DW_H, DW_C0, DW_C1, kNumDw };
__host__ __device__ constexpr int dw_n(int l) { return l == DW_D2 || l == DW_V2 || l == DW_B1 || l == DW_C1 ? 16 : l == DW_H ? 32 : 64; }
__host__ __device__ constexpr int dw_k(int l) { return l == DW_D0 ? 16 : l == DW_V0 ? 32 : l == DW_B0 ? 48 : l == DW_H ? 16 : l == DW_C0 ? 32 : 64; }
__host__ __device__ constexpr int dw_off(int l) {
    int o = 0;
    for (int i = 0; i < l; i++) o += dw_n(i) * dw_k(i);
    return o;
}
constexpr int kDwFloatsNoClip = dw_off(DW_C0);
constexpr int kDwFloatsClip = dw_off(kNumDw);

// =====================================================================================================================
// forward
// =====================================================================================================================
template <bool CLIP>
__global__ void __launch_bounds__(kFusedWarps * 32, 1)
k_field_train_fwd(const float* __restrict__ xyzs, const float* __restrict__ dirs, uint32_t M, pnerf_palette_train f,
                  uint32_t* __restrict__ xbuf, float* __restrict__ sigma, float* __restrict__ rgb, float* __restrict__ flex) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TrainSmem* sm = reinterpret_cast<TrainSmem*>(smem_raw);
    uint2* wts = reinterpret_cast<uint2*>(smem_raw + ((sizeof(TrainSmem) + 15) & ~(size_t)15));
    constexpr int kWU = CLIP ? kWUnitsClip : kWUnitsNoClip;
    WarpScratch* scratch = reinterpret_cast<WarpScratch*>(wts + kWU);
    if (threadIdx.x < f.L) make_level(sm->lp[threadIdx.x], threadIdx.x, f.offsets, f.S, f.H, 3, 0, false);
    const int slow = __syncthreads_or(threadIdx.x < f.L && sm->lp[threadIdx.x].mask == 0u);
    if (threadIdx.x == 0) sm->fast_wrap = slow ? 0u : 1u;
    if (threadIdx.x < kNB * 3) sm->palette[threadIdx.x] = f.palette[threadIdx.x];
    {
        const uint4* src = reinterpret_cast<const uint4*>(f.wfwd);
        uint4* dst = reinterpret_cast<uint4*>(wts);
        for (int i = threadIdx.x; i < kWU / 2; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpScratch& ws = scratch[wid];
    constexpr int UX = CLIP ? kUXClip : kUXNoClip;
    const uint32_t cd = f.clip_dim, nflex = 13 + cd + kNB;
    if (f.m_dev) M = min(M, (uint32_t)__ldg(f.m_dev));      // static-capacity mode: the sample count lives on the device
    const uint32_t n_tiles = ceil_div(M, 32u);

    for (uint32_t tile = blockIdx.x * kFusedWarps + wid; tile < n_tiles; tile += gridDim.x * kFusedWarps) {
        const uint32_t s = tile * 32 + lane;
        const bool active = s < M;
        float x = 0, y = 0, z = 0, dx = 0, dy = 0, dz = 1;
        if (active) {
            x = xyzs[(size_t)s * 3]; y = xyzs[(size_t)s * 3 + 1]; z = xyzs[(size_t)s * 3 + 2];
            dx = dirs[(size_t)s * 3]; dy = dirs[(size_t)s * 3 + 1]; dz = dirs[(size_t)s * 3 + 2];
        }
        const float u = (x + f.bound) / (2 * f.bound), v = (y + f.bound) / (2 * f.bound), w = (z + f.bound) / (2 * f.bound);
        const bool in_range = active && !((u < 0 || u > 1) || (v < 0 || v > 1) || (w < 0 || w > 1));
        uint32_t* xtile = xbuf + (size_t)tile * 2 * UX * 128;
        uint32_t* carry = reinterpret_cast<uint32_t*>(&ws.out[lane][O_CLIP]);   // [t][6], see fused.cu::eval_field

        // ---- phase 1: density grid -> sigma net -> geo; geo -> diffuse net ----
        // The palette grid shares the density grid's geometry: with the two tables interleaved entry by entry
        // (table_sigma_palette) both are read here with ONE set of corner indices and one 8-byte load per corner; the
        // palette features wait (fp16 pairs) in this lane's output row, columns O_OFFRAD.. that phase 3 writes last
        // (same scheme as fused.cu::eval_field).
        uint32_t* park = reinterpret_cast<uint32_t*>(&ws.out[lane][O_OFFRAD]);   // 16 words
        const bool paired = sm->fast_wrap && f.table_sigma_palette != nullptr && f.L == 16;     // warp-uniform
        if (paired) {
            uint32_t* const rows[2] = {reinterpret_cast<uint32_t*>(ws.feat[lane]), park};
            gather_fast<2, 2>(f.table_sigma_palette, sm->lp, u, v, w, in_range, rows);
        } else {
            gather_features((const __half*)f.table_sigma, sm->lp, f.L, u, v, w, in_range, ws.feat[lane]);
        }
        __syncwarp();
#pragma unroll 1
        for (int t = 0; t < 2; t++) {
            uint32_t* xb = xtile + t * UX * 128;
            uint32_t a2[2][4];
            ldmatrix_a(a2[0], &ws.feat[0][0], 16 * t, 0, lane);
            ldmatrix_a(a2[1], &ws.feat[0][0], 16 * t, 16, lane);
            float c8[8][4];
            mma_layer<2, 8>(wts + layer_off(LS0), a2, c8, lane);
            uint32_t a4[4][4];
            chain<8, ACT_RELU>(c8, a4);
            float c2[2][4];
            mma_layer<4, 2>(wts + layer_off(LS1), a4, c2, lane);
            if ((lane & 3) == 0) {
                ws.out[16 * t + (lane >> 2)][O_SIGMA] = c2[0][0];
                ws.out[16 * t + (lane >> 2) + 8][O_SIGMA] = c2[0][2];
            }
            uint32_t geo[1][4];
            chain<2, ACT_NONE>(c2, geo);
            st_unit(xb, XD0, geo[0], lane);
#pragma unroll
            for (int i = 0; i < 4; i++) carry[t * 6 + i] = geo[0][i];
            mma_layer<1, 8>(wts + layer_off(LD0), geo, c8, lane);
            chain<8, ACT_RELU>(c8, a4);
#pragma unroll
            for (int j = 0; j < 4; j++) st_unit(xb, XD1 + j, a4[j], lane);
            mma_layer<4, 8>(wts + layer_off(LD1), a4, c8, lane);
            chain<8, ACT_RELU>(c8, a4);
#pragma unroll
            for (int j = 0; j < 4; j++) st_unit(xb, XD2 + j, a4[j], lane);
            float c1[1][4];
            mma_layer<4, 1>(wts + layer_off(LD2), a4, c1, lane);
#pragma unroll
            for (int i = 0; i < 4; i++) c1[0][i] = sigmoidf_(c1[0][i]);
            store_out<1>(ws.out, 16 * t, O_DIFF, 3, c1, lane);
            carry[t * 6 + 4] = pack_h2(c1[0][0], c1[0][1]);
            carry[t * 6 + 5] = pack_h2(c1[0][2], c1[0][3]);
        }
        __syncwarp();

        // ---- phase 2: SH(4) ++ geo -> view-dependent colour net ----
        {
            float sh[16];
            sh_eval<4, false>(dx, dy, dz, sh, nullptr, nullptr, nullptr);
#pragma unroll
            for (int i = 0; i < 8; i++) reinterpret_cast<__half2*>(ws.feat[lane])[i] = __floats2half2_rn(sh[2 * i], sh[2 * i + 1]);
        }
        __syncwarp();
#pragma unroll 1
        for (int t = 0; t < 2; t++) {
            uint32_t* xb = xtile + t * UX * 128;
            uint32_t a2[2][4];
            ldmatrix_a(a2[0], &ws.feat[0][0], 16 * t, 0, lane);
#pragma unroll
            for (int i = 0; i < 4; i++) a2[1][i] = carry[t * 6 + i];
            st_unit(xb, XV0, a2[0], lane);
            st_unit(xb, XV0 + 1, a2[1], lane);
            float c8[8][4];
            mma_layer<2, 8>(wts + layer_off(LV0), a2, c8, lane);
            uint32_t a4[4][4];
            chain<8, ACT_RELU>(c8, a4);
#pragma unroll
            for (int j = 0; j < 4; j++) st_unit(xb, XV1 + j, a4[j], lane);
            mma_layer<4, 8>(wts + layer_off(LV1), a4, c8, lane);
            chain<8, ACT_RELU>(c8, a4);
#pragma unroll
            for (int j = 0; j < 4; j++) st_unit(xb, XV2 + j, a4[j], lane);
            float c1[1][4];
            mma_layer<4, 1>(wts + layer_off(LV2), a4, c1, lane);
#pragma unroll
            for (int i = 0; i < 4; i++) c1[0][i] = sigmoidf_(c1[0][i]);
            store_out<1>(ws.out, 16 * t, O_VIEW, 3, c1, lane);
        }
        __syncwarp();

        // ---- phase 3: palette grid ++ diffuse(detached) -> basis net -> heads (bias folded into column 15) ----
        if (paired) {
#pragma unroll
            for (int i = 0; i < 16; i++) reinterpret_cast<uint32_t*>(ws.feat[lane])[i] = park[i];
        } else {
            gather_features((const __half*)f.table_palette, sm->lp, f.L, u, v, w, in_range, ws.feat[lane]);
        }
        __syncwarp();
#pragma unroll 1
        for (int t = 0; t < 2; t++) {
            uint32_t* xb = xtile + t * UX * 128;
            uint32_t a3[3][4];
            ldmatrix_a(a3[0], &ws.feat[0][0], 16 * t, 0, lane);
            ldmatrix_a(a3[1], &ws.feat[0][0], 16 * t, 16, lane);
            a3[2][0] = carry[t * 6 + 4];
            a3[2][1] = carry[t * 6 + 5];
            a3[2][2] = 0u;
            a3[2][3] = 0u;
#pragma unroll
            for (int j = 0; j < 3; j++) st_unit(xb, XB0 + j, a3[j], lane);
            float c8[8][4];
            mma_layer<3, 8>(wts + layer_off(LB0), a3, c8, lane);
            uint32_t a4[4][4];
            chain<8, ACT_ELU>(c8, a4);
#pragma unroll
            for (int j = 0; j < 4; j++) st_unit(xb, XB1 + j, a4[j], lane);
            float c2[2][4];
            mma_layer<4, 2>(wts + layer_off(LB1), a4, c2, lane);
            uint32_t a1[1][4];
            chain<2, ACT_NONE>(c2, a1);
            if ((lane & 3) == 3) {   // column 15 of the 16-wide head input := 1.0 (fp16 0x3C00) -> the head bias is a weight
                a1[0][2] = (a1[0][2] & 0x0000ffffu) | 0x3C000000u;
                a1[0][3] = (a1[0][3] & 0x0000ffffu) | 0x3C000000u;
            }
            st_unit(xb, XH, a1[0], lane);
            float c3[3][4];
            mma_layer<1, 3>(wts + layer_off(LH), a1, c3, lane);
            store_out<3>(ws.out, 16 * t, O_OFFRAD, 13 + kNB, c3, lane);
        }
        __syncwarp();

        // ---- phase 4 (optional): semantic grid -> clip net ----
        if (CLIP) {
            gather_features((const __half*)f.table_clip, sm->lp, f.L, u, v, w, in_range, ws.feat[lane]);
            __syncwarp();
#pragma unroll 1
            for (int t = 0; t < 2; t++) {
                uint32_t* xb = xtile + t * UX * 128;
                uint32_t a2[2][4];
                ldmatrix_a(a2[0], &ws.feat[0][0], 16 * t, 0, lane);
                ldmatrix_a(a2[1], &ws.feat[0][0], 16 * t, 16, lane);
                st_unit(xb, XC0, a2[0], lane);
                st_unit(xb, XC0 + 1, a2[1], lane);
                float c8[8][4];
                mma_layer<2, 8>(wts + layer_off(LC0), a2, c8, lane);
                uint32_t a4[4][4];
                chain<8, ACT_RELU>(c8, a4);
#pragma unroll
                for (int j = 0; j < 4; j++) st_unit(xb, XC1 + j, a4[j], lane);
                float c2[2][4];
                mma_layer<4, 2>(wts + layer_off(LC1), a4, c2, lane);
                store_out<2>(ws.out, 16 * t, O_CLIP, (int)cd, c2, lane);
            }
            __syncwarp();
        }

        // ---- this lane's sample: training blend + regulariser channels (ref: palette/renderer.py:333-359, 384-385) ----
        float* row = ws.out[lane];
        const float sig = f.density_scale * __expf(row[O_SIGMA]);
        float diffuse[3], view_dep[3], off_rad[13], omega[kNB], clipv[kClipMax];
#pragma unroll
        for (int i = 0; i < 3; i++) { diffuse[i] = row[O_DIFF + i]; view_dep[i] = row[O_VIEW + i]; }
#pragma unroll
        for (int i = 0; i < 13; i++) off_rad[i] = row[O_OFFRAD + i];
        float osum = 0.f;
#pragma unroll
        for (int b = 0; b < kNB; b++) { omega[b] = softplusf_(row[O_OMEGA + b]) + 0.05f; osum += omega[b]; }
        const float rinv = 1.0f / osum;
#pragma unroll
        for (int b = 0; b < kNB; b++) omega[b] *= rinv;
#pragma unroll
        for (int i = 0; i < kClipMax; i++) clipv[i] = (CLIP && i < (int)cd) ? row[O_CLIP + i] : 0.f;
        __syncwarp();

        const float sp = softplusf_(off_rad[12]);
        float col[3] = {0.f, 0.f, 0.f}, s1 = 0.f, s2 = 0.f, onorm = 0.f, vnorm = 0.f;
#pragma unroll
        for (int b = 0; b < kNB; b++) {
            s1 += omega[b]; s2 += omega[b] * omega[b];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float off = off_rad[b * 3 + c];
                col[c] += omega[b] * (sp * (sm->palette[b * 3 + c] + off));
                onorm += off * off;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) { col[c] += view_dep[c]; vnorm += view_dep[c] * view_dep[c]; }
        // stage the flex row + rgb in this lane's output row, then copy out coalesced
        row[0] = s1 / (s2 + 1e-6f) - 1.0f;
        row[1] = vnorm;
        row[2] = onorm;
        row[3] = 0.f;                                    // smooth_norm: unfused path only
#pragma unroll
        for (int c = 0; c < 3; c++) { row[4 + c] = view_dep[c]; row[7 + c] = diffuse[c] + view_dep[c]; row[10 + c] = diffuse[c]; }
        for (uint32_t i = 0; i < cd; i++) row[13 + i] = i < (uint32_t)kClipMax ? clipv[i] : 0.f;
#pragma unroll
        for (int b = 0; b < kNB; b++) row[13 + cd + b] = omega[b];
        row[37] = col[0]; row[38] = col[1]; row[39] = col[2];
        if (active) sigma[s] = sig;
        __syncwarp();
        {
            const uint32_t rows = min(32u, M - tile * 32);
            float* fo = flex + (size_t)tile * 32 * nflex;
            for (uint32_t i = lane; i < rows * nflex; i += 32) { const uint32_t r = i / nflex; fo[i] = ws.out[r][i - r * nflex]; }
            float* ro = rgb + (size_t)tile * 96;
            for (uint32_t i = lane; i < rows * 3; i += 32) { const uint32_t r = i / 3; ro[i] = ws.out[r][37 + i - r * 3]; }
        }
        __syncwarp();
    }
}

// =====================================================================================================================
// backward (data gradients)
// =====================================================================================================================
struct BwdScratch {
    float out[32][kOutStride];
    __half d[32][kDStride];
};

template <bool CLIP>
__global__ void __launch_bounds__(kTrainWarps * 32, 1)
k_field_train_bwd(uint32_t M, pnerf_palette_train f, const uint32_t* __restrict__ xbuf, uint32_t* __restrict__ ybuf,
                  const float* __restrict__ grad_rgb, const float* __restrict__ grad_flex, const float* __restrict__ flex,
                  float* __restrict__ d_enc, float* __restrict__ d_enc_clip, float* __restrict__ d_palette) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* palette = reinterpret_cast<float*>(smem_raw);                       // [12] (+4 pad)
    uint2* wh = reinterpret_cast<uint2*>(smem_raw + 64);                        // forward head layer: 3 n-tiles x 1 k-step
    uint2* wt = wh + 3 * 32;                                                    // transposed layers
    constexpr int kTU = CLIP ? kTUnitsClip : kTUnitsNoClip;
    BwdScratch* scratch = reinterpret_cast<BwdScratch*>(wt + kTU);
    if (threadIdx.x < kNB * 3) palette[threadIdx.x] = f.palette[threadIdx.x];
    if (threadIdx.x < 96) wh[threadIdx.x] = reinterpret_cast<const uint2*>(f.wfwd)[layer_off(LH) + threadIdx.x];
    {
        const uint4* src = reinterpret_cast<const uint4*>(f.wbwd);
        uint4* dst = reinterpret_cast<uint4*>(wt);
        for (int i = threadIdx.x; i < kTU / 2; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    BwdScratch& bs = scratch[wid];
    constexpr int UX = CLIP ? kUXClip : kUXNoClip, UY = CLIP ? kUYClip : kUYNoClip;
    const uint32_t cd = f.clip_dim, nflex = 13 + cd + kNB;
    if (f.m_dev) M = min(M, (uint32_t)__ldg(f.m_dev));
    const uint32_t n_tiles = ceil_div(M, 32u);

    for (uint32_t tile = blockIdx.x * kTrainWarps + wid; tile < n_tiles; tile += gridDim.x * kTrainWarps) {
        const uint32_t s = tile * 32 + lane;
        const bool active = s < M;
        const uint32_t* xtile = xbuf + (size_t)tile * 2 * UX * 128;
        uint32_t* ytile = ybuf + (size_t)tile * 2 * UY * 128;

        // ---- recompute the head outputs (one k-step MMA on the saved head input) ----
#pragma unroll 1
        for (int t = 0; t < 2; t++) {
            uint32_t a1[1][4];
            ld_unit(xtile + t * UX * 128, XH, a1[0], lane);
            float c3[3][4];
            mma_layer<1, 3>(wh, a1, c3, lane);
            store_out<3>(bs.out, 16 * t, 0, 13 + kNB, c3, lane);
        }
        __syncwarp();

        // ---- owner-lane math: blend, regularisers, softplus-normalise and sigmoid derivatives ----
        {
            const float* row = bs.out[lane];
            float off[13], zz[kNB], u[kNB], om[kNB], usum = 0.f;
#pragma unroll
            for (int i = 0; i < 13; i++) off[i] = row[i];
#pragma unroll
            for (int b = 0; b < kNB; b++) { zz[b] = row[13 + b]; u[b] = softplusf_(zz[b]) + 0.05f; usum += u[b]; }
            const float rinv = 1.0f / usum;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int b = 0; b < kNB; b++) { om[b] = u[b] * rinv; s1 += om[b]; s2 += om[b] * om[b]; }
            const float sp = softplusf_(off[12]), sgr = sigmoidf_(off[12]);
            float g_rgb[3] = {0.f, 0.f, 0.f}, g_sp = 0.f, g_vn = 0.f, g_on = 0.f;
            float g_vd[3] = {0.f, 0.f, 0.f}, g_dir[3] = {0.f, 0.f, 0.f}, g_df[3] = {0.f, 0.f, 0.f}, g_om[kNB] = {0.f, 0.f, 0.f, 0.f};
            float vd[3] = {0.f, 0.f, 0.f}, df[3] = {0.f, 0.f, 0.f};
            __half* drow = bs.d[lane];
            if (active) {
                const float* gf = grad_flex + (size_t)s * nflex;
                const float* fv = flex + (size_t)s * nflex;
                g_sp = gf[0]; g_vn = gf[1]; g_on = gf[2];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    g_rgb[c] = grad_rgb[(size_t)s * 3 + c];
                    g_vd[c] = gf[4 + c]; g_dir[c] = gf[7 + c]; g_df[c] = gf[10 + c];
                    vd[c] = fv[4 + c]; df[c] = fv[10 + c];
                }
#pragma unroll
                for (int b = 0; b < kNB; b++) g_om[b] = gf[13 + cd + b];
                for (uint32_t i = 0; i < 16; i++) drow[DC_CLIP + i] = __float2half_rn((CLIP && i < cd) ? gf[13 + i] : 0.f);
            } else {
                for (uint32_t i = 0; i < 16; i++) drow[DC_CLIP + i] = __float2half_rn(0.f);
            }
            const float inv = 1.0f / (s2 + 1e-6f);
            float d_om[kNB], d_rad = 0.f, dot = 0.f, dpal[kNB * 3];
#pragma unroll
            for (int b = 0; b < kNB; b++) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float colv = palette[b * 3 + c] + off[b * 3 + c];
                    acc += g_rgb[c] * colv;
                    dpal[b * 3 + c] = g_rgb[c] * om[b] * sp;
                    drow[DC_HEAD + b * 3 + c] = __float2half_rn(dpal[b * 3 + c] + g_on * 2.0f * off[b * 3 + c]);
                }
                d_om[b] = g_om[b] + acc * sp + g_sp * (inv - s1 * 2.0f * om[b] * inv * inv);
                d_rad += acc * om[b];
                dot += d_om[b] * om[b];
            }
            drow[DC_HEAD + 12] = __float2half_rn(d_rad * sgr);
#pragma unroll
            for (int b = 0; b < kNB; b++) drow[DC_HEAD + 13 + b] = __float2half_rn((d_om[b] - dot) * rinv * sigmoidf_(zz[b]));
            for (int i = 17; i < 32; i++) drow[DC_HEAD + i] = __float2half_rn(0.f);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float dv = g_vd[c] + g_dir[c] + g_vn * 2.0f * vd[c];
                drow[DC_VIEW + c] = __float2half_rn(dv * vd[c] * (1.0f - vd[c]));
                const float dd = g_dir[c] + g_df[c];
                drow[DC_DIFF + c] = __float2half_rn(dd * df[c] * (1.0f - df[c]));
            }
            for (int i = 3; i < 16; i++) { drow[DC_VIEW + i] = __float2half_rn(0.f); drow[DC_DIFF + i] = __float2half_rn(0.f); }
            if (d_palette) {   // gradient of the (clamped) palette colours: sum over the samples of this tile
#pragma unroll
                for (int k = 0; k < kNB * 3; k++) {
                    const float tot = warp_sum(dpal[k]);
                    if (lane == 0 && tot != 0.f) atomicAdd(d_palette + k, tot);
                }
            }
        }
        __syncwarp();

        // ---- fragment math, per 16-sample half ----
#pragma unroll 1
        for (int t = 0; t < 2; t++) {
            const uint32_t* xb = xtile + t * UX * 128;
            uint32_t* yb = ytile + t * UY * 128;
            const uint32_t s0 = tile * 32 + 16 * t;
            float c8[8][4];
            uint32_t a4[4][4];
            // heads -> basis net -> palette grid features
            {
                uint32_t ah[2][4];
                ldmatrix_a_s<kDStride>(ah[0], &bs.d[0][0], 16 * t, DC_HEAD, lane);
                ldmatrix_a_s<kDStride>(ah[1], &bs.d[0][0], 16 * t, DC_HEAD + 16, lane);
                st_unit(yb, YH, ah[0], lane);
                st_unit(yb, YH + 1, ah[1], lane);
                float c2[2][4];
                mma_layer<2, 2>(wt + tl_off(T_H), ah, c2, lane);
                uint32_t a1[1][4];
                chain<2, ACT_NONE>(c2, a1);
                st_unit(yb, YB1, a1[0], lane);
                mma_layer<1, 8>(wt + tl_off(T_B1), a1, c8, lane);
                deriv_pack<DRV_ELU>(c8, xb, XB1, yb, YB0, a4, lane);
                float c4[4][4];
                mma_layer<4, 4>(wt + tl_off(T_B0), a4, c4, lane);
                store_denc(d_enc, s0, M, c4, lane);
            }
            // diffuse net (its input, the geo features, is detached: no gradient leaves the first layer)
            {
                uint32_t ad[1][4];
                ldmatrix_a_s<kDStride>(ad[0], &bs.d[0][0], 16 * t, DC_DIFF, lane);
                st_unit(yb, YD2, ad[0], lane);
                mma_layer<1, 8>(wt + tl_off(T_D2), ad, c8, lane);
                deriv_pack<DRV_RELU>(c8, xb, XD2, yb, YD1, a4, lane);
                mma_layer<4, 8>(wt + tl_off(T_D1), a4, c8, lane);
                deriv_pack<DRV_RELU>(c8, xb, XD1, yb, YD0, a4, lane);
            }
            // view-dependent colour net (SH and geo inputs carry no gradient)
            {
                uint32_t av[1][4];
                ldmatrix_a_s<kDStride>(av[0], &bs.d[0][0], 16 * t, DC_VIEW, lane);
                st_unit(yb, YV2, av[0], lane);
                mma_layer<1, 8>(wt + tl_off(T_V2), av, c8, lane);
                deriv_pack<DRV_RELU>(c8, xb, XV2, yb, YV1, a4, lane);
                mma_layer<4, 8>(wt + tl_off(T_V1), a4, c8, lane);
                deriv_pack<DRV_RELU>(c8, xb, XV1, yb, YV0, a4, lane);
            }
            if (CLIP) {
                uint32_t ac[1][4];
                ldmatrix_a_s<kDStride>(ac[0], &bs.d[0][0], 16 * t, DC_CLIP, lane);
                st_unit(yb, YC1, ac[0], lane);
                mma_layer<1, 8>(wt + tl_off(T_C1), ac, c8, lane);
                deriv_pack<DRV_RELU>(c8, xb, XC1, yb, YC0, a4, lane);
                float c4[4][4];
                mma_layer<4, 4>(wt + tl_off(T_C0), a4, c4, lane);
                store_denc(d_enc_clip, s0, M, c4, lane);
            }
        }
        __syncwarp();
    }
}

// =====================================================================================================================
// weight gradients: dW[n_out][k_in] = sum_s dY[s][n_out] X[s][k_in]
// =====================================================================================================================
// One warp owns ALL 16-row blocks of one layer for a chunk of half-tiles, so every saved fragment is read exactly once
// (203 MB at 122 k samples instead of 371 MB when a warp owned a single 16-row block), and the loads of the next
// half-tile are issued before the MMAs of the current one.
__host__ inline void add_job(WJobs& J, int yslot, int ny, int xslot, int ux, int dwl) {
    add_job_at(J, yslot, ny, xslot, ux, dw_k(dwl), dw_off(dwl));
}

constexpr int kWgradWarps = 4;
#ifndef PNERF_WGRAD_MIN_CHUNK
#define PNERF_WGRAD_MIN_CHUNK 16
#endif

// One warp: split-K partial sum of one layer's dW over its chunks of half-tiles (registers), then the CTA's four partial
// sums are added in shared memory (the warps take turns: no shared-memory atomics, fixed order) and the CTA issues ONE
// fp32 reduction per weight — a quarter of the global reductions of a per-warp flush, which is what lets the chunks be
// half as long (twice the warps in flight for the same latency-bound loop).
template <int NY, int UX>
__device__ __forceinline__ void wgrad_job(const uint32_t* __restrict__ xbuf, const uint32_t* __restrict__ ybuf, uint32_t n_half,
                                          uint32_t per_chunk, uint32_t chunk0, uint32_t chunk_stride, uint32_t UXT, uint32_t UYT,
                                          const WJob& job, float* __restrict__ dwbuf, float* __restrict__ sdw, int lane, int wid) {
    float c[NY][2 * UX][4];
#pragma unroll
    for (int m = 0; m < NY; m++)
#pragma unroll
        for (int i = 0; i < 2 * UX; i++) c[m][i][0] = c[m][i][1] = c[m][i][2] = c[m][i][3] = 0.f;
    uint32_t p[NY][4], q[UX][4];
    auto load = [&](uint32_t h) {
        const uint32_t* yb = ybuf + (size_t)h * UYT * 128;
        const uint32_t* xb = xbuf + (size_t)h * UXT * 128;
#pragma unroll
        for (int m = 0; m < NY; m++) ld_unit(yb, job.yslot + m, p[m], lane);
#pragma unroll
        for (int j = 0; j < UX; j++) ld_unit(xb, job.xslot + j, q[j], lane);
    };
#pragma unroll 1
    for (uint32_t chunk = chunk0; chunk * per_chunk < n_half; chunk += chunk_stride) {
        const uint32_t h0 = chunk * per_chunk, h1 = min(n_half, h0 + per_chunk);
        load(h0);
#pragma unroll 1
        for (uint32_t h = h0; h < h1; h++) {
            // transposes of the current half-tile's fragments (registers), then prefetch the next half-tile
            uint32_t a[NY][4], b[UX][4];
#pragma unroll
            for (int m = 0; m < NY; m++) {   // A = (dY block)^T: 8x8 transposes + swap of the off-diagonal blocks
                a[m][0] = movmatrix_t(p[m][0]); a[m][1] = movmatrix_t(p[m][2]); a[m][2] = movmatrix_t(p[m][1]); a[m][3] = movmatrix_t(p[m][3]);
            }
#pragma unroll
            for (int j = 0; j < UX; j++) {   // B (k = sample, n = k_in), .col fragment order = transposes of the stored blocks
#pragma unroll
                for (int i = 0; i < 4; i++) b[j][i] = movmatrix_t(q[j][i]);
            }
            if (h + 1 < h1) load(h + 1);
#pragma unroll
            for (int m = 0; m < NY; m++)
#pragma unroll
                for (int j = 0; j < UX; j++) {
                    mma16816(c[m][2 * j], a[m], b[j][0], b[j][1]);
                    mma16816(c[m][2 * j + 1], a[m], b[j][2], b[j][3]);
                }
        }
    }
    // CTA-level sum: rows m * 16 + g (+8), compact row stride 16 * UX
    constexpr int KP = 16 * UX;
    const int g = lane >> 2, q2 = (lane & 3) * 2;
#pragma unroll 1
    for (int w = 0; w < kWgradWarps; w++) {
        if (wid == w) {
#pragma unroll
            for (int m = 0; m < NY; m++) {
                float* r0 = sdw + (m * 16 + g) * KP + q2;
                float* r1 = r0 + 8 * KP;
#pragma unroll
                for (int nt = 0; nt < 2 * UX; nt++) {
                    float2 v0 = make_float2(c[m][nt][0], c[m][nt][1]), v1 = make_float2(c[m][nt][2], c[m][nt][3]);
                    if (w != 0) {
                        const float2 o0 = *reinterpret_cast<const float2*>(r0 + nt * 8), o1 = *reinterpret_cast<const float2*>(r1 + nt * 8);
                        v0.x += o0.x; v0.y += o0.y; v1.x += o1.x; v1.y += o1.y;
                    }
                    *reinterpret_cast<float2*>(r0 + nt * 8) = v0;
                    *reinterpret_cast<float2*>(r1 + nt * 8) = v1;
                }
            }
        }
        __syncthreads();
    }
    float* dw = dwbuf + job.dwoff;
    for (int i = threadIdx.x; i < NY * 16 * KP; i += kWgradWarps * 32) {
        const int r = i / KP, col = i - r * KP;
        const float v = sdw[i];
        if (v != 0.f) atomicAdd(dw + r * job.kpad + col, v);
    }
}

__global__ void __launch_bounds__(kWgradWarps * 32)
k_field_wgrad(const uint32_t* __restrict__ xbuf, const uint32_t* __restrict__ ybuf, uint32_t M, const int32_t* __restrict__ m_dev,
              uint32_t UX, uint32_t UY, const __grid_constant__ WJobs jobs, float* __restrict__ dwbuf) {
    __shared__ __align__(16) float sdw[64 * 64];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const WJob job = jobs.j[blockIdx.y];
    if (m_dev) M = min(M, (uint32_t)__ldg(m_dev));
    const uint32_t n_half = ceil_div(M, 32u) * 2;
    // every warp of this job gets one chunk of half-tiles when there is enough work, but never fewer than
    // PNERF_WGRAD_MIN_CHUNK half-tiles so that the closing reductions (one per weight per CTA) stay negligible; long inputs loop.
    const uint32_t n_warps = gridDim.x * kWgradWarps;
    const uint32_t per_chunk = min(128u, max((uint32_t)PNERF_WGRAD_MIN_CHUNK, ceil_div(n_half, n_warps)));
    const uint32_t chunk0 = blockIdx.x * kWgradWarps;
    if (chunk0 * per_chunk >= n_half) return;                 // no warp of this CTA has work (CTA-uniform)
    const int shape = job.ny * 8 + job.ux;                    // uniform per blockIdx.y
    switch (shape) {
        case 4 * 8 + 1: wgrad_job<4, 1>(xbuf, ybuf, n_half, per_chunk, chunk0 + wid, n_warps, UX, UY, job, dwbuf, sdw, lane, wid); break;
        case 4 * 8 + 2: wgrad_job<4, 2>(xbuf, ybuf, n_half, per_chunk, chunk0 + wid, n_warps, UX, UY, job, dwbuf, sdw, lane, wid); break;
        case 4 * 8 + 3: wgrad_job<4, 3>(xbuf, ybuf, n_half, per_chunk, chunk0 + wid, n_warps, UX, UY, job, dwbuf, sdw, lane, wid); break;
        case 4 * 8 + 4: wgrad_job<4, 4>(xbuf, ybuf, n_half, per_chunk, chunk0 + wid, n_warps, UX, UY, job, dwbuf, sdw, lane, wid); break;
        case 1 * 8 + 4: wgrad_job<1, 4>(xbuf, ybuf, n_half, per_chunk, chunk0 + wid, n_warps, UX, UY, job, dwbuf, sdw, lane, wid); break;
        case 2 * 8 + 1: wgrad_job<2, 1>(xbuf, ybuf, n_half, per_chunk, chunk0 + wid, n_warps, UX, UY, job, dwbuf, sdw, lane, wid); break;
        default: break;
    }
}

int launch_field_wgrad(const uint32_t* xbuf, const uint32_t* ybuf, uint32_t M, const int32_t* m_dev, uint32_t UX, uint32_t UY,
                       const WJobs& J, float* dwbuf, cudaStream_t stream, const char* what) {
    const uint32_t n_half_cap = ceil_div(M, 32u) * 2;
    const uint32_t grid_x = min(ceil_div(ceil_div(n_half_cap, (uint32_t)PNERF_WGRAD_MIN_CHUNK), (uint32_t)kWgradWarps), 4u * (uint32_t)kNumSMs);
    const dim3 grid(max(grid_x, 1u), J.n, 1);
    k_field_wgrad<<<grid, kWgradWarps * 32, 0, stream>>>(xbuf, ybuf, M, m_dev, UX, UY, J, dwbuf);
    return check_launch(what);
}

}  // namespace pnerf

using namespace pnerf;

extern "C" {

uint64_t pnerf_palette_train_xbuf_bytes(uint32_t M, uint32_t pred_clip) {
    return (uint64_t)ceil_div(M, 32u) * 2 * (pred_clip ? kUXClip : kUXNoClip) * 512;
}
uint64_t pnerf_palette_train_ybuf_bytes(uint32_t M, uint32_t pred_clip) {
    return (uint64_t)ceil_div(M, 32u) * 2 * (pred_clip ? kUYClip : kUYNoClip) * 512;
}
uint32_t pnerf_palette_train_dw_floats(uint32_t pred_clip) { return pred_clip ? kDwFloatsClip : kDwFloatsNoClip; }
uint32_t pnerf_palette_train_wfwd_units(uint32_t pred_clip) { return pred_clip ? kWUnitsClip : kWUnitsNoClip; }
uint32_t pnerf_palette_train_wbwd_units(uint32_t pred_clip) { return pred_clip ? kTUnitsClip : kTUnitsNoClip; }

static int train_args_ok(const pnerf_palette_train* p) {
    if (!p || !p->offsets || !p->wfwd || !p->wbwd || !p->palette) return PNERF_ERR_INVALID_ARG;
    if (!p->table_sigma_palette && (!p->table_sigma || !p->table_palette)) return PNERF_ERR_INVALID_ARG;
    if (p->pred_clip && !p->table_clip) return PNERF_ERR_INVALID_ARG;
    if (p->L != 16 || p->clip_dim > (uint32_t)kClipMax) return PNERF_ERR_UNSUPPORTED;
    return PNERF_OK;
}

int pnerf_palette_train_forward(const float* xyzs, const float* dirs, uint32_t M, const pnerf_palette_train* p, void* xbuf,
                                float* sigma, float* rgb, float* flex, void* stream) {
    if (M == 0) return PNERF_OK;
    PNERF_REQUIRE(xyzs && dirs && xbuf && sigma && rgb && flex);
    if (int st = train_args_ok(p)) return st;
    cudaStream_t s = (cudaStream_t)stream;
    const bool clip = p->pred_clip != 0;
    const size_t smem = ((sizeof(TrainSmem) + 15) & ~(size_t)15) + (size_t)(clip ? kWUnitsClip : kWUnitsNoClip) * sizeof(uint2) +
                        sizeof(WarpScratch) * kFusedWarps;
    const uint32_t grid = min(ceil_div(ceil_div(M, 32u), (uint32_t)kFusedWarps), (uint32_t)kNumSMs);
    static bool attr_done[2] = {false, false};   // set once per process (keeps cudaFuncSetAttribute out of graph capture)
    if (!attr_done[clip]) {
        cudaError_t e = clip ? cudaFuncSetAttribute(k_field_train_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                             : cudaFuncSetAttribute(k_field_train_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_cuda_error(e, "palette_train_forward attr"); return PNERF_ERR_CUDA; }
        attr_done[clip] = true;
    }
    if (clip) {
        k_field_train_fwd<true><<<grid, kFusedWarps * 32, smem, s>>>(xyzs, dirs, M, *p, (uint32_t*)xbuf, sigma, rgb, flex);
    } else {
        k_field_train_fwd<false><<<grid, kFusedWarps * 32, smem, s>>>(xyzs, dirs, M, *p, (uint32_t*)xbuf, sigma, rgb, flex);
    }
    return check_launch("palette_train_forward");
}

int pnerf_palette_train_backward(uint32_t M, const pnerf_palette_train* p, const void* xbuf, void* ybuf, const float* grad_rgb,
                                 const float* grad_flex, const float* flex, float* d_enc, float* d_enc_clip, float* d_palette,
                                 void* stream) {
    if (M == 0) return PNERF_OK;
    PNERF_REQUIRE(xbuf && ybuf && grad_rgb && grad_flex && flex && d_enc);
    if (int st = train_args_ok(p)) return st;
    const bool clip = p->pred_clip != 0;
    if (clip) PNERF_REQUIRE(d_enc_clip != nullptr);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = 64 + (size_t)(3 * 32 + (clip ? kTUnitsClip : kTUnitsNoClip)) * sizeof(uint2) + sizeof(BwdScratch) * kTrainWarps;
    const uint32_t grid = min(ceil_div(ceil_div(M, 32u), (uint32_t)kTrainWarps), 2u * (uint32_t)kNumSMs);
    static bool attr_done[2] = {false, false};
    if (!attr_done[clip]) {
        cudaError_t e = clip ? cudaFuncSetAttribute(k_field_train_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                             : cudaFuncSetAttribute(k_field_train_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_cuda_error(e, "palette_train_backward attr"); return PNERF_ERR_CUDA; }
        attr_done[clip] = true;
    }
    if (clip) {
        k_field_train_bwd<true><<<grid, kTrainWarps * 32, smem, s>>>(M, *p, (const uint32_t*)xbuf, (uint32_t*)ybuf, grad_rgb,
                                                                   grad_flex, flex, d_enc, d_enc_clip, d_palette);
    } else {
        k_field_train_bwd<false><<<grid, kTrainWarps * 32, smem, s>>>(M, *p, (const uint32_t*)xbuf, (uint32_t*)ybuf, grad_rgb,
                                                                    grad_flex, flex, d_enc, d_enc_clip, d_palette);
    }
    return check_launch("palette_train_backward");
}

int pnerf_palette_train_wgrad(uint32_t M, uint32_t flags, const void* xbuf, const void* ybuf, float* dwbuf,
                              const int32_t* m_dev, void* stream) {
    if (M == 0) return PNERF_OK;
    PNERF_REQUIRE(xbuf && ybuf && dwbuf);
    const uint32_t pred_clip = flags & 1u;          // PNERF_TRAIN_WGRAD_CLIP
    const bool basis = (flags & 2u) != 0;           // PNERF_TRAIN_WGRAD_BASIS_NET
    WJobs J;
    J.n = 0;
    add_job(J, YD0, 4, XD0, 1, DW_D0);
    add_job(J, YD1, 4, XD1, 4, DW_D1);
    add_job(J, YD2, 1, XD2, 4, DW_D2);
    add_job(J, YV0, 4, XV0, 2, DW_V0);
    add_job(J, YV1, 4, XV1, 4, DW_V1);
    add_job(J, YV2, 1, XV2, 4, DW_V2);
    if (basis) {   // the reference never steps basis_net (absent from get_params, palette/network.py:283-308): its weight
                   // gradients are computed on request only; their slots of dwbuf stay as the caller initialised them
        add_job(J, YB0, 4, XB0, 3, DW_B0);
        add_job(J, YB1, 1, XB1, 4, DW_B1);
    }
    add_job(J, YH, 2, XH, 1, DW_H);
    if (pred_clip) {
        add_job(J, YC0, 4, XC0, 2, DW_C0);
        add_job(J, YC1, 1, XC1, 4, DW_C1);
    }
    return launch_field_wgrad((const uint32_t*)xbuf, (const uint32_t*)ybuf, M, m_dev, pred_clip ? kUXClip : kUXNoClip,
                              pred_clip ? kUYClip : kUYNoClip, J, dwbuf, (cudaStream_t)stream, "palette_train_wgrad");
}

}  // extern "C"
