// fused_field.cuh — the fused per-sample field of PaletteNeRF (hash grids -> sigma / diffuse / view-dependent /
// palette-basis / semantic MLPs -> palette heads) as device functions shared by the batch kernel (fused.cu), the
// lane-per-ray renderer (fused.cu) and the warp-per-ray renderer (render_rays.cu). See fused.cu for the design notes.
#pragma once
#include "fused_common.cuh"

#ifndef PNERF_GATHER_LV
#define PNERF_GATHER_LV 2     // levels per gather iteration (x 8 corners x 8 B loads in flight per lane); A/B on B200 at
                              // 800x800: LV 1 7.47 ms, LV 2 7.30 ms, LV 4 7.99 ms per view (profiles/README.md)
#endif

#ifdef PNERF_TILE_UNROLL
#define PNERF_TILE_LOOP _Pragma("unroll")
#else
#define PNERF_TILE_LOOP _Pragma("unroll 1")
#endif

namespace pnerf {

// ------------------------------------------------------------------------------------------------
// the fused field: 32 samples per warp (lane = sample), L must be 16 (feature rows are 32 wide)
// ------------------------------------------------------------------------------------------------
//   COOP: gather with lane pairs (gather_coop: half the L1/TEX wavefronts, FHFMA interpolation) instead of one lane per
//         sample (gather_fast); requires the interleaved density+palette table and mask-wrapped levels (fast_wrap)
template <bool CLIP, bool COOP = false>
__device__ __forceinline__ void eval_field(const pnerf_palette_field& f, const FusedSmem& sm, const uint2* __restrict__ wts,
                                           WarpScratch& ws, float x, float y, float z, float dx, float dy, float dz,
                                           bool active, int lane, FieldOut& o) {
    const float u = (x + f.bound) / (2 * f.bound), v = (y + f.bound) / (2 * f.bound), w = (z + f.bound) / (2 * f.bound);
    const bool in_range = active && !((u < 0 || u > 1) || (v < 0 || v > 1) || (w < 0 || w > 1));

    // Per-tile fragments that outlive a phase ([logit | geo15] k-step: 4 words, [diffuse3 | pad] k-step: 2 words) are
    // parked in this lane's own output row, in the columns the clip head only writes in phase 4 (after they are dead).
    uint32_t* carry = reinterpret_cast<uint32_t*>(&ws.out[lane][O_CLIP]);   // [t][6]

    // ---------------- phase 1: density grid -> sigma net -> geo; geo -> diffuse net ----------------
    // The palette grid shares the density grid's geometry: both tables are read here with ONE set of corner indices;
    // the palette features wait (as fp16 pairs) in this lane's output row, columns O_OFFRAD.. that phase 3 writes last.
    uint32_t* park = reinterpret_cast<uint32_t*>(&ws.out[lane][O_OFFRAD]);   // 16 words
    const bool paired = COOP || (sm.fast_wrap && f.table_sigma_palette != nullptr);      // warp-uniform
    if (COOP) {
        auto st = [&ws](int e, int s, int l0, const uint32_t (&words)[PNERF_COOP_LV]) {
            uint32_t* dst = e == 0 ? reinterpret_cast<uint32_t*>(ws.feat[s]) : reinterpret_cast<uint32_t*>(&ws.out[s][O_OFFRAD]);
#pragma unroll
            for (int j = 0; j < PNERF_COOP_LV; j++) dst[l0 + j] = words[j];
        };
        gather_coop<2, PNERF_COOP_LV, PNERF_GATHER_HACC != 0>(f.table_sigma_palette, sm.lp, u, v, w, in_range, lane, st);
    } else if (paired) {
        uint32_t* const rows[2] = {reinterpret_cast<uint32_t*>(ws.feat[lane]), park};
        gather_fast<2, PNERF_GATHER_LV, PNERF_GATHER_HACC != 0>(f.table_sigma_palette, sm.lp, u, v, w, in_range, rows);
    } else {
        gather_features((const __half*)f.table_sigma, sm.lp, f.L, u, v, w, in_range, ws.feat[lane]);
    }
    __syncwarp();
PNERF_TILE_LOOP
    for (int t = 0; t < 2; t++) {
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
#pragma unroll
        for (int i = 0; i < 4; i++) carry[t * 6 + i] = geo[0][i];
        // diffuse net 15 -> 64 -> 64 -> 3
        mma_layer<1, 8>(wts + layer_off(LD0), geo, c8, lane);
        chain<8, ACT_RELU>(c8, a4);
        mma_layer<4, 8>(wts + layer_off(LD1), a4, c8, lane);
        chain<8, ACT_RELU>(c8, a4);
        float c1[1][4];
        mma_layer<4, 1>(wts + layer_off(LD2), a4, c1, lane);
#pragma unroll
        for (int i = 0; i < 4; i++) c1[0][i] = sigmoidf_(c1[0][i]);
        store_out<1>(ws.out, 16 * t, O_DIFF, 3, c1, lane);
        carry[t * 6 + 4] = pack_h2(c1[0][0], c1[0][1]);
        carry[t * 6 + 5] = pack_h2(c1[0][2], c1[0][3]);
    }
    __syncwarp();

    // ---------------- phase 2: SH(4) of the view direction ++ geo -> view-dependent colour net ----------------
    {
        float sh[16];
        sh_eval<4, false>(dx, dy, dz, sh, nullptr, nullptr, nullptr);
#pragma unroll
        for (int i = 0; i < 8; i++) reinterpret_cast<__half2*>(ws.feat[lane])[i] = __floats2half2_rn(sh[2 * i], sh[2 * i + 1]);
    }
    __syncwarp();
PNERF_TILE_LOOP
    for (int t = 0; t < 2; t++) {
        uint32_t a2[2][4];
        ldmatrix_a(a2[0], &ws.feat[0][0], 16 * t, 0, lane);
#pragma unroll
        for (int i = 0; i < 4; i++) a2[1][i] = carry[t * 6 + i];
        float c8[8][4];
        mma_layer<2, 8>(wts + layer_off(LV0), a2, c8, lane);
        uint32_t a4[4][4];
        chain<8, ACT_RELU>(c8, a4);
        mma_layer<4, 8>(wts + layer_off(LV1), a4, c8, lane);
        chain<8, ACT_RELU>(c8, a4);
        float c1[1][4];
        mma_layer<4, 1>(wts + layer_off(LV2), a4, c1, lane);
#pragma unroll
        for (int i = 0; i < 4; i++) c1[0][i] = sigmoidf_(c1[0][i]);
        store_out<1>(ws.out, 16 * t, O_VIEW, 3, c1, lane);
    }
    __syncwarp();

    // ---------------- phase 3: palette grid ++ diffuse -> basis net -> offsets/radiance + omega heads ----------------
    if (paired) {
#pragma unroll
        for (int i = 0; i < 16; i++) reinterpret_cast<uint32_t*>(ws.feat[lane])[i] = park[i];
    } else {
        gather_features((const __half*)f.table_palette, sm.lp, f.L, u, v, w, in_range, ws.feat[lane]);
    }
    __syncwarp();
PNERF_TILE_LOOP
    for (int t = 0; t < 2; t++) {
        uint32_t a3[3][4];
        ldmatrix_a(a3[0], &ws.feat[0][0], 16 * t, 0, lane);
        ldmatrix_a(a3[1], &ws.feat[0][0], 16 * t, 16, lane);
        a3[2][0] = carry[t * 6 + 4];
        a3[2][1] = carry[t * 6 + 5];
        a3[2][2] = 0u;
        a3[2][3] = 0u;
        float c8[8][4];
        mma_layer<3, 8>(wts + layer_off(LB0), a3, c8, lane);
        uint32_t a4[4][4];
        chain<8, ACT_ELU>(c8, a4);
        float c2[2][4];
        mma_layer<4, 2>(wts + layer_off(LB1), a4, c2, lane);
        uint32_t a1[1][4];
        chain<2, ACT_NONE>(c2, a1);
        float c3[3][4];
        mma_layer<1, 3>(wts + layer_off(LH), a1, c3, lane);
        store_out<3>(ws.out, 16 * t, O_OFFRAD, 13 + kNB, c3, lane);   // cols 7..19 offsets/radiance, 20..23 omega logits
    }
    __syncwarp();

    // ---------------- phase 4 (optional): semantic grid -> clip net ----------------
    if (CLIP) {
        if (COOP) {
            auto st = [&ws](int, int s, int l0, const uint32_t (&words)[PNERF_COOP_LV]) {
#pragma unroll
                for (int j = 0; j < PNERF_COOP_LV; j++) reinterpret_cast<uint32_t*>(ws.feat[s])[l0 + j] = words[j];
            };
            gather_coop<1, PNERF_COOP_LV, PNERF_GATHER_HACC != 0>(f.table_clip, sm.lp, u, v, w, in_range, lane, st);
        } else if (sm.fast_wrap) {
            uint32_t* const rows[1] = {reinterpret_cast<uint32_t*>(ws.feat[lane])};
            gather_fast<1, PNERF_GATHER_LV, PNERF_GATHER_HACC != 0>(f.table_clip, sm.lp, u, v, w, in_range, rows);
        } else {
            gather_features((const __half*)f.table_clip, sm.lp, f.L, u, v, w, in_range, ws.feat[lane]);
        }
        __syncwarp();
PNERF_TILE_LOOP
        for (int t = 0; t < 2; t++) {
            uint32_t a2[2][4];
            ldmatrix_a(a2[0], &ws.feat[0][0], 16 * t, 0, lane);
            ldmatrix_a(a2[1], &ws.feat[0][0], 16 * t, 16, lane);
            float c8[8][4];
            mma_layer<2, 8>(wts + layer_off(LC0), a2, c8, lane);
            uint32_t a4[4][4];
            chain<8, ACT_RELU>(c8, a4);
            float c2[2][4];
            mma_layer<4, 2>(wts + layer_off(LC1), a4, c2, lane);
            store_out<2>(ws.out, 16 * t, O_CLIP, (int)f.clip_dim, c2, lane);
        }
        __syncwarp();
    }

    // ---------------- collect this lane's sample ----------------
    const float* row = ws.out[lane];
    o.sigma = __expf(row[O_SIGMA]);
#pragma unroll
    for (int i = 0; i < 3; i++) { o.diffuse[i] = row[O_DIFF + i]; o.view_dep[i] = row[O_VIEW + i]; }
#pragma unroll
    for (int i = 0; i < 13; i++) o.off_rad[i] = row[O_OFFRAD + i] + sm.head_bias[i];
    float osum = 0.f;
#pragma unroll
    for (int b = 0; b < kNB; b++) { o.omega[b] = softplusf_(row[O_OMEGA + b]) + 0.05f; osum += o.omega[b]; }
    const float rinv = 1.0f / osum;
#pragma unroll
    for (int b = 0; b < kNB; b++) o.omega[b] *= rinv;
#pragma unroll
    for (int i = 0; i < kClipMax; i++) o.clip[i] = (CLIP && i < (int)f.clip_dim) ? row[O_CLIP + i] : 0.f;
    __syncwarp();
}

// palette blend of one sample (ref: palette/renderer.py:470-494): basis_rgb[b][c] = omega_b * softplus(radiance) *
// (palette_bc + offsets_weight * offsets_bc); rgb = sum_b basis_rgb + view_dep_weight * view_dep
__device__ __forceinline__ void blend(const pnerf_palette_field& f, const FusedSmem& sm, const FieldOut& o, float (&rgb)[3],
                                      float (&basis_rgb)[kNB * 3], float (&unscaled)[kNB * 3]) {
    const float sp = softplusf_(o.off_rad[12]);
    rgb[0] = rgb[1] = rgb[2] = 0.f;
#pragma unroll
    for (int b = 0; b < kNB; b++) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float off = o.off_rad[b * 3 + c];
            unscaled[b * 3 + c] = sm.palette[b * 3 + c] + off;
            basis_rgb[b * 3 + c] = o.omega[b] * (sp * (sm.palette[b * 3 + c] + f.offsets_weight * off));
            rgb[c] += basis_rgb[b * 3 + c];
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) rgb[c] += f.view_dep_weight * o.view_dep[c];
}

// ------------------------------------------------------------------------------------------------
// CTA prologue shared by both kernels: level table, head bias, palette, weights -> shared memory
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fused_prologue(const pnerf_palette_field& f, FusedSmem* sm, uint2* wts) {
    if (threadIdx.x < f.L) make_level(sm->lp[threadIdx.x], threadIdx.x, f.offsets, f.S, f.H, 3, 0, false);
    const int slow = __syncthreads_or(threadIdx.x < f.L && sm->lp[threadIdx.x].mask == 0u);
    if (threadIdx.x == 0) sm->fast_wrap = slow ? 0u : 1u;
    if (threadIdx.x < 16) sm->head_bias[threadIdx.x] = f.head_bias[threadIdx.x];
    if (threadIdx.x < kNB * 3) sm->palette[threadIdx.x] = f.palette[threadIdx.x];
    const int units = f.pred_clip ? kWUnitsClip : kWUnitsNoClip;   // uint2 units; both counts are even
    const uint4* src = reinterpret_cast<const uint4*>(f.wpack);
    uint4* dst = reinterpret_cast<uint4*>(wts);
    for (int i = threadIdx.x; i < units / 2; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
}

__host__ __device__ constexpr size_t fused_smem_bytes(bool clip, bool aux, bool clip_acc = false) {
    return sizeof(FusedSmem) + (size_t)(clip ? kWUnitsClip : kWUnitsNoClip) * sizeof(uint2) +
           sizeof(WarpScratch) * kFusedWarps + (aux ? sizeof(WarpAux) * kFusedWarps : 0) +
           (clip_acc ? sizeof(WarpClip) * kFusedWarps : 0) + 16;
}

}  // namespace pnerf
