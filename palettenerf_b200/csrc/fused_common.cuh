// fused_common.cuh — pieces shared by the fused inference kernels (fused.cu) and the fused training kernels
// (fused_train.cu): layer table of the packed weight blob, mma.sync / ldmatrix helpers, activation helpers and the
// per-lane hash-grid gather.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "grid_common.cuh"
#include "march_common.cuh"
#include "sh_common.cuh"

namespace pnerf {

#ifndef PNERF_GATHER_HACC
#define PNERF_GATHER_HACC 1   // packed-half interpolation in the INFERENCE renderers' gathers (gather_coop / gather_fast <.., HACC>)
#endif
#ifndef PNERF_COOP_LV
#define PNERF_COOP_LV 4       // levels per iteration of the lane-pair gather (x 4 corners x 8 B loads in flight per lane)
#endif

constexpr int kNB = 4;            // palette bases supported by the fused path (reference default, main_palette.py:99)
constexpr int kClipMax = 16;      // semantic feature width supported by the fused path (main_palette.py:76)
#ifndef PNERF_FUSED_WARPS
#define PNERF_FUSED_WARPS 12
#endif
constexpr int kFusedWarps = PNERF_FUSED_WARPS;   // one CTA per SM: 12 warps <= 168 registers/thread, 10 warps <= 204
constexpr int kAuxCh = 3 + 3 + kNB + 2 * kNB * 3;   // direct_rgb, view_dep_rgb, basis_acc, basis_rgb, unscaled_basis_rgb
constexpr int kFeatStride = 40;   // halfs per feature row: 32 + 8 pad -> ldmatrix rows hit distinct bank groups
#ifndef PNERF_OUT_STRIDE
#define PNERF_OUT_STRIDE 41
#endif
constexpr int kOutStride = PNERF_OUT_STRIDE;    // floats per output row (40 used); odd stride -> conflict-free row-per-lane reads

// packed weight blob: per layer [NT][KS][32 lanes] x uint2 (= the m16n8k16 B fragment of that (n-tile, k-step))
enum Layer { LS0, LS1, LD0, LD1, LD2, LV0, LV1, LV2, LB0, LB1, LH, LC0, LC1, kNumLayers };
__host__ __device__ constexpr int layer_ks(int l) {
    return l == LS0 ? 2 : l == LS1 ? 4 : l == LD0 ? 1 : l == LD1 ? 4 : l == LD2 ? 4 : l == LV0 ? 2 : l == LV1 ? 4
         : l == LV2 ? 4 : l == LB0 ? 3 : l == LB1 ? 4 : l == LH ? 1 : l == LC0 ? 2 : 4;
}
__host__ __device__ constexpr int layer_nt(int l) {
    return l == LS0 ? 8 : l == LS1 ? 2 : l == LD0 ? 8 : l == LD1 ? 8 : l == LD2 ? 1 : l == LV0 ? 8 : l == LV1 ? 8
         : l == LV2 ? 1 : l == LB0 ? 8 : l == LB1 ? 2 : l == LH ? 3 : l == LC0 ? 8 : 2;
}
__host__ __device__ constexpr int layer_off(int l) {  // in uint2 units
    int o = 0;
    for (int i = 0; i < l; i++) o += layer_ks(i) * layer_nt(i) * 32;
    return o;
}
constexpr int kWUnitsNoClip = layer_off(LC0);
constexpr int kWUnitsClip = layer_off(kNumLayers);

// output staging columns
enum OutCol { O_SIGMA = 0, O_DIFF = 1, O_VIEW = 4, O_OFFRAD = 7, O_OMEGA = 20, O_CLIP = 24 };

struct WarpScratch {
    __half feat[32][kFeatStride];
    float out[32][kOutStride];
};
// per-warp auxiliary-map accumulators of the renderer, channel-major so that lane-per-ray accesses are conflict-free
struct WarpAux {
    float acc[kAuxCh][32];
};

// per-warp accumulators of the semantic-feature map (renderer with pred_clip): channel-major like WarpAux
struct WarpClip {
    float acc[kClipMax][32];
};

struct FusedSmem {
    LevelParams lp[kMaxLevels];
    float head_bias[16];
    float palette[kNB * 3];
    uint32_t fast_wrap;   // every level wraps with a mask (power-of-two hashed level or dense level that fits its table)
    // followed by: uint2 weights[units]; WarpScratch scratch[kFusedWarps]
};

}  // namespace pnerf

namespace pnerf {

// ------------------------------------------------------------------------------------------------
// tensor-core helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragment (m16 x k16, fp16) from a row-major shared-memory tile: rows [row0, row0+16), cols [col0, col0+16)
__device__ __forceinline__ void ldmatrix_a(uint32_t (&a)[4], const __half* tile, int row0, int col0, int lane) {
    const __half* p = tile + (row0 + (lane & 7) + ((lane >> 3) & 1) * 8) * kFeatStride + col0 + (lane >> 4) * 8;
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                 : "r"(addr));
}

template <int KS, int NT>
__device__ __forceinline__ void mma_layer(const uint2* __restrict__ w, const uint32_t (&a)[KS][4], float (&c)[NT][4],
                                          int lane) {
#ifdef PNERF_MMA_KS_OUTER
    // k-step outer: consecutive MMAs go to different accumulators (no back-to-back dependent HMMAs); every accumulator
    // still sums its k-steps in the order 0..KS-1, so the result is bit-identical to the n-tile-outer order
#pragma unroll
    for (int nt = 0; nt < NT; nt++) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            const uint2 b = w[(nt * KS + ks) * 32 + lane];
            mma16816(c[nt], a[ks], b.x, b.y);
        }
    }
#else
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
        c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            const uint2 b = w[(nt * KS + ks) * 32 + lane];
            mma16816(c[nt], a[ks], b.x, b.y);
        }
    }
#endif
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// exp(v) as ONE multiply + ONE MUFU.EX2 (ex2.approx.ftz: 2 ulp, denormal results flush to zero). __expf() spends ~6 more
// instructions per value on scaling denormal results, which no consumer here can tell from zero.
__device__ __forceinline__ float fast_exp(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v * 1.4426950408889634f));
    return r;
}
__device__ __forceinline__ float fast_rcp(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

enum Act { ACT_NONE, ACT_RELU, ACT_ELU };
template <int ACT>
__device__ __forceinline__ float activate(float v) {
    if (ACT == ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == ACT_ELU) return v > 0.f ? v : (fast_exp(v) - 1.0f);
    return v;
}

// pack two activated accumulators. ReLU is applied to the packed pair (one HMNMX2 instead of two FMNMX):
// round-to-nearest is monotonic and sign-preserving, so max(rn(x), 0) == rn(max(x, 0)) for every finite x.
template <int ACT>
__device__ __forceinline__ uint32_t pack_act(float lo, float hi) {
    if (ACT == ACT_RELU) {   // ONE instruction (F2FP.RELU): convert, clamp at zero and pack
        uint32_t h;
        asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(hi), "f"(lo));
        return h;
    }
    return pack_h2(activate<ACT>(lo), activate<ACT>(hi));
}

// accumulators of NT n-tiles -> A fragments of NT/2 k-steps (layout identity of mma.m16n8k16, no shuffles)
template <int NT, int ACT>
__device__ __forceinline__ void chain(const float (&c)[NT][4], uint32_t (&a)[NT / 2][4]) {
#pragma unroll
    for (int j = 0; j < NT / 2; j++) {
        a[j][0] = pack_act<ACT>(c[2 * j][0], c[2 * j][1]);
        a[j][1] = pack_act<ACT>(c[2 * j][2], c[2 * j][3]);
        a[j][2] = pack_act<ACT>(c[2 * j + 1][0], c[2 * j + 1][1]);
        a[j][3] = pack_act<ACT>(c[2 * j + 1][2], c[2 * j + 1][3]);
    }
}

// write an n-tile accumulator (rows r / r+8 of the m16 tile, cols 2q, 2q+1) into the fp32 output staging
template <int NT>
__device__ __forceinline__ void store_out(float (*out)[kOutStride], int row0, int col0, int ncols, const float (&c)[NT][4],
                                          int lane) {
    const int r = lane >> 2, q = lane & 3;
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
        const int col = nt * 8 + 2 * q;
        if (col < ncols) { out[row0 + r][col0 + col] = c[nt][0]; out[row0 + r + 8][col0 + col] = c[nt][2]; }
        if (col + 1 < ncols) { out[row0 + r][col0 + col + 1] = c[nt][1]; out[row0 + r + 8][col0 + col + 1] = c[nt][3]; }
    }
}

__device__ __forceinline__ float sigmoidf_(float v) { return fast_rcp(1.0f + fast_exp(-v)); }
// torch threshold 20; log(1 + e^v) instead of log1p(e^v): the absolute error is < 6e-8 (the consumers add 0.05 / multiply
// an O(1) colour), at a quarter of the instructions
__device__ __forceinline__ float softplusf_(float v) { return v > 20.f ? v : __logf(1.0f + fast_exp(v)); }

// ------------------------------------------------------------------------------------------------
// hash-grid gather of one sample (this lane) into its fp16 feature row
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ void gather_features(const __half* __restrict__ table, const LevelParams* __restrict__ lp,
                                                uint32_t L, float u, float v, float w, bool active,
                                                __half* __restrict__ row) {
    for (uint32_t l0 = 0; l0 < L; l0 += 2) {
        float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        if (active) {
            float2 val[2][8];
            float wt[2][8];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                if (l0 + j < L) {
                    const LevelParams& p = lp[l0 + j];
                    uint32_t idx[8];
                    corner_setup(p, u, v, w, idx, wt[j], false);
                    const __half2* g = reinterpret_cast<const __half2*>(table) + p.offset;
#pragma unroll
                    for (int c = 0; c < 8; c++) val[j][c] = __half22float2(__ldg(g + idx[c]));
                }
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                if (l0 + j < L) {
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        acc[j].x += wt[j][c] * val[j][c].x;
                        acc[j].y += wt[j][c] * val[j][c].y;
                    }
                }
            }
        }
        reinterpret_cast<__half2*>(row)[l0] = __floats2half2_rn(acc[0].x, acc[0].y);
        if (l0 + 1 < L) reinterpret_cast<__half2*>(row)[l0 + 1] = __floats2half2_rn(acc[1].x, acc[1].y);
    }
}

// ------------------------------------------------------------------------------------------------
// gather_fast: the hot-path gather. The density, palette and semantic grids of the model share ONE geometry (same
// level table), so the 8 corner indices and trilinear weights of a level are computed once; with the density and
// palette tables interleaved entry by entry (pnerf_palette_field::table_sigma_palette, EW = 2) one 8-byte load per
// corner fetches both grids' features: half the loads and half the L2 sectors of two separate gathers.
// Requires every level to wrap with a mask (FusedSmem::fast_wrap, true for every table GridEncoder can construct:
// hashed levels hold 2^log2T entries, dense levels fit); the generic gather_features above stays as the fallback.
// Address arithmetic is one mad.wide.u32 per load (32-bit entry index against a per-level 64-bit base).
// Out-of-range samples produce zero features like the reference kernel (gridencoder.cu:118-130).
//   EW: 32-bit words per table entry (1: one F=2 fp16 table, 2: two interleaved tables); rows[e] receives table e.
//   LV: levels per iteration (LV x 8 loads in flight per lane).
// ------------------------------------------------------------------------------------------------
// floor of a coordinate 0 <= p < 2^22 WITHOUT the conversion pipe: t = p + 2^23 rounded toward -inf is 2^23 + floor(p)
// exactly (FADD.RM, fma pipe, 4 cycles), its low 23 bits are the integer and t - 2^23 is the floor as a float (exact) — the
// same bits floorf() and the float -> int conversion deliver (FRND + F2I: quarter-rate XU pipe, tracked by the short
// scoreboard, in the middle of the address chain of every gather). `raw` keeps the exponent bits (0x4B000000 + floor).
__device__ __forceinline__ void floor_split(float p, float& fl, uint32_t& raw) {
    const float t = __fadd_rd(p, 8388608.0f);
    raw = __float_as_uint(t);
    fl = t - 8388608.0f;
}
constexpr uint32_t kFloorRawMask = 0x007fffffu;   // raw & mask = floor(p) as an integer

template <int EW>
__device__ __forceinline__ void ldg_entry(uint64_t base, uint32_t idx, uint32_t (&v)[EW]) {
    uint64_t addr;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(addr) : "r"(idx), "n"(EW * 4), "l"(base));
    if (EW == 1) asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v[0]) : "l"(addr));
    else asm volatile("ld.global.nc.v2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[EW - 1]) : "l"(addr));
}
// (Tried: the entry size as a run-time register operand so that the address is one IMAD.WIDE.U32 instead of the LEA +
// LEA.HI.X pair ptxas makes of a power-of-two immediate — ptxas then emits IMAD.WIDE + IADD3 + IMAD.X, three slots. Kept as is.)

template <int EW, int LV, bool HACC = false>
__device__ __forceinline__ void gather_fast(const void* __restrict__ table, const LevelParams* __restrict__ lp,
                                            float u, float v, float w, bool in_range, uint32_t* const (&rows)[EW]) {
    static_assert(EW == 1 || EW == 2, "one table or two interleaved tables");
    u = fminf(fmaxf(u, 0.f), 1.f); v = fminf(fmaxf(v, 0.f), 1.f); w = fminf(fmaxf(w, 0.f), 1.f);   // keeps the loads in bounds
#pragma unroll 1
    for (int l0 = 0; l0 < 16; l0 += LV) {
        uint32_t val[LV][8][EW];
        float wt[LV][8];
#pragma unroll
        for (int j = 0; j < LV; j++) {
            const LevelParams& p = lp[l0 + j];
            const float px = fmaf(u, p.scale, 0.5f), py = fmaf(v, p.scale, 0.5f), pz = fmaf(w, p.scale, 0.5f);
            float fx0, fy0, fz0;
            uint32_t gx, gy, gz;                       // raw words (see floor_split): stripped in the dense branch only
            floor_split(px, fx0, gx); floor_split(py, fy0, gy); floor_split(pz, fz0, gz);
            const float rx = px - fx0, ry = py - fy0, rz = pz - fz0;
            const float wx[2] = {1 - rx, rx}, wy[2] = {1 - ry, ry}, wz[2] = {1 - rz, rz};
            const float wxy[4] = {wx[0] * wy[0], wx[1] * wy[0], wx[0] * wy[1], wx[1] * wy[1]};   // (wx*wy)*wz: the reference's order
#pragma unroll
            for (int c = 0; c < 8; c++) wt[j][c] = wxy[c & 3] * wz[c >> 2];
            uint32_t idx[8];
            if (p.use_hash) {
                // hashed level, table of <= 2^24 entries (make_level): the exponent bits of the raw words reach index bits >= 24
                // only (x: as they are; y, z: 0x4B000000 * prime has no bit below 24), which the mask drops
                const uint32_t hx1 = gx + 1, hy0 = gy * 2654435761u, hz0 = gz * 805459861u;
                const uint32_t hy1 = hy0 + 2654435761u, hz1 = hz0 + 805459861u;
                const uint32_t yz[4] = {hy0 ^ hz0, hy1 ^ hz0, hy0 ^ hz1, hy1 ^ hz1};
#pragma unroll
                for (int c = 0; c < 8; c++) idx[c] = (((c & 1) ? hx1 : gx) ^ yz[c >> 1]) & p.mask;
            } else {
                gx &= kFloorRawMask; gy &= kFloorRawMask; gz &= kFloorRawMask;
                const uint32_t ix0 = gx * p.stride[0], iy0 = gy * p.stride[1], iz0 = gz * p.stride[2];
                const uint32_t yz[4] = {iy0 + iz0, iy0 + p.stride[1] + iz0, iy0 + iz0 + p.stride[2],
                                        iy0 + p.stride[1] + iz0 + p.stride[2]};
#pragma unroll
                for (int c = 0; c < 8; c++) idx[c] = (((c & 1) ? ix0 + p.stride[0] : ix0) + yz[c >> 1]) & p.mask;
            }
            const uint64_t base = (uint64_t)(uintptr_t)table + (uint64_t)p.offset * (uint32_t)(EW * 4);
#pragma unroll
            for (int c = 0; c < 8; c++) ldg_entry<EW>(base, idx[c], val[j][c]);
        }
#pragma unroll
        for (int j = 0; j < LV; j++) {
            if (HACC) {
                // packed-half accumulation: one HMUL2 / HFMA2 per corner and table with the weight broadcast to both halves
                // (one fp16 rounding per corner = the reference kernel's arithmetic, gridencoder.cu:142-165) instead of two
                // conversions + two FFMA
                uint32_t w2[8];
#pragma unroll
                for (int c = 0; c < 8; c++) asm("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(w2[c]) : "f"(wt[j][c]));
#pragma unroll
                for (int e = 0; e < EW; e++) {
                    uint32_t a2;
                    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(a2) : "r"(val[j][0][e]), "r"(w2[0]));
#pragma unroll
                    for (int c = 1; c < 8; c++) asm("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(a2) : "r"(val[j][c][e]), "r"(w2[c]));
                    rows[e][l0 + j] = in_range ? a2 : 0u;
                }
                continue;
            }
#pragma unroll
            for (int e = 0; e < EW; e++) {
                float ax = 0.f, ay = 0.f;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&val[j][c][e]));
                    ax = fmaf(wt[j][c], h.x, ax);
                    ay = fmaf(wt[j][c], h.y, ay);
                }
                rows[e][l0 + j] = in_range ? pack_h2(ax, ay) : 0u;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// gather_coop: LANE-PAIR cooperative gather of one warp tile (32 samples).
//
// Why: a fully scattered warp-wide load costs the L1/TEX pipe one wavefront per distinct 128-byte line it touches
// (round 1: 81 % L1/TEX pipe utilisation, 128 scattered 8-byte loads per sample). The two x-neighbour corners of a lattice
// cell are ADJACENT table entries — dense levels: x has stride 1; hashed levels: the x prime is 1, so
// idx(x+1) = idx(x) ^ 1 whenever x is even and differs only in the low bits otherwise — hence they share a 128-byte line
// 15 times out of 16. Here a load instruction serves 16 samples x 2 x-corners: lanes 2p and 2p+1 fetch the x / x+1 corner
// of the same (y, z) combination of sample p, so one instruction touches ~17 lines instead of 32; consecutive samples of
// ONE ray (the renderer's and the training path's sample order) additionally share lines on the coarse levels.
// Each lane accumulates its 4 (y, z) corners; one shuffle pair per (level, table) combines the two x halves; with two
// interleaved tables lane 2p ends up owning table 0's feature and lane 2p+1 table 1's, so both keep busy.
//
// Interpolation arithmetic: weights (wx*wy)*wz in fp32 (the reference's order), rounded ONCE to fp16, then
// fma.rn.f32.f16 (SASS FHFMA, new on sm_100): exact fp16 x fp16 products accumulated in fp32 — no fp16 -> fp32
// conversion instructions at all. Error vs fp32 weights <= 2^-11 * sum|w_c v_c| per feature, the size of the final fp16
// rounding of the feature itself (the reference accumulates all 8 terms in fp16, gridencoder.cu:142-165).
//   st(e, s, l0, words[LV]) : stores the packed fp16 feature pairs of table e, sample s, levels l0 .. l0 + LV - 1
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t f2h_bits(float v) {
    uint16_t h;
    asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    return h;
}
// ax += lo(v) * w, ay += hi(v) * w  (v = packed half2 table entry, w = fp16 weight): two FHFMA
__device__ __forceinline__ void fhfma2(float& ax, float& ay, uint32_t v, uint16_t w) {
    asm("{\n\t.reg .f16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tfma.rn.f32.f16 %0, lo, %3, %0;\n\tfma.rn.f32.f16 %1, hi, %3, %1;\n\t}"
        : "+f"(ax), "+f"(ay)
        : "r"(v), "h"(w));
}
// two fp32 weights -> one register of two fp16 weights with ONE conversion instruction (F2FP.F16.F32.PACK_AB), each rounded
// to nearest exactly as cvt.rn.f16.f32 would
__device__ __forceinline__ uint32_t f2h_pair(float lo, float hi) {
    uint32_t p;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi), "f"(lo));
    return p;
}
// the same two FHFMA with the weight taken from the low (HI = false) or high half of a weight pair
template <bool HI>
__device__ __forceinline__ void fhfma2_sel(float& ax, float& ay, uint32_t v, uint32_t wpair) {
    if (HI)
        asm("{\n\t.reg .f16 lo, hi, wl, wh;\n\tmov.b32 {lo, hi}, %2;\n\tmov.b32 {wl, wh}, %3;\n\tfma.rn.f32.f16 %0, lo, wh, %0;\n\t"
            "fma.rn.f32.f16 %1, hi, wh, %1;\n\t}"
            : "+f"(ax), "+f"(ay)
            : "r"(v), "r"(wpair));
    else
        asm("{\n\t.reg .f16 lo, hi, wl, wh;\n\tmov.b32 {lo, hi}, %2;\n\tmov.b32 {wl, wh}, %3;\n\tfma.rn.f32.f16 %0, lo, wl, %0;\n\t"
            "fma.rn.f32.f16 %1, hi, wl, %1;\n\t}"
            : "+f"(ax), "+f"(ay)
            : "r"(v), "r"(wpair));
}

// acc (half2) += v (half2) * w, w = the low (HI = false) / high half of a packed weight pair, broadcast to both halves:
// ptxas folds the broadcast into the operand selector of ONE HFMA2 (R.H0_H0 / R.H1_H1)
template <bool HI>
__device__ __forceinline__ void hfma2_bcast(uint32_t& acc, uint32_t v, uint32_t wpair) {
    if (HI)
        asm("{\n\t.reg .f16 lo, hi;\n\t.reg .b32 w;\n\tmov.b32 {lo, hi}, %2;\n\tmov.b32 w, {hi, hi};\n\tfma.rn.f16x2 %0, %1, w, %0;\n\t}"
            : "+r"(acc) : "r"(v), "r"(wpair));
    else
        asm("{\n\t.reg .f16 lo, hi;\n\t.reg .b32 w;\n\tmov.b32 {lo, hi}, %2;\n\tmov.b32 w, {lo, lo};\n\tfma.rn.f16x2 %0, %1, w, %0;\n\t}"
            : "+r"(acc) : "r"(v), "r"(wpair));
}
__device__ __forceinline__ uint32_t hmul2_bcast_lo(uint32_t v, uint32_t wpair) {
    uint32_t r;
    asm("{\n\t.reg .f16 lo, hi;\n\t.reg .b32 w;\n\tmov.b32 {lo, hi}, %2;\n\tmov.b32 w, {lo, lo};\n\tmul.rn.f16x2 %0, %1, w;\n\t}"
        : "=r"(r) : "r"(v), "r"(wpair));
    return r;
}

template <int EW, int LV, bool HACC = false, bool ZERO_OOR = true, typename StoreFn>
__device__ __forceinline__ void gather_coop(const void* __restrict__ table, const LevelParams* __restrict__ lp, float u, float v,
                                            float w, bool in_range, int lane, StoreFn st) {
    static_assert(EW == 1 || EW == 2, "one table or two interleaved tables");
    u = fminf(fmaxf(u, 0.f), 1.f); v = fminf(fmaxf(v, 0.f), 1.f); w = fminf(fmaxf(w, 0.f), 1.f);   // keeps the loads in bounds
    const uint32_t xsel = (uint32_t)lane & 1u;
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        const int s = half * 16 + (lane >> 1);
        const float us = __shfl_sync(0xffffffffu, u, s), vs = __shfl_sync(0xffffffffu, v, s), wsm = __shfl_sync(0xffffffffu, w, s);
        const bool inr = __shfl_sync(0xffffffffu, (int)in_range, s) != 0;
#pragma unroll 1
        for (int l0 = 0; l0 < 16; l0 += LV) {
            uint32_t val[LV][4][EW];
            uint32_t wt[LV][2];                            // fp16 weight pairs {corner 0, 1}, {corner 2, 3}
#pragma unroll
            for (int j = 0; j < LV; j++) {
                const LevelParams& p = lp[l0 + j];
                const float px = fmaf(us, p.scale, 0.5f), py = fmaf(vs, p.scale, 0.5f), pz = fmaf(wsm, p.scale, 0.5f);
                float fx0, fy0, fz0;
                uint32_t gx, gy, gz;                   // raw words (see floor_split): stripped in the dense branch only
                floor_split(px, fx0, gx); floor_split(py, fy0, gy); floor_split(pz, fz0, gz);
                const float rx = px - fx0, ry = py - fy0, rz = pz - fz0;
                const float wx = xsel ? rx : 1.f - rx;
                const float wxy0 = wx * (1.f - ry), wxy1 = wx * ry;        // (wx*wy)*wz: the reference's order
                // (fp16 weight products {w0, w1} = {wxy0, wxy1} * wz as two packed multiplies: 4.70 vs 4.72 ms per view, one more
                // rounding — measured, not kept)
                wt[j][0] = f2h_pair(wxy0 * (1.f - rz), wxy1 * (1.f - rz));
                wt[j][1] = f2h_pair(wxy0 * rz, wxy1 * rz);
                uint32_t idx[4];
                if (p.use_hash) {   // raw words: see gather_fast
                    const uint32_t hx = gx + xsel;
                    const uint32_t hy0 = gy * 2654435761u, hz0 = gz * 805459861u;
                    const uint32_t hy1 = hy0 + 2654435761u, hz1 = hz0 + 805459861u;
                    idx[0] = (hx ^ hy0 ^ hz0) & p.mask; idx[1] = (hx ^ hy1 ^ hz0) & p.mask;
                    idx[2] = (hx ^ hy0 ^ hz1) & p.mask; idx[3] = (hx ^ hy1 ^ hz1) & p.mask;
                } else {
                    const uint32_t hx = (gx & kFloorRawMask) + xsel;
                    gy &= kFloorRawMask; gz &= kFloorRawMask;
                    const uint32_t ix = hx * p.stride[0], iy0 = gy * p.stride[1], iz0 = gz * p.stride[2];
                    const uint32_t iy1 = iy0 + p.stride[1], iz1 = iz0 + p.stride[2];
                    idx[0] = (ix + iy0 + iz0) & p.mask; idx[1] = (ix + iy1 + iz0) & p.mask;
                    idx[2] = (ix + iy0 + iz1) & p.mask; idx[3] = (ix + iy1 + iz1) & p.mask;
                }
                const uint64_t base = (uint64_t)(uintptr_t)table + (uint64_t)p.offset * (uint32_t)(EW * 4);
#pragma unroll
                for (int c = 0; c < 4; c++) ldg_entry<EW>(base, idx[c], val[j][c]);
            }
            uint32_t words[LV];
#pragma unroll
            for (int j = 0; j < LV; j++) {
                if (HACC) {
                    // packed-half accumulation (HMUL2 + 3 HFMA2 per table instead of 8 FHFMA, the weight broadcast by the operand
                    // selector of the instruction; the partner's half arrives as ONE
                    // packed word): one fp16 rounding per corner, which is the reference kernel's own arithmetic
                    // (gridencoder.cu:142-165 accumulates in scalar_t = half)
                    uint32_t a2[EW];
#pragma unroll
                    for (int e = 0; e < EW; e++) {
                        a2[e] = hmul2_bcast_lo(val[j][0][e], wt[j][0]);
                        hfma2_bcast<true>(a2[e], val[j][1][e], wt[j][0]);
                        hfma2_bcast<false>(a2[e], val[j][2][e], wt[j][1]);
                        hfma2_bcast<true>(a2[e], val[j][3][e], wt[j][1]);
                    }
                    uint32_t mine, send;
                    if (EW == 2) { mine = xsel ? a2[EW - 1] : a2[0]; send = xsel ? a2[0] : a2[EW - 1]; }
                    else { mine = a2[0]; send = a2[0]; }
                    const uint32_t got = __shfl_xor_sync(0xffffffffu, send, 1);
                    uint32_t sum;
                    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(sum) : "r"(mine), "r"(got));
                    words[j] = (!ZERO_OOR || inr) ? sum : 0u;   // ZERO_OOR = false: the caller clamps every sample into the grid
                    continue;
                }
                float acc[EW][2];
#pragma unroll
                for (int e = 0; e < EW; e++) {
                    acc[e][0] = acc[e][1] = 0.f;
#pragma unroll
                    for (int c = 0; c < 4; c += 2) {
                        fhfma2_sel<false>(acc[e][0], acc[e][1], val[j][c][e], wt[j][c >> 1]);
                        fhfma2_sel<true>(acc[e][0], acc[e][1], val[j][c + 1][e], wt[j][c >> 1]);
                    }
                }
                float rx_, ry_;
                if (EW == 2) {
                    // lane 2p keeps table 0 and receives the partner's table-0 half; lane 2p+1 the same for table 1
                    const float sx = xsel ? acc[0][0] : acc[EW - 1][0], sy = xsel ? acc[0][1] : acc[EW - 1][1];
                    const float mx = xsel ? acc[EW - 1][0] : acc[0][0], my = xsel ? acc[EW - 1][1] : acc[0][1];
                    rx_ = mx + __shfl_xor_sync(0xffffffffu, sx, 1);
                    ry_ = my + __shfl_xor_sync(0xffffffffu, sy, 1);
                } else {
                    rx_ = acc[0][0] + __shfl_xor_sync(0xffffffffu, acc[0][0], 1);
                    ry_ = acc[0][1] + __shfl_xor_sync(0xffffffffu, acc[0][1], 1);
                }
                words[j] = inr ? pack_h2(rx_, ry_) : 0u;
            }
            if (EW == 2 || !xsel) st(EW == 2 ? (int)xsel : 0, s, l0, words);
        }
    }
}

// (Measured and rejected: batches whose four levels are all hashed as straight-line code — no branch per level, ptxas then
// interleaves the four address chains and clusters the 16 loads behind them — 4.90 vs 4.62 ms per view: the loads of a level are no longer issued while the next level's addresses are computed (they reach the first use later), more values live at the 128-register
// cap, more spills; the per-level branch is a useful fence.)
// (Measured and rejected: a software-pipelined version — batches of two levels, the 8 loads of batch i + 1 issued before batch i
// is interpolated, so that a warp always has gathers in flight instead of a burst of 16 followed by a drain — 5.12 vs 4.92 ms
// per 800x800 view: the other three warps of the scheduler already cover the drain, the extra live batch costs scheduling
// freedom at the 128-register cap.)
// (Measured and rejected, profiles/README.md "gather variants": computing the corner indices once for the density and
// palette grids — same geometry — and reading both tables with them cuts ~30 % of the gather's instructions but does not
// speed the renderer up: the gather is bound by L1/L2 load latency, the index arithmetic hides under it, and the second
// feature row costs shared memory, i.e. L1 capacity.)

// per-sample result of the field, held by the lane that owns the sample
struct FieldOut {
    float sigma;          // exp(logit) (NOT yet scaled by density_scale)
    float diffuse[3], view_dep[3];
    float off_rad[13];    // offsets (12) + radiance (1), bias added
    float omega[kNB];     // normalised blending weights
    float clip[kClipMax];
};

}  // namespace pnerf
