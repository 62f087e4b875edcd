// train_common.cuh — pieces shared by the fused TRAINING kernels of the palette field (fused_train.cu) and of the stage-1
// NeRF field (nerf_train.cu): the saved-activation unit format (one m16 x k16 fp16 block in mma A-fragment order = 512
// bytes, stored / loaded as four 128-byte coalesced rows), the activation-derivative pack, the d_enc row store and the job
// table of the split-K weight-gradient kernel (k_field_wgrad, defined once in fused_train.cu).
#pragma once
#include "fused_common.cuh"

namespace pnerf {

constexpr int kTrainWarps = 8;     // backward CTA: 8 warps, <= 255 registers
constexpr int kDStride = 88;       // halfs per row of the per-warp gradient staging tile (176 B: ldmatrix rows conflict-free)
enum DCol { DC_HEAD = 0, DC_VIEW = 32, DC_DIFF = 48, DC_CLIP = 64 };

__device__ __forceinline__ void st_unit(uint32_t* __restrict__ base, int unit, const uint32_t (&a)[4], int lane) {
    uint32_t* p = base + unit * 128 + lane;
    p[0] = a[0]; p[32] = a[1]; p[64] = a[2]; p[96] = a[3];
}
__device__ __forceinline__ void ld_unit(const uint32_t* __restrict__ base, int unit, uint32_t (&a)[4], int lane) {
    const uint32_t* p = base + unit * 128 + lane;
    a[0] = __ldg(p); a[1] = __ldg(p + 32); a[2] = __ldg(p + 64); a[3] = __ldg(p + 96);
}

template <int STRIDE>
__device__ __forceinline__ void ldmatrix_a_s(uint32_t (&a)[4], const __half* tile, int row0, int col0, int lane) {
    const __half* p = tile + (row0 + (lane & 7) + ((lane >> 3) & 1) * 8) * STRIDE + col0 + (lane >> 4) * 8;
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                 : "r"(addr));
}

__device__ __forceinline__ uint32_t movmatrix_t(uint32_t v) {
    uint32_t r;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;\n" : "=r"(r) : "r"(v));
    return r;
}

__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }

// accumulators of 8 n-tiles (a 16 x 64 block of dL/d(post-activation)) times the activation derivative taken from the
// saved post-activation fragments -> A fragments of dL/d(pre-activation), also stored as the layer's dY units
enum Deriv { DRV_RELU, DRV_ELU };
template <int DRV>
__device__ __forceinline__ void deriv_pack(const float (&c)[8][4], const uint32_t* __restrict__ xbase, int xslot,
                                           uint32_t* __restrict__ ybase, int yslot, uint32_t (&a)[4][4], int lane) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t h[4];
        ld_unit(xbase, xslot + j, h, lane);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float2 hv = unpack_h2(h[i]);
            const float g0 = c[2 * j + (i >> 1)][(i & 1) * 2], g1 = c[2 * j + (i >> 1)][(i & 1) * 2 + 1];
            float d0, d1;
            if (DRV == DRV_RELU) { d0 = hv.x > 0.f ? g0 : 0.f; d1 = hv.y > 0.f ? g1 : 0.f; }
            else { d0 = hv.x > 0.f ? g0 : g0 * (hv.x + 1.0f); d1 = hv.y > 0.f ? g1 : g1 * (hv.y + 1.0f); }   // ELU' = elu + 1
            a[j][i] = pack_h2(d0, d1);
        }
        st_unit(ybase, yslot + j, a[j], lane);
    }
}

// 16 x 32 fp32 block of grid-feature gradients -> d_enc rows (C-fragment layout: rows g / g+8, cols nt*8 + 2q, +1)
__device__ __forceinline__ void store_denc(float* __restrict__ d_enc, uint32_t s0, uint32_t M, const float (&c)[4][4], int lane) {
    const uint32_t r0 = s0 + (lane >> 2), r1 = r0 + 8;
    const int q = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
        if (r0 < M) *reinterpret_cast<float2*>(d_enc + (size_t)r0 * 32 + nt * 8 + 2 * q) = make_float2(c[nt][0], c[nt][1]);
        if (r1 < M) *reinterpret_cast<float2*>(d_enc + (size_t)r1 * 32 + nt * 8 + 2 * q) = make_float2(c[nt][2], c[nt][3]);
    }
}

struct TrainSmem {
    LevelParams lp[kMaxLevels];
    float palette[kNB * 3];
    uint32_t fast_wrap;   // every level wraps with a mask (see FusedSmem::fast_wrap)
    // followed by: uint2 weights[...]; per-warp scratch
};

// ---- weight-gradient jobs: dW[n_out][k_in] = sum_s dY[s][n_out] X[s][k_in]; one job = one layer ------------------------
// yslot / xslot: first saved unit of dY / X inside a half-tile, ny / ux: their widths in 16-column units,
// kpad: row stride of the layer's block in the packed fp32 gradient buffer, dwoff: its offset (floats)
struct WJob { uint16_t yslot, xslot, ny, ux, kpad, pad; uint32_t dwoff; };
struct WJobs { WJob j[16]; uint32_t n; };

__host__ inline void add_job_at(WJobs& J, int yslot, int ny, int xslot, int ux, int kpad, int dwoff) {
    WJob& w = J.j[J.n++];
    w.yslot = (uint16_t)yslot; w.xslot = (uint16_t)xslot; w.ny = (uint16_t)ny; w.ux = (uint16_t)ux;
    w.kpad = (uint16_t)kpad; w.pad = 0; w.dwoff = (uint32_t)dwoff;
}

// launches k_field_wgrad (fused_train.cu) for `J` on saved buffers with UX / UY units per half-tile; supported job shapes
// (ny, ux): (4,1) (4,2) (4,3) (4,4) (1,4) (2,1)
int launch_field_wgrad(const uint32_t* xbuf, const uint32_t* ybuf, uint32_t M, const int32_t* m_dev, uint32_t UX, uint32_t UY,
                       const WJobs& J, float* dwbuf, cudaStream_t stream, const char* what);

}  // namespace pnerf
