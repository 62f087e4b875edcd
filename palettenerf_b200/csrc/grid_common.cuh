// grid_common.cuh — per-level constants and corner indexing of the multiresolution hash grid, shared by the
// stand-alone encoder kernels (gridenc.cu) and the fused field / render kernels (fused.cu).
#pragma once
#include "common.cuh"

namespace pnerf {

constexpr int kMaxLevels = 32;

struct LevelParams {
    float scale;         // exp2f(level * S) * H - 1
    uint32_t stride[5];  // dense strides per dimension; 0 for dimensions the reference's loop never reaches
    uint32_t size;       // hashmap_size = offsets[l+1] - offsets[l]
    uint32_t offset;     // offsets[l] (in entries)
    uint32_t use_hash;   // gridtype == hash && (res+1)^D overflowed the table
    uint32_t mask;       // size - 1 if size is a power of two; all ones if indices cannot reach size (dense level); else 0
};

// ref: gridencoder.cu:54-72,97-99 — the index loop stops multiplying once stride > hashmap_size.
__device__ __forceinline__ void make_level(LevelParams& p, uint32_t level, const int32_t* offsets, float S, uint32_t H,
                                           uint32_t D, uint32_t gridtype, bool align_corners) {
    p.offset = (uint32_t)offsets[level];
    p.size = (uint32_t)offsets[level + 1] - (uint32_t)offsets[level];
    p.scale = exp2f(level * S) * H - 1.0f;
    const uint32_t resolution = (uint32_t)ceilf(p.scale) + 1;
    uint32_t stride = 1;
    for (uint32_t d = 0; d < 5; d++) {
        if (d < D && stride <= p.size) {
            p.stride[d] = stride;
            stride *= align_corners ? resolution : (resolution + 1);
        } else {
            p.stride[d] = 0;
        }
    }
    p.use_hash = (gridtype == 0 && stride > p.size) ? 1u : 0u;
    p.mask = ((p.size & (p.size - 1)) == 0) ? (p.size - 1) : 0u;
    // the mask-wrapping fast paths (fused_common.cuh::gather_fast / gather_coop) form hashed indices from raw float words whose
    // exponent bits sit at bit 24 and above: tables of more than 2^24 entries per level take the generic (modulo) path
    if (p.use_hash && p.mask > 0x00ffffffu) p.mask = 0u;
    // dense level whose full (res+1)^D lattice fits the table: every index is < stride <= size, so the reference's
    // `index % hashmap_size` (gridencoder.cu:71) is the identity and the integer division can be dropped.
    // NOT with align_corners: the stride multiplier is `resolution` there while an input of exactly 1.0 on an integer
    // scale reaches lattice coordinate `resolution` (pos_grid + 1), so an index can exceed stride — keep the modulo.
    if (!align_corners && !p.use_hash && stride <= p.size && D <= 5 && p.stride[D - 1] != 0) p.mask = 0xffffffffu;
}

__device__ __forceinline__ uint32_t wrap_index(uint32_t index, const LevelParams& p) {
    return p.mask ? (index & p.mask) : (index % p.size);
}

__device__ __forceinline__ void corner_setup(const LevelParams& p, float x, float y, float z, uint32_t (&idx)[8],
                                             float (&w)[8], bool align_corners, uint32_t* cell = nullptr) {
    const float half = align_corners ? 0.0f : 0.5f;
    float px = x * p.scale + half, py = y * p.scale + half, pz = z * p.scale + half;
    const float fx0 = floorf(px), fy0 = floorf(py), fz0 = floorf(pz);
    const uint32_t gx = (uint32_t)fx0, gy = (uint32_t)fy0, gz = (uint32_t)fz0;
    if (cell) { cell[0] = gx; cell[1] = gy; cell[2] = gz; }
    px -= (float)gx; py -= (float)gy; pz -= (float)gz;
    const float wx[2] = {1 - px, px}, wy[2] = {1 - py, py}, wz[2] = {1 - pz, pz};
    if (p.use_hash) {
        const uint32_t hx[2] = {gx, gx + 1};  // prime 1
        const uint32_t hy[2] = {gy * 2654435761u, (gy + 1) * 2654435761u};
        const uint32_t hz[2] = {gz * 805459861u, (gz + 1) * 805459861u};
#pragma unroll
        for (int c = 0; c < 8; c++) {
            idx[c] = wrap_index(hx[c & 1] ^ hy[(c >> 1) & 1] ^ hz[(c >> 2) & 1], p);
            w[c] = wx[c & 1] * wy[(c >> 1) & 1] * wz[(c >> 2) & 1];
        }
    } else {
        const uint32_t ix[2] = {gx * p.stride[0], (gx + 1) * p.stride[0]};
        const uint32_t iy[2] = {gy * p.stride[1], (gy + 1) * p.stride[1]};
        const uint32_t iz[2] = {gz * p.stride[2], (gz + 1) * p.stride[2]};
#pragma unroll
        for (int c = 0; c < 8; c++) {
            idx[c] = wrap_index(ix[c & 1] + iy[(c >> 1) & 1] + iz[(c >> 2) & 1], p);
            w[c] = wx[c & 1] * wy[(c >> 1) & 1] * wz[(c >> 2) & 1];
        }
    }
}


}  // namespace pnerf
