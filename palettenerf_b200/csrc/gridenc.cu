// gridenc.cu — multiresolution hash/tiled grid encoder, forward + backward, for B200 (sm_100a).
//
// Replaces gridencoder/src/gridencoder.cu:35-479 of the reference.
//
// What is different from the reference schedule (one thread per (point, level), level on blockIdx.y, [L,B,C]
// output + a permute copy, per-level re-read of the coordinates):
//   * the fast path (D=3, C=2, fp16 or fp32 tables — the only shape PaletteNeRF instantiates) is point-major:
//     a thread reads its coordinate once, walks all L levels with the 8 corner gathers of several levels in
//     flight, accumulates in fp32 and writes the final [B, L*C] row with 128-bit stores (layout BLC), or the
//     reference's [L,B,C] layout on request;
//   * per-level constants (scale, strides, table offset/size, hash-or-dense) are computed once per CTA into
//     shared memory with exactly the reference's fp32 expressions (exp2f(level*S)*H-1, ceil) so cell indices agree;
//   * the backward fast path reads the [B, L*C] gradient row once (128-bit loads) and issues packed
//     red.global.add.f16x2 / red.global.add.v2.f32 reductions (one per corner, both features at once).
// The tables (24 MiB fp16 / 48 MiB fp32 per grid) are L2-resident on B200 (126 MB L2); consecutive threads hold
// consecutive samples of a ray, so coarse-level corners hit L1.
#include <type_traits>

#include "common.cuh"
#include "grid_common.cuh"
#include "fused_common.cuh"   // gather_coop: lane-pair cooperative gather with FHFMA interpolation

namespace pnerf {

// ------------------------------------------------------------------------------------------------
// typed helpers
// ------------------------------------------------------------------------------------------------
template <typename T> struct Pair;  // two consecutive features of one entry
template <> struct Pair<__half> {
    using type = __half2;
    static __device__ __forceinline__ float2 load(const __half* p) {
        return __half22float2(__ldg(reinterpret_cast<const __half2*>(p)));
    }
    static __device__ __forceinline__ void red(__half* p, float a, float b) {
        atomicAdd(reinterpret_cast<__half2*>(p), __floats2half2_rn(a, b));
    }
};
template <> struct Pair<float> {
    using type = float2;
    static __device__ __forceinline__ float2 load(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
    static __device__ __forceinline__ void red(float* p, float a, float b) {
        atomicAdd(reinterpret_cast<float2*>(p), make_float2(a, b));  // red.global.add.v2.f32 (sm_90+)
    }
};

// the eight corner reductions of one cell. fp32 tables: the two corners that differ in x are neighbours in memory
// whenever the pair is 16-byte aligned (dense levels: x even; hashed levels: x even, since the x prime is 1 the two
// hashes differ in bit 0 only) — they go out as ONE red.global.add.v4.f32. The L2 reduction rate is per operation, not
// per byte, so this removes a quarter of the operations on average.
template <typename T>
__device__ __forceinline__ void red_cell(T* gg, const uint32_t (&idx)[8], const float2 (&acc)[8]) {
#pragma unroll
    for (int c = 0; c < 8; c++) Pair<T>::red(gg + (size_t)idx[c] * 2, acc[c].x, acc[c].y);
}
template <>
__device__ __forceinline__ void red_cell<float>(float* gg, const uint32_t (&idx)[8], const float2 (&acc)[8]) {
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
        const uint32_t i0 = idx[c], i1 = idx[c + 1];
        if ((i0 ^ i1) == 1u && (reinterpret_cast<uintptr_t>(gg) & 15u) == 0u) {
            const bool swap = (i0 & 1u) != 0u;          // i1 is the even (16-byte aligned) one
            const float2 lo = swap ? acc[c + 1] : acc[c], hi = swap ? acc[c] : acc[c + 1];
            atomicAdd(reinterpret_cast<float4*>(gg + (size_t)(i0 & ~1u) * 2), make_float4(lo.x, lo.y, hi.x, hi.y));
        } else {
            Pair<float>::red(gg + (size_t)i0 * 2, acc[c].x, acc[c].y);
            Pair<float>::red(gg + (size_t)i1 * 2, acc[c + 1].x, acc[c + 1].y);
        }
    }
}

template <typename T> struct AccOf { using type = float; };
template <> struct AccOf<double> { using type = double; };  // fp64 tables (gradcheck) accumulate in fp64 like the reference
template <typename T> __device__ __forceinline__ typename AccOf<T>::type to_float(T v) { return (typename AccOf<T>::type)v; }
template <> __device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_float(typename AccOf<T>::type v) { return (T)v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }

template <typename T> __device__ __forceinline__ void atomic_add_t(T* p, typename AccOf<T>::type v) { atomicAdd(p, (T)v); }
template <> __device__ __forceinline__ void atomic_add_t<__half>(__half* p, float v) { atomicAdd(p, __float2half_rn(v)); }

// ------------------------------------------------------------------------------------------------
// fast forward: D = 3, C = 2
// ------------------------------------------------------------------------------------------------
template <typename T, int LV>  // LV = levels processed per unrolled group
__global__ void __launch_bounds__(256) k_grid_fwd_d3c2(const float* __restrict__ inputs, const T* __restrict__ grid,
                                                       const int32_t* __restrict__ offsets, T* __restrict__ outputs,
                                                       uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                                       bool align_corners, bool layout_blc) {
    __shared__ LevelParams lp[kMaxLevels];
    if (threadIdx.x < L) make_level(lp[threadIdx.x], threadIdx.x, offsets, S, H, 3, gridtype, align_corners);
    __syncthreads();
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float x = inputs[(size_t)b * 3 + 0], y = inputs[(size_t)b * 3 + 1], z = inputs[(size_t)b * 3 + 2];
    const bool oob = (x < 0 || x > 1) || (y < 0 || y > 1) || (z < 0 || z > 1);

    for (uint32_t l0 = 0; l0 < L; l0 += LV) {
        float2 acc[LV];
#pragma unroll
        for (int j = 0; j < LV; j++) acc[j] = make_float2(0.f, 0.f);
        if (!oob) {
            float2 v[LV][8];
            float w[LV][8];
#pragma unroll
            for (int j = 0; j < LV; j++) {
                if (l0 + j < L) {
                    const LevelParams& p = lp[l0 + j];
                    uint32_t idx[8];
                    corner_setup(p, x, y, z, idx, w[j], align_corners);
                    const T* g = grid + (size_t)p.offset * 2;
#pragma unroll
                    for (int c = 0; c < 8; c++) v[j][c] = Pair<T>::load(g + (size_t)idx[c] * 2);
                }
            }
#pragma unroll
            for (int j = 0; j < LV; j++) {
                if (l0 + j < L) {
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        acc[j].x += w[j][c] * v[j][c].x;
                        acc[j].y += w[j][c] * v[j][c].y;
                    }
                }
            }
        }
        // store
        if (layout_blc) {
            T* out = outputs + (size_t)b * L * 2 + (size_t)l0 * 2;
            if (LV == 4 && l0 + LV <= L && (L % 4) == 0) {
                if (sizeof(T) == 2) {
                    uint4 pk;
                    __half2 h0 = __floats2half2_rn(acc[0].x, acc[0].y), h1 = __floats2half2_rn(acc[1 % LV].x, acc[1 % LV].y);
                    __half2 h2 = __floats2half2_rn(acc[2 % LV].x, acc[2 % LV].y), h3 = __floats2half2_rn(acc[3 % LV].x, acc[3 % LV].y);
                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(out) = pk;
                } else {
                    float4* o4 = reinterpret_cast<float4*>(out);
                    o4[0] = make_float4(acc[0].x, acc[0].y, acc[1 % LV].x, acc[1 % LV].y);
                    o4[1] = make_float4(acc[2 % LV].x, acc[2 % LV].y, acc[3 % LV].x, acc[3 % LV].y);
                }
            } else {
#pragma unroll
                for (int j = 0; j < LV; j++) {
                    if (l0 + j < L) {
                        out[j * 2 + 0] = from_float<T>(acc[j].x);
                        out[j * 2 + 1] = from_float<T>(acc[j].y);
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < LV; j++) {
                if (l0 + j < L) {
                    T* out = outputs + ((size_t)(l0 + j) * B + b) * 2;
                    out[0] = from_float<T>(acc[j].x);
                    out[1] = from_float<T>(acc[j].y);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// generic forward (any D <= 5, C <= 8, any dtype, optional dy_dx): one thread per (point, level)
// ref: gridencoder.cu:75-223
// ------------------------------------------------------------------------------------------------
template <uint32_t D>
__device__ __forceinline__ uint32_t grid_index(const LevelParams& p, const uint32_t (&pg)[D]) {
    uint32_t index;
    if (p.use_hash) {
        constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
        index = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) index ^= pg[d] * primes[d];
    } else {
        index = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) index += pg[d] * p.stride[d];
    }
    return wrap_index(index, p);
}

template <typename T, uint32_t D>
__global__ void __launch_bounds__(256) k_grid_fwd_generic(const float* __restrict__ inputs, const T* __restrict__ grid,
                                                          const int32_t* __restrict__ offsets, T* __restrict__ outputs,
                                                          uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H,
                                                          T* __restrict__ dy_dx, uint32_t gridtype, bool align_corners,
                                                          bool layout_blc) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t level = blockIdx.y;
    if (b >= B) return;
    LevelParams p;
    make_level(p, level, offsets, S, H, D, gridtype, align_corners);
    const float* in = inputs + (size_t)b * D;
    T* out = layout_blc ? outputs + ((size_t)b * L + level) * C : outputs + ((size_t)level * B + b) * C;
    T* dd = dy_dx ? dy_dx + ((size_t)b * L + level) * D * C : nullptr;

    bool oob = false;
    float pos[D];
    uint32_t pg[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        const float v = in[d];
        if (v < 0 || v > 1) oob = true;
        pos[d] = v * p.scale + (align_corners ? 0.0f : 0.5f);
        const float f = floorf(pos[d]);
        pg[d] = (uint32_t)f;
        pos[d] -= (float)pg[d];
    }
    if (oob) {
        for (uint32_t ch = 0; ch < C; ch++) out[ch] = from_float<T>(0.f);
        if (dd) for (uint32_t i = 0; i < D * C; i++) dd[i] = from_float<T>(0.f);
        return;
    }
    const T* g = grid + (size_t)p.offset * C;
    using A = typename AccOf<T>::type;
    A res[8];
#pragma unroll
    for (int ch = 0; ch < 8; ch++) res[ch] = 0;
#pragma unroll
    for (uint32_t c = 0; c < (1u << D); c++) {
        float w = 1;
        uint32_t pl[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            if ((c & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
            else { w *= pos[d]; pl[d] = pg[d] + 1; }
        }
        const uint32_t index = grid_index<D>(p, pl);
#pragma unroll
        for (uint32_t ch = 0; ch < 8; ch++)
            if (ch < C) res[ch] += w * to_float<T>(g[(size_t)index * C + ch]);
    }
#pragma unroll
    for (uint32_t ch = 0; ch < 8; ch++)
        if (ch < C) out[ch] = from_float<T>(res[ch]);

    if (dd) {  // ref: gridencoder.cu:182-222
#pragma unroll
        for (uint32_t gd = 0; gd < D; gd++) {
            A rg[8];
#pragma unroll
            for (int ch = 0; ch < 8; ch++) rg[ch] = 0;
#pragma unroll
            for (uint32_t c = 0; c < (1u << (D - 1)); c++) {
                float w = p.scale;
                uint32_t pl[D];
#pragma unroll
                for (uint32_t nd = 0; nd < D - 1; nd++) {
                    const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                    if ((c & (1u << nd)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
                    else { w *= pos[d]; pl[d] = pg[d] + 1; }
                }
                pl[gd] = pg[gd];
                const uint32_t il = grid_index<D>(p, pl);
                pl[gd] = pg[gd] + 1;
                const uint32_t ir = grid_index<D>(p, pl);
#pragma unroll
                for (uint32_t ch = 0; ch < 8; ch++)
                    if (ch < C) rg[ch] += w * (to_float<T>(g[(size_t)ir * C + ch]) - to_float<T>(g[(size_t)il * C + ch]));
            }
#pragma unroll
            for (uint32_t ch = 0; ch < 8; ch++)
                if (ch < C) dd[gd * C + ch] = from_float<T>(rg[ch]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fast backward: D = 3, C = 2, grad layout [B, L*2]; thread per (point, group of LV levels)
// ref: gridencoder.cu:226-313
// ------------------------------------------------------------------------------------------------
template <typename T, int LV>
__global__ void __launch_bounds__(256) k_grid_bwd_d3c2(const T* __restrict__ grad, const float* __restrict__ inputs,
                                                       const int32_t* __restrict__ offsets, T* __restrict__ grad_grid,
                                                       uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                                       bool align_corners, bool layout_blc) {
    __shared__ LevelParams lp[kMaxLevels];
    if (threadIdx.x < L) make_level(lp[threadIdx.x], threadIdx.x, offsets, S, H, 3, gridtype, align_corners);
    __syncthreads();
    const uint32_t groups = ceil_div(L, (uint32_t)LV);
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // level-group major: all points of one level group are adjacent in the grid, so one group's tables are hot in L2
    const uint32_t grp = (uint32_t)(tid / B), b = (uint32_t)(tid % B);
    if (grp >= groups) return;
    const float x = inputs[(size_t)b * 3 + 0], y = inputs[(size_t)b * 3 + 1], z = inputs[(size_t)b * 3 + 2];
    if ((x < 0 || x > 1) || (y < 0 || y > 1) || (z < 0 || z > 1)) return;
    const uint32_t l0 = grp * LV;

    float2 g[LV];
    if (layout_blc && LV == 4 && l0 + LV <= L && (L % 4) == 0) {
        const T* gp = grad + (size_t)b * L * 2 + (size_t)l0 * 2;
        if (sizeof(T) == 2) {
            const uint4 pk = __ldg(reinterpret_cast<const uint4*>(gp));
            const uint32_t u[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int j = 0; j < LV; j++) g[j] = __half22float2(*reinterpret_cast<const __half2*>(&u[j % 4]));
        } else {
            const float4 a = __ldg(reinterpret_cast<const float4*>(gp));
            const float4 c = __ldg(reinterpret_cast<const float4*>(gp) + 1);
            const float2 t[4] = {make_float2(a.x, a.y), make_float2(a.z, a.w), make_float2(c.x, c.y), make_float2(c.z, c.w)};
#pragma unroll
            for (int j = 0; j < LV; j++) g[j] = t[j % 4];
        }
    } else {
#pragma unroll
        for (int j = 0; j < LV; j++) {
            g[j] = make_float2(0.f, 0.f);
            if (l0 + j < L) {
                const T* gp = layout_blc ? grad + ((size_t)b * L + l0 + j) * 2 : grad + ((size_t)(l0 + j) * B + b) * 2;
                g[j] = make_float2(to_float<T>(gp[0]), to_float<T>(gp[1]));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < LV; j++) {
        if (l0 + j < L) {
            const LevelParams& p = lp[l0 + j];
            uint32_t idx[8];
            float w[8];
            corner_setup(p, x, y, z, idx, w, align_corners);
            T* gg = grad_grid + (size_t)p.offset * 2;
#pragma unroll
            for (int c = 0; c < 8; c++) Pair<T>::red(gg + (size_t)idx[c] * 2, w[c] * g[j].x, w[c] * g[j].y);
        }
    }
}

// run-length backward: D = 3, C = 2. One thread owns K consecutive points at ONE level and keeps the 8 corner
// contributions of the current cell in registers while consecutive points stay in that cell; it issues the 8 packed
// reductions only when the cell changes. Consecutive points of a training batch are consecutive samples of a ray
// (step 2*sqrt(3)/1024), so at the coarse levels (cells of 1/16 .. 1/300) a run covers the whole segment and the
// same-address contention that serialises the reference's atomics at those levels (4 920 entries receiving 8*B adds)
// disappears; at the fine levels every point is its own run and the kernel degenerates to one reduction per corner.
// Threads of a warp hold 16 levels x 2 segments: the [B, L*2] gradient row is read as a contiguous 64 B.
// TG = type of the incoming gradient, TA = type of the table the reductions go to (TA = float with TG = half: fp32
// workspace, see pnerf_grid_encode_backward_ws).
template <typename TG, typename TA, int K>
__global__ void __launch_bounds__(256) k_grid_bwd_runs(const TG* __restrict__ grad, const float* __restrict__ inputs,
                                                       const int32_t* __restrict__ offsets, TA* __restrict__ grad_grid,
                                                       uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                                       bool align_corners, bool layout_blc,
                                                       const int32_t* __restrict__ count_dev = nullptr, float bound = 0.f) {
    __shared__ LevelParams lp[kMaxLevels];
    if (threadIdx.x < L) make_level(lp[threadIdx.x], threadIdx.x, offsets, S, H, 3, gridtype, align_corners);
    __syncthreads();
    if (count_dev) B = min(B, (uint32_t)__ldg(count_dev));
    // grid-stride over (segment, level) work items: with a device-side count the launch is sized for the capacity
    const uint64_t n_items = ceil_div<uint64_t>(B, K) * L, stride = (uint64_t)gridDim.x * blockDim.x;
#pragma unroll 1
    for (uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; tid < n_items; tid += stride) {
        const uint32_t level = (uint32_t)(tid % L);
        const uint64_t b0 = (tid / L) * K;
        const LevelParams& p = lp[level];
        TA* gg = grad_grid + (size_t)p.offset * 2;

        bool have = false;
        uint32_t cur_cell[3] = {0, 0, 0}, cur_idx[8];
        float2 acc[8];
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            const uint64_t b = b0 + k;
            if (b >= B) break;
            float x = inputs[b * 3 + 0], y = inputs[b * 3 + 1], z = inputs[b * 3 + 2];
            if (bound > 0.f) {   // world coordinates: the [0,1] map of GridEncoder.forward (gridencoder/grid.py:142), same fp32 ops
                x = (x + bound) / (2 * bound); y = (y + bound) / (2 * bound); z = (z + bound) / (2 * bound);
            }
            if ((x < 0 || x > 1) || (y < 0 || y > 1) || (z < 0 || z > 1)) continue;
            const TG* gp = layout_blc ? grad + (b * L + level) * 2 : grad + ((size_t)level * B + b) * 2;
            const float2 g = Pair<TG>::load(gp);
            uint32_t idx[8], cell[3];
            float w[8];
            corner_setup(p, x, y, z, idx, w, align_corners, cell);
            const bool same = have && cell[0] == cur_cell[0] && cell[1] == cur_cell[1] && cell[2] == cur_cell[2];
            if (!same) {
                if (have) {
#pragma unroll
                    red_cell<TA>(gg, cur_idx, acc);
                }
#pragma unroll
                for (int c = 0; c < 8; c++) { cur_idx[c] = idx[c]; acc[c] = make_float2(w[c] * g.x, w[c] * g.y); }
                cur_cell[0] = cell[0]; cur_cell[1] = cell[1]; cur_cell[2] = cell[2];
                have = true;
            } else {
#pragma unroll
                for (int c = 0; c < 8; c++) { acc[c].x += w[c] * g.x; acc[c].y += w[c] * g.y; }
            }
        }
        if (have) {
#pragma unroll
            red_cell<TA>(gg, cur_idx, acc);
        }
    }
}

// generic backward: thread per (point, level), scalar atomics per feature
template <typename T, uint32_t D>
__global__ void __launch_bounds__(256) k_grid_bwd_generic(const T* __restrict__ grad, const float* __restrict__ inputs,
                                                          const int32_t* __restrict__ offsets, T* __restrict__ grad_grid,
                                                          uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H,
                                                          uint32_t gridtype, bool align_corners, bool layout_blc) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t level = blockIdx.y;
    if (b >= B) return;
    LevelParams p;
    make_level(p, level, offsets, S, H, D, gridtype, align_corners);
    const float* in = inputs + (size_t)b * D;
    const T* gp = layout_blc ? grad + ((size_t)b * L + level) * C : grad + ((size_t)level * B + b) * C;
    float pos[D];
    uint32_t pg[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        const float v = in[d];
        if (v < 0 || v > 1) return;
        pos[d] = v * p.scale + (align_corners ? 0.0f : 0.5f);
        const float f = floorf(pos[d]);
        pg[d] = (uint32_t)f;
        pos[d] -= (float)pg[d];
    }
    using A = typename AccOf<T>::type;
    A gc[8];
#pragma unroll
    for (uint32_t ch = 0; ch < 8; ch++) gc[ch] = (ch < C) ? to_float<T>(gp[ch]) : (A)0;
    T* gg = grad_grid + (size_t)p.offset * C;
#pragma unroll
    for (uint32_t c = 0; c < (1u << D); c++) {
        float w = 1;
        uint32_t pl[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            if ((c & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
            else { w *= pos[d]; pl[d] = pg[d] + 1; }
        }
        const uint32_t index = grid_index<D>(p, pl);
#pragma unroll
        for (uint32_t ch = 0; ch < 8; ch++)
            if (ch < C) atomic_add_t<T>(gg + (size_t)index * C + ch, w * gc[ch]);
    }
}

// grad_inputs[b,d] = sum_{l,ch} grad[l,b,ch] * dy_dx[b,l,d,ch]   (ref: gridencoder.cu:316-342)
template <typename T>
__global__ void __launch_bounds__(256) k_grid_input_bwd(const T* __restrict__ grad, const T* __restrict__ dy_dx,
                                                        T* __restrict__ grad_inputs, uint32_t B, uint32_t D, uint32_t C,
                                                        uint32_t L, bool layout_blc) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const T* dd = dy_dx + (size_t)b * L * D * C;
    typename AccOf<T>::type result = 0;
    for (uint32_t l = 0; l < L; l++) {
        const T* gp = layout_blc ? grad + ((size_t)b * L + l) * C : grad + ((size_t)l * B + b) * C;
        for (uint32_t ch = 0; ch < C; ch++) result += to_float<T>(gp[ch]) * to_float<T>(dd[((size_t)l * D + d) * C + ch]);
    }
    grad_inputs[t] = from_float<T>(result);
}

// ------------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// fast forward for the configuration every PaletteNeRF grid uses: fp16 table, D = 3, C = 2, L = 16, hash grid, not
// align_corners, [B, L*C] output. A warp owns 32 points; the gather is the lane-pair cooperative one of the fused kernels
// (gather_coop: the two x-neighbour corners of a cell are adjacent table entries, so the pair of lanes that fetch them share
// a 128-byte line — a load instruction touches ~16 lines instead of 32, which is what bounds this kernel: round 1's
// thread-per-point version ran at 86.6 % of the L1/TEX pipe) with exact fp16 x fp16 products accumulated in fp32 (FHFMA).
// Rows are staged in shared memory and leave as one contiguous 2 KB burst per warp.
// Levels that cannot wrap with a mask (a hashed level whose size is not a power of two) take the per-thread path below.
// ------------------------------------------------------------------------------------------------
constexpr int kCoopWarps = 8;
__global__ void __launch_bounds__(kCoopWarps * 32) k_grid_fwd_coop_h(const float* __restrict__ inputs, const __half* __restrict__ grid,
                                                                     const int32_t* __restrict__ offsets, __half* __restrict__ outputs,
                                                                     uint32_t B, float S, uint32_t H, bool layout_blc) {
    __shared__ LevelParams lp[16];
    __shared__ __align__(16) uint32_t stage[kCoopWarps][32][16 + 4];      // +4 words: rows start in different bank groups
    __shared__ int slow_s;
    if (threadIdx.x < 16) make_level(lp[threadIdx.x], threadIdx.x, offsets, S, H, 3, 0, false);
    const int slow = __syncthreads_or(threadIdx.x < 16 && lp[threadIdx.x].mask == 0u);
    if (threadIdx.x == 0) slow_s = slow;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t n_tiles = ceil_div(B, 32u);
    for (uint32_t tile = blockIdx.x * kCoopWarps + wid; tile < n_tiles; tile += gridDim.x * kCoopWarps) {
        const uint32_t b = tile * 32 + lane;
        const bool active = b < B;
        float x = 0.f, y = 0.f, z = 0.f;
        if (active) { x = inputs[(size_t)b * 3]; y = inputs[(size_t)b * 3 + 1]; z = inputs[(size_t)b * 3 + 2]; }
        const bool in_range = active && !((x < 0 || x > 1) || (y < 0 || y > 1) || (z < 0 || z > 1));
        uint32_t (*rows)[20] = stage[wid];
        if (!slow) {
            auto st = [rows](int, int s, int l0, const uint32_t (&wd)[4]) {
                *reinterpret_cast<uint4*>(&rows[s][l0]) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            };
            gather_coop<1, 4>(grid, lp, x, y, z, in_range, lane, st);
        } else {
            for (int l = 0; l < 16; l++) {
                float2 acc = make_float2(0.f, 0.f);
                if (in_range) {
                    uint32_t idx[8];
                    float w[8];
                    corner_setup(lp[l], x, y, z, idx, w, false);
                    const __half2* g = reinterpret_cast<const __half2*>(grid) + lp[l].offset;
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const float2 v = __half22float2(__ldg(g + idx[c]));
                        acc.x += w[c] * v.x; acc.y += w[c] * v.y;
                    }
                }
                rows[lane][l] = pack_h2(acc.x, acc.y);
            }
        }
        __syncwarp();
        if (layout_blc) {
            // 32 rows x 64 B = 128 uint4, contiguous in the output: lane i writes uint4 i, i + 32, i + 64, i + 96
            uint4* out = reinterpret_cast<uint4*>(outputs + (size_t)tile * 32 * 32);
            const uint32_t rows_here = min(32u, B - tile * 32);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int q = lane + 32 * k, r = q >> 2, c4 = q & 3;
                if ((uint32_t)r < rows_here) out[q] = *reinterpret_cast<const uint4*>(&rows[r][c4 * 4]);
            }
        } else if (active) {
            // the reference's [L, B, C] layout: per level the warp writes 32 consecutive feature pairs (128 B)
            uint32_t* out = reinterpret_cast<uint32_t*>(outputs);
#pragma unroll
            for (int l = 0; l < 16; l++) out[(size_t)l * B + b] = rows[lane][l];
        }
        __syncwarp();
    }
}

template <typename T>
int grid_forward_t(const float* inputs, const T* emb, const int32_t* offsets, T* outputs, uint32_t B, uint32_t D,
                   uint32_t C, uint32_t L, float S, uint32_t H, T* dy_dx, uint32_t gridtype, bool align, bool blc,
                   cudaStream_t s) {
    if constexpr (std::is_same<T, __half>::value) {
        if (D == 3 && C == 2 && L == 16 && dy_dx == nullptr && gridtype == 0 && !align) {
            const uint32_t grid = min(ceil_div(ceil_div(B, 32u), (uint32_t)kCoopWarps), 8u * (uint32_t)kNumSMs);
            k_grid_fwd_coop_h<<<grid, kCoopWarps * 32, 0, s>>>(inputs, emb, offsets, outputs, B, S, H, blc);
            return check_launch("grid_encode_forward");
        }
    }
    if (D == 3 && C == 2 && dy_dx == nullptr && sizeof(T) <= 4) {
        if constexpr (!std::is_same<T, double>::value) {
            k_grid_fwd_d3c2<T, 4><<<ceil_div(B, 256u), 256, 0, s>>>(inputs, emb, offsets, outputs, B, L, S, H, gridtype,
                                                                   align, blc);
            return check_launch("grid_encode_forward");
        }
    }
    const dim3 grid(ceil_div(B, 256u), L, 1);
    switch (D) {
        case 1: k_grid_fwd_generic<T, 1><<<grid, 256, 0, s>>>(inputs, emb, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, align, blc); break;
        case 2: k_grid_fwd_generic<T, 2><<<grid, 256, 0, s>>>(inputs, emb, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, align, blc); break;
        case 3: k_grid_fwd_generic<T, 3><<<grid, 256, 0, s>>>(inputs, emb, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, align, blc); break;
        case 4: k_grid_fwd_generic<T, 4><<<grid, 256, 0, s>>>(inputs, emb, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, align, blc); break;
        case 5: k_grid_fwd_generic<T, 5><<<grid, 256, 0, s>>>(inputs, emb, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, align, blc); break;
        default: return PNERF_ERR_UNSUPPORTED;
    }
    return check_launch("grid_encode_forward");
}

template <typename T>
int grid_backward_t(const T* grad, const float* inputs, const int32_t* offsets, T* grad_emb, uint32_t B, uint32_t D,
                    uint32_t C, uint32_t L, float S, uint32_t H, const T* dy_dx, T* grad_inputs, uint32_t gridtype,
                    bool align, bool blc, cudaStream_t s) {
    bool done = false;
    if (D == 3 && C == 2) {
        if constexpr (!std::is_same<T, double>::value) {
            constexpr int K = 8;
            const uint64_t threads = ceil_div<uint64_t>(B, K) * L;
            k_grid_bwd_runs<T, T, K><<<(uint32_t)ceil_div<uint64_t>(threads, 256), 256, 0, s>>>(
                grad, inputs, offsets, grad_emb, B, L, S, H, gridtype, align, blc);
            done = true;
        }
    }
    if (!done) {
        const dim3 grid(ceil_div(B, 256u), L, 1);
        switch (D) {
            case 1: k_grid_bwd_generic<T, 1><<<grid, 256, 0, s>>>(grad, inputs, offsets, grad_emb, B, C, L, S, H, gridtype, align, blc); break;
            case 2: k_grid_bwd_generic<T, 2><<<grid, 256, 0, s>>>(grad, inputs, offsets, grad_emb, B, C, L, S, H, gridtype, align, blc); break;
            case 3: k_grid_bwd_generic<T, 3><<<grid, 256, 0, s>>>(grad, inputs, offsets, grad_emb, B, C, L, S, H, gridtype, align, blc); break;
            case 4: k_grid_bwd_generic<T, 4><<<grid, 256, 0, s>>>(grad, inputs, offsets, grad_emb, B, C, L, S, H, gridtype, align, blc); break;
            case 5: k_grid_bwd_generic<T, 5><<<grid, 256, 0, s>>>(grad, inputs, offsets, grad_emb, B, C, L, S, H, gridtype, align, blc); break;
            default: return PNERF_ERR_UNSUPPORTED;
        }
    }
    if (dy_dx && grad_inputs) {
        k_grid_input_bwd<T><<<ceil_div(B * D, 256u), 256, 0, s>>>(grad, dy_dx, grad_inputs, B, D, C, L, blc);
    }
    return check_launch("grid_encode_backward");
}

// grad_embeddings(fp16) += round(workspace(fp32)); 8 entries (4 pairs) per thread, 128-bit loads
__global__ void __launch_bounds__(256) k_grid_fold_ws(const float4* __restrict__ ws, uint2* __restrict__ out, uint64_t n4) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = ld_stream4(ws + i);
        uint2 o = out[i];
        __half2 a = *reinterpret_cast<__half2*>(&o.x), b = *reinterpret_cast<__half2*>(&o.y);
        a = __hadd2(a, __floats2half2_rn(v.x, v.y));
        b = __hadd2(b, __floats2half2_rn(v.z, v.w));
        o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
        out[i] = o;
    }
}

}  // namespace pnerf

using namespace pnerf;

extern "C" {

int pnerf_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs,
                              uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, void* dy_dx,
                              uint32_t gridtype, int align_corners, int dtype, int out_layout, void* stream) {
    if (B == 0) return PNERF_OK;
    PNERF_REQUIRE(inputs && embeddings && offsets && outputs);
    PNERF_REQUIRE(gridtype <= 1 && (out_layout == PNERF_LAYOUT_LBC || out_layout == PNERF_LAYOUT_BLC));
    if (D < 1 || D > 5 || !(C == 1 || C == 2 || C == 4 || C == 8) || L < 1 || L > (uint32_t)kMaxLevels)
        return PNERF_ERR_UNSUPPORTED;  // ref: gridencoder.cu:354,372 throws for the same shapes
    cudaStream_t s = (cudaStream_t)stream;
    const bool blc = out_layout == PNERF_LAYOUT_BLC, al = align_corners != 0;
    switch (dtype) {
        case PNERF_F16: return grid_forward_t<__half>(inputs, (const __half*)embeddings, offsets, (__half*)outputs, B, D, C, L, S, H, (__half*)dy_dx, gridtype, al, blc, s);
        case PNERF_F32: return grid_forward_t<float>(inputs, (const float*)embeddings, offsets, (float*)outputs, B, D, C, L, S, H, (float*)dy_dx, gridtype, al, blc, s);
        case PNERF_F64: return grid_forward_t<double>(inputs, (const double*)embeddings, offsets, (double*)outputs, B, D, C, L, S, H, (double*)dy_dx, gridtype, al, blc, s);
        default: return PNERF_ERR_INVALID_ARG;
    }
}

int pnerf_grid_encode_backward_counted(const void* grad, const float* inputs, const int32_t* offsets, void* grad_embeddings,
                                       uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                       int dtype, int grad_layout, const int32_t* count_dev, float bound, void* stream) {
    if (B == 0) return PNERF_OK;
    PNERF_REQUIRE(grad && inputs && offsets && grad_embeddings && count_dev);
    PNERF_REQUIRE(gridtype <= 1 && (grad_layout == PNERF_LAYOUT_LBC || grad_layout == PNERF_LAYOUT_BLC));
    if (L < 1 || L > (uint32_t)kMaxLevels) return PNERF_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const bool blc = grad_layout == PNERF_LAYOUT_BLC, al = align_corners != 0;
    constexpr int K = 8;
    const uint64_t threads = ceil_div<uint64_t>(B, K) * L;
    const uint64_t blocks = ceil_div<uint64_t>(threads, 256);
    const uint32_t grid = (uint32_t)(blocks < 8ull * kNumSMs ? blocks : 8ull * kNumSMs);   // grid-stride kernel
    if (dtype == PNERF_F32)
        k_grid_bwd_runs<float, float, K><<<grid, 256, 0, s>>>((const float*)grad, inputs, offsets, (float*)grad_embeddings, B, L, S, H,
                                                      gridtype, al, blc, count_dev, bound);
    else if (dtype == PNERF_F16)
        k_grid_bwd_runs<__half, __half, K><<<grid, 256, 0, s>>>((const __half*)grad, inputs, offsets, (__half*)grad_embeddings, B, L, S,
                                                       H, gridtype, al, blc, count_dev, bound);
    else
        return PNERF_ERR_UNSUPPORTED;
    return check_launch("grid_encode_backward_counted");
}

int pnerf_grid_encode_backward_ws(const void* grad, const float* inputs, const int32_t* offsets, void* grad_embeddings,
                                  float* workspace, uint64_t n_entries, uint32_t B, uint32_t L, float S, uint32_t H,
                                  uint32_t gridtype, int align_corners, int grad_layout, void* stream) {
    if (B == 0) return PNERF_OK;
    PNERF_REQUIRE(grad && inputs && offsets && grad_embeddings && workspace && n_entries > 0);
    PNERF_REQUIRE((n_entries % 2) == 0);   // 2 entries (4 floats) per 128-bit access; GridEncoder pads levels to 8 entries
    PNERF_REQUIRE(gridtype <= 1 && (grad_layout == PNERF_LAYOUT_LBC || grad_layout == PNERF_LAYOUT_BLC));
    if (L < 1 || L > (uint32_t)kMaxLevels) return PNERF_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const bool blc = grad_layout == PNERF_LAYOUT_BLC, al = align_corners != 0;
    if (cudaMemsetAsync(workspace, 0, n_entries * 2 * sizeof(float), s) != cudaSuccess) return check_launch("grid_encode_backward_ws");
    constexpr int K = 8;
    const uint64_t threads = ceil_div<uint64_t>(B, K) * L;
    k_grid_bwd_runs<__half, float, K><<<(uint32_t)ceil_div<uint64_t>(threads, 256), 256, 0, s>>>(
        (const __half*)grad, inputs, offsets, workspace, B, L, S, H, gridtype, al, blc);
    const uint64_t n4 = n_entries / 2;
    const uint64_t blocks = ceil_div<uint64_t>(n4, 256);
    k_grid_fold_ws<<<(uint32_t)(blocks < 16ull * kNumSMs ? blocks : 16ull * kNumSMs), 256, 0, s>>>(
        (const float4*)workspace, (uint2*)grad_embeddings, n4);
    return check_launch("grid_encode_backward_ws");
}

int pnerf_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings, const int32_t* offsets,
                               void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                               uint32_t H, const void* dy_dx, void* grad_inputs, uint32_t gridtype, int align_corners,
                               int dtype, int grad_layout, void* stream) {
    if (B == 0) return PNERF_OK;
    (void)embeddings;  // the gradient does not depend on the table values (ref kernel takes but never reads them)
    PNERF_REQUIRE(grad && inputs && offsets && grad_embeddings);
    PNERF_REQUIRE(gridtype <= 1 && (grad_layout == PNERF_LAYOUT_LBC || grad_layout == PNERF_LAYOUT_BLC));
    if (D < 1 || D > 5 || !(C == 1 || C == 2 || C == 4 || C == 8) || L < 1 || L > (uint32_t)kMaxLevels)
        return PNERF_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const bool blc = grad_layout == PNERF_LAYOUT_BLC, al = align_corners != 0;
    switch (dtype) {
        case PNERF_F16: return grid_backward_t<__half>((const __half*)grad, inputs, offsets, (__half*)grad_embeddings, B, D, C, L, S, H, (const __half*)dy_dx, (__half*)grad_inputs, gridtype, al, blc, s);
        case PNERF_F32: return grid_backward_t<float>((const float*)grad, inputs, offsets, (float*)grad_embeddings, B, D, C, L, S, H, (const float*)dy_dx, (float*)grad_inputs, gridtype, al, blc, s);
        case PNERF_F64: return grid_backward_t<double>((const double*)grad, inputs, offsets, (double*)grad_embeddings, B, D, C, L, S, H, (const double*)dy_dx, (double*)grad_inputs, gridtype, al, blc, s);
        default: return PNERF_ERR_INVALID_ARG;
    }
}

}  // extern "C"
