// march_common.cuh — Morton bit tricks and the per-ray marcher shared by raymarch.cu and the fused render kernel.
#pragma once
#include "common.cuh"

namespace pnerf {

// ------------------------------------------------------------------------------------------------
// bit tricks (ref: raymarching.cu:59-84)
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {
    // 10 input bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton_encode(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t compact3(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

// exponent e with v = m * 2^e, m in [0.5, 1) — frexpf semantics for the values that matter here
// (normal floats; zero/denormals give e <= 0 which every caller clamps to 0, as frexpf's would be).
__device__ __forceinline__ int frexp_exponent(float v) {
    return (int)((__float_as_uint(v) >> 23) & 0xffu) - 126;
}

// ------------------------------------------------------------------------------------------------
// Per-ray marcher. One instance per thread; `probe(t)` evaluates the lattice point t and either reports an
// occupied sample or advances t past the empty voxel. Arithmetic follows raymarching.cu:364-403 exactly.
// ------------------------------------------------------------------------------------------------
struct Marcher {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, dt_gamma, dt_min, dt_max;
    float rH, Hf, halfH, Hm1, H3f, rbound;
    int Ci1;
    const uint8_t* __restrict__ grid;

    __device__ __forceinline__ void init(const float* __restrict__ o, const float* __restrict__ d, float bound_,
                                         float dt_gamma_, uint32_t max_steps, uint32_t C, uint32_t H,
                                         const uint8_t* __restrict__ grid_) {
        ox = o[0]; oy = o[1]; oz = o[2];
        dx = d[0]; dy = d[1]; dz = d[2];
        rdx = 1 / dx; rdy = 1 / dy; rdz = 1 / dz;
        bound = bound_;
        dt_gamma = dt_gamma_;
        const float two_sqrt3 = 2 * 1.7320508075688772f;
        dt_min = two_sqrt3 / max_steps;
        dt_max = two_sqrt3 * (1u << (C - 1)) / H;
        Hf = (float)H;
        rH = 1 / Hf;
        halfH = 0.5f * Hf;  // exact; (0.5 * v * H) in double rounds once, same as v * halfH in float
        Hm1 = (float)(H - 1);
        H3f = (float)(H * H * H);
        Ci1 = (int)C - 1;
        rbound = 1 / bound_;
        grid = grid_;
    }

    __device__ __forceinline__ float step_size(float t) const { return clampf(t * dt_gamma, dt_min, dt_max); }

    // Classify lattice point t: returns true if (x,y,z) is an occupied sample; otherwise `tt` receives the ray
    // parameter at which the ray leaves the empty voxel (ref: raymarching.cu:364-403). Does not advance t.
    // sample position and step at ray parameter t (ref: raymarching.cu:366-371); ONE definition, so the list-driven
    // writer (k_march_train_emit) produces the bits the walk produced
    __device__ __forceinline__ void position(float t, float& x, float& y, float& z, float& dt) const {
        x = clampf(ox + t * dx, -bound, bound);
        y = clampf(oy + t * dy, -bound, bound);
        z = clampf(oz + t * dz, -bound, bound);
        dt = step_size(t);
    }

    // first lattice point of a training ray (ref: raymarching.cu:352-355: t = near; t += clamp(t * dt_gamma, ..) * noise)
    __device__ __forceinline__ float first_t(float near, float noise) const {
        float t = near;
        t += step_size(t) * noise;
        return t;
    }

    // Ray parameter at which the ray leaves the box `occ` = world bounds of every occupied cell, padded by one cell
    // (pnerf_occupied_bounds). Behind it no lattice point can fall into an occupied cell, so a walk may stop there
    // instead of at `far`: the samples found are the same, the empty tail behind the object is not marched.
    // A ray that misses the box gets -1 (no samples at all).
    __device__ __forceinline__ float occupied_exit(const float* __restrict__ occ) const {
        float t_in = -3.402823466e+38f, t_out = 3.402823466e+38f;
        const float o[3] = {ox, oy, oz}, rd[3] = {rdx, rdy, rdz}, d[3] = {dx, dy, dz};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float lo = occ[k], hi = occ[3 + k];
            if (d[k] == 0.f) {
                if (o[k] < lo || o[k] > hi) return -1.f;
            } else {
                const float a = (lo - o[k]) * rd[k], b = (hi - o[k]) * rd[k];
                t_in = fmaxf(t_in, fminf(a, b));
                t_out = fminf(t_out, fmaxf(a, b));
            }
        }
        return (t_in <= t_out) ? t_out : -1.f;
    }

    // STRETCH (constant step only, dt_gamma == 0): for an OCCUPIED point `tt` receives a ray parameter up to which the ray
    // provably stays inside the point's voxel — the exit through the voxel shrunk by 1e-3 of its size on the exit sides,
    // four orders of magnitude more than the fp32 rounding of position and index (~3e-5 voxel). Every lattice point
    // below it lands in the same voxel at the same cascade (the cascade can only change across |x| = 2^k planes, which
    // are voxel faces unless the cascade is capped by the bound; the step-size term is constant), so it is occupied without
    // being probed.
    template <bool STRETCH = false>
    __device__ __forceinline__ bool probe_point(float t, float& x, float& y, float& z, float& dt, float& tt) const {
        position(t, x, y, z, dt);

        // cascade from position and from step size (ref: raymarching.cu:45-57)
        // (integer clamps: the reference clamps the exponents as floats, fminf(C-1, fmaxf(0, e)) — same integers)
        const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
        const int level = min(max(max(frexp_exponent(mx), frexp_exponent(dt * Hf * 0.5f)), 0), Ci1);

        // mip_bound = min(2^level, bound) and its reciprocal without scalbnf / a division: 2^level and 2^-level are
        // exponent-field constructions (exact), and 1 / bound is the same correctly-rounded quotient computed once
        const float pow2 = __int_as_float((127 + level) << 23);
        const bool capped = pow2 > bound;
        const float mip_bound = capped ? bound : pow2;
        const float mip_rbound = capped ? rbound : __int_as_float((127 - level) << 23);

        const int nx = (int)clampf((x * mip_rbound + 1) * halfH, 0.0f, Hm1);
        const int ny = (int)clampf((y * mip_rbound + 1) * halfH, 0.0f, Hm1);
        const int nz = (int)clampf((z * mip_rbound + 1) * halfH, 0.0f, Hm1);

        // the reference forms this index in fp32 (level * H3 + morton); keep its rounding behaviour
        const uint32_t index = (uint32_t)((float)level * H3f + (float)morton_encode(nx, ny, nz));
        const bool occ = grid[index >> 3] & (1u << (index & 7u));
        if (occ && !STRETCH) return true;

        // distance to the exit face of this voxel along each axis
        const float sx = copysignf(1.0f, dx), sy = copysignf(1.0f, dy), sz = copysignf(1.0f, dz);
        if (STRETCH && occ) {
            const float ux = (((nx + 0.5f + 0.499f * sx) * rH * 2 - 1) * mip_bound - x) * rdx;
            const float uy = (((ny + 0.5f + 0.499f * sy) * rH * 2 - 1) * mip_bound - y) * rdy;
            const float uz = (((nz + 0.5f + 0.499f * sz) * rH * 2 - 1) * mip_bound - z) * rdz;
            // (NaN from 0 * inf on an axis the ray does not move along is dropped by fminf. bound < 2^level: the voxel grid is
            // scaled to the bound, the |x| = 2^k planes are no longer faces — no stretch there)
            tt = capped ? t : t + fminf(ux, fminf(uy, uz));
            return true;
        }
        const float tx = (((nx + 0.5f + 0.5f * sx) * rH * 2 - 1) * mip_bound - x) * rdx;
        const float ty = (((ny + 0.5f + 0.5f * sy) * rH * 2 - 1) * mip_bound - y) * rdy;
        const float tz = (((nz + 0.5f + 0.5f * sz) * rH * 2 - 1) * mip_bound - z) * rdz;
        tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
        return false;
    }

    // Evaluate lattice point t. Returns true if (x,y,z) is an occupied sample (t is NOT advanced; caller adds dt).
    // Otherwise advances t to the first lattice point at/after the voxel exit and returns false.
    __device__ __forceinline__ bool probe(float& t, float& x, float& y, float& z, float& dt) const {
        float tt;
        if (probe_point(t, x, y, z, dt, tt)) return true;
        t = advance_past(t, tt);
        return false;
    }

    // first lattice point at or behind tt: the reference's `do { t += clamp(t * dt_gamma, dt_min, dt_max); } while (t < tt)`
    // (raymarching.cu:399-402). With dt_gamma == 0 (every Blender config) the step is the same constant in every
    // iteration — clamp(+0, dt_min, dt_max) — so the loop is one dependent FADD per lattice point instead of
    // FMUL + FMNMX + FMNMX + FADD; the sequence of sums, hence the result, is bit-identical.
    __device__ __forceinline__ float advance_past(float t, float tt) const {
        if (dt_gamma == 0.f) {
            const float s = clampf(0.f, dt_min, dt_max);
            do { t += s; } while (t < tt);
        } else {
            do { t += step_size(t); } while (t < tt);
        }
        return t;
    }
};


// ------------------------------------------------------------------------------------------------
// Lattice window: lane i receives t_i = t_start advanced i times by the (serial, fp32) step, i < nvalid (>= 1).
// Constant step (dt_gamma == 0, every Blender config): inside one binade fl(t + dt) = t + dq ulps with a constant
// integer dq (dt rounded to the binade's ulp; round-half-even ties and binade crossings excluded), so the lattice
// points are consecutive-integer-spaced BIT PATTERNS and lane i gets its point in O(1). The window is accepted
// only if every lane verifies t_i == fl(t_{i-1} + dt) with a real fp32 add, i.e. it is the reference's serial
// sequence by construction; otherwise (tie, crossing at lane 1, dt_gamma > 0) the serial prefix is used.
// All 32 lanes must call it with the same t_start.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lattice_window(const Marcher& m, float t_start, uint32_t lane, float& t, uint32_t& nvalid) {
    const uint32_t dbits = __float_as_uint(m.dt_min);
    t = t_start;
    nvalid = 32;
    bool fast = false;
    if (m.dt_gamma == 0.f) {
        const uint32_t tb = __float_as_uint(t_start);
        const int shift = (int)((tb >> 23) & 0xffu) - (int)((dbits >> 23) & 0xffu);
        if (shift >= 1 && shift <= 23 && (tb >> 23) != 0u && (tb >> 23) < 0xffu) {
            const uint32_t md = (dbits & 0x7fffffu) | 0x800000u;
            const uint32_t rem = md & ((1u << shift) - 1u), half = 1u << (shift - 1);
            const uint32_t dq = (md >> shift) + (rem > half ? 1u : 0u);
            const uint32_t bi = tb + lane * dq;
            const bool wv = bi < ((tb | 0x7fffffu) + 1u);            // still inside t_start's binade
            const float ti = __uint_as_float(bi);
            const float pa = __shfl_up_sync(0xffffffffu, ti + m.dt_min, 1);
            const bool ok = (lane == 0) || !wv || (pa == ti);
            const uint32_t wmask = __ballot_sync(0xffffffffu, wv);
            if (rem != half && dq != 0u && __all_sync(0xffffffffu, ok)) {
                fast = true;
                nvalid = __popc(wmask);                               // >= 1 (lane 0 is always inside)
                t = wv ? ti : t_start;
            }
        }
    }
    if (!fast) {
#pragma unroll 1
        for (uint32_t i = 0; i < 31; i++) {
            const float tn = t + m.step_size(t);
            if (i < lane) t = tn;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Warp-cooperative walk of ONE ray (all 32 lanes call it with the same ray). The reference walks the lattice
// t_{k+1} = t_k + clamp(t_k * dt_gamma, dt_min, dt_max) serially, probing a point, then either emitting a sample
// or skipping to the first lattice point past the empty voxel. Here a window of 32 consecutive lattice points is
// generated (serial fp32 adds, so the values are the reference's bit for bit), all 32 are classified in parallel,
// and the serial walk over the window is replayed with ballots/shuffles (~10 instructions per visited point).
// Visited-and-occupied points are exactly the reference's samples, in order.
//   WRITE = false: returns the sample count (<= budget)
//   WRITE = true : additionally writes xyz / dir / (dt, real delta) of sample k to row (k) of the output pointers
// ------------------------------------------------------------------------------------------------
//   t_list (optional): receives the ray parameter t of sample k at t_list[k] (the list-driven writer needs nothing else)
template <bool WRITE>
__device__ __forceinline__ uint32_t warp_walk(const Marcher& m, float t0, float far, uint32_t budget, uint32_t lane,
                                              float* __restrict__ xyzs, float* __restrict__ dirs,
                                              float* __restrict__ deltas, float* __restrict__ t_list = nullptr) {
    uint32_t count = 0;
    float t_start = t0;
    float last_t = t0;  // end of the previous sample (start of the real-delta interval)
    const uint32_t lt_mask = (1u << lane) - 1u;
    while (t_start < far && count < budget) {
        float t;
        uint32_t nvalid;
        lattice_window(m, t_start, lane, t, nvalid);
        const bool wv = lane < nvalid;
        const float t_after = t + m.step_size(t);                 // lattice point following this lane's
        const bool inside = wv && t < far;
        float x = 0.f, y = 0.f, z = 0.f, dt = 0.f, tt = 0.f;
        bool occ = false;
        if (inside) occ = m.probe_point(t, x, y, z, dt, tt);
        const uint32_t in_mask = __ballot_sync(0xffffffffu, inside);
        const uint32_t occ_mask = __ballot_sync(0xffffffffu, occ);
        // replay the serial walk over the window (warp-uniform)
        uint32_t take = 0, cur = 0;
        bool finished = false, jumped = false;
        float t_jump = 0.f;
        while (cur < nvalid) {
            if (!((in_mask >> cur) & 1u)) { finished = true; break; }            // t >= far
            if ((occ_mask >> cur) & 1u) {
                // a run of consecutive occupied points is taken in one step (the serial walk would visit them one by
                // one: same points, same order), cut at the sample budget (max_steps / num_steps)
                const uint32_t room = budget - (count + __popc(take));
                if (room == 0u) { finished = true; break; }
                const uint32_t inv = ~(occ_mask >> cur);
                const uint32_t run = min(inv ? (uint32_t)(__ffs(inv) - 1) : 32u, room);      // >= 1; occupied => inside the window
                take |= (run >= 32u ? 0xffffffffu : ((1u << run) - 1u)) << cur;
                cur += run;
            } else {
                const float tt_cur = __shfl_sync(0xffffffffu, tt, cur);
                const uint32_t above = (cur >= 31) ? 0u : (0xffffffffu << (cur + 1));
                const uint32_t ge = __ballot_sync(0xffffffffu, wv && t >= tt_cur) & above;
                if (ge) {
                    cur = __ffs(ge) - 1;
                } else {   // the empty voxel extends past the window: keep stepping from the last lattice point
                    t_jump = m.advance_past(__shfl_sync(0xffffffffu, t, nvalid - 1), tt_cur);
                    jumped = true;
                    break;
                }
            }
        }
        if (WRITE && take) {
            // real delta of a sample = (t_i + dt_i) - end of the previous sample
            const uint32_t below = take & lt_mask;
            const int prev = below ? (31 - __clz(below)) : -1;
            const float prev_end = __shfl_sync(0xffffffffu, t_after, prev < 0 ? 0 : prev);
            if ((take >> lane) & 1u) {
                const uint32_t k = count + __popc(below);
                const float lt = prev < 0 ? last_t : prev_end;
                xyzs[(size_t)k * 3 + 0] = x; xyzs[(size_t)k * 3 + 1] = y; xyzs[(size_t)k * 3 + 2] = z;
                dirs[(size_t)k * 3 + 0] = m.dx; dirs[(size_t)k * 3 + 1] = m.dy; dirs[(size_t)k * 3 + 2] = m.dz;
                reinterpret_cast<float2*>(deltas)[k] = make_float2(dt, t_after - lt);
            }
        }
        if (!WRITE && t_list && ((take >> lane) & 1u)) t_list[count + __popc(take & lt_mask)] = t;
        if (take) {
            const int top = 31 - __clz(take);
            last_t = __shfl_sync(0xffffffffu, t_after, top);
            count += __popc(take);
        }
        if (finished) break;
        t_start = jumped ? t_jump : __shfl_sync(0xffffffffu, t_after, nvalid - 1);
    }
    return count;
}

}  // namespace pnerf
