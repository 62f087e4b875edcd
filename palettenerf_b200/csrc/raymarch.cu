// raymarch.cu — occupancy-grid ray marching for B200 (sm_100a).
//
// Replaces raymarching/src/raymarching.cu:95-493,848-1021 of the reference (near/far, sphere coords, Morton,
// packbits, training march, inference march, spread). The float arithmetic of one marching step mirrors the
// reference expression by expression so that sample positions, counts and offsets are
// bit-identical; the *schedule* is new:
//   - training march = count kernel -> single-CTA exclusive scan in ray order -> write kernel, instead of two
//     global atomics per ray; slot order is deterministic (ray order), which is one of the orders the
//     reference's atomic race may produce;
//   - both marching passes are warp-per-ray (march_common.cuh::warp_walk): 32 lattice points are classified per
//     step and the lanes of a warp write consecutive samples as coalesced runs;
//   - packbits reads 2x float4 per byte, morton/near-far read and write through coalesced vector accesses.
#include "common.cuh"
#include "march_common.cuh"
#include "ray_common.cuh"

namespace pnerf {

// ------------------------------------------------------------------------------------------------
// near / far  (ref: raymarching.cu:95-148)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_near_far(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                  const float* __restrict__ aabb, uint32_t N, float min_near,
                                                  float* __restrict__ nears, float* __restrict__ fars) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3 + 0], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3 + 0], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    float near, far;
    slab_near_far(ox, oy, oz, dx, dy, dz, aabb, min_near, near, far);
    nears[n] = near;
    fars[n] = far;
}

// ------------------------------------------------------------------------------------------------
// background sphere coordinates (ref: raymarching.cu:166-201)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sph_from_ray(const float* __restrict__ rays_o,
                                                      const float* __restrict__ rays_d, float radius, uint32_t N,
                                                      float* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3 + 0], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3 + 0], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float Bh = ox * dx + oy * dy + oz * dz;  // half of the linear coefficient
    const float Cc = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-Bh + sqrtf(Bh * Bh - A * Cc)) / A;  // far intersection
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    const float rpi = 0.3183098861837907f;
    reinterpret_cast<float2*>(coords)[n] = make_float2(2 * theta * rpi - 1, phi * rpi);
}

// ------------------------------------------------------------------------------------------------
// Morton (ref: raymarching.cu:217-257)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_morton3D(const int32_t* __restrict__ coords, uint32_t N,
                                                  int32_t* __restrict__ indices) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    indices[n] = (int32_t)morton_encode((uint32_t)coords[n * 3 + 0], (uint32_t)coords[n * 3 + 1],
                                        (uint32_t)coords[n * 3 + 2]);
}

__global__ void __launch_bounds__(256) k_morton3D_invert(const int32_t* __restrict__ indices, uint32_t N,
                                                         int32_t* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int32_t ind = indices[n];  // arithmetic shifts of the signed value, like the reference
    coords[n * 3 + 0] = (int32_t)compact3((uint32_t)(ind >> 0));
    coords[n * 3 + 1] = (int32_t)compact3((uint32_t)(ind >> 1));
    coords[n * 3 + 2] = (int32_t)compact3((uint32_t)(ind >> 2));
}

// ------------------------------------------------------------------------------------------------
// packbits (ref: raymarching.cu:271-292). One WARP packs 128 bytes = 1024 cells: in each of 8 steps the lanes read 32
// consecutive float4 (512 contiguous bytes per load instruction), turn them into 4-bit nibbles, and three xor-shuffles
// OR the nibbles of 8 neighbouring lanes into a 32-bit word; every lane keeps one of the 32 words and the warp stores
// them into one 128-byte line. Ragged tails and unaligned pointers take the byte-per-thread path.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_packbits(const float* __restrict__ grid, uint32_t N, float thresh,
                                                  uint8_t* __restrict__ bitfield) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    const uint32_t n0 = warp * 128;                         // first byte of this warp's chunk
    if (n0 >= N) return;
    const bool fast = n0 + 128 <= N && ((reinterpret_cast<uintptr_t>(grid) & 15u) == 0) &&
                      ((reinterpret_cast<uintptr_t>(bitfield) & 3u) == 0);
    if (fast) {
        const float4* g4 = reinterpret_cast<const float4*>(grid) + (size_t)warp * 256;
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = ld_stream4(g4 + i * 32 + lane);
        uint32_t mine = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t w = ((v[i].x > thresh ? 1u : 0u) | (v[i].y > thresh ? 2u : 0u) | (v[i].z > thresh ? 4u : 0u) |
                          (v[i].w > thresh ? 8u : 0u)) << ((lane & 7u) * 4);
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            w |= __shfl_xor_sync(0xffffffffu, w, 4);        // word (4 i + lane / 8) of the chunk, in all 8 lanes of the group
            if ((lane & 7u) == (uint32_t)i) mine = w;
        }
        reinterpret_cast<uint32_t*>(bitfield)[(size_t)warp * 32 + (lane & 7u) * 4 + (lane >> 3)] = mine;
    } else {
        for (uint32_t n = n0 + lane; n < N && n < n0 + 128; n += 32) {
            uint8_t bits = 0;
            for (int i = 0; i < 8; i++) bits |= (grid[(size_t)n * 8 + i] > thresh) ? (uint8_t)(1u << i) : 0;
            bitfield[n] = bits;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// training march (ref: raymarching.cu:315-483)
// ------------------------------------------------------------------------------------------------

// pass 1: count occupied steps per ray; rays[n] = (n, <offset later>, count). One WARP per ray (march_common.cuh:
// warp_walk): 32 lattice points are classified per step instead of one, which matters because a training batch is
// only 4096 rays — thread-per-ray leaves the chip latency-bound at 32 resident warps.
__global__ void __launch_bounds__(256) k_march_train_count(const float* __restrict__ rays_o,
                                                           const float* __restrict__ rays_d,
                                                           const uint8_t* __restrict__ grid, float bound,
                                                           float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C,
                                                           uint32_t H, const float* __restrict__ nears,
                                                           const float* __restrict__ fars,
                                                           const float* __restrict__ noises,
                                                           int32_t* __restrict__ rays,
                                                           float* __restrict__ t_list = nullptr,
                                                           const float* __restrict__ occ = nullptr) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (n >= N) return;
    Marcher m;
    m.init(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, bound, dt_gamma, max_steps, C, H, grid);
    const float t = m.first_t(nears[n], noises[n]);
    float far = fars[n];
    if (occ) far = fminf(far, m.occupied_exit(occ));     // -1 for a ray that misses every occupied cell: no walk at all
    const uint32_t num_steps = warp_walk<false>(m, t, far, max_steps, lane, nullptr, nullptr, nullptr,
                                                t_list ? t_list + (size_t)n * max_steps : nullptr);
    if (lane == 0) {
        rays[n * 3 + 0] = (int32_t)n;
        rays[n * 3 + 2] = (int32_t)num_steps;
    }
}

// pass 2: single-CTA exclusive scan of the counts in ray order. 1024 threads, each owns a contiguous run.
// valid (optional): number of leading rows that will be written = min(total, offset of the first ray that does not fit M)
__global__ void __launch_bounds__(1024) k_march_train_scan(int32_t* __restrict__ rays, uint32_t N,
                                                           int32_t* __restrict__ counter, uint32_t M = 0,
                                                           int32_t* __restrict__ valid = nullptr) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t base_s, first_bad_s, total_s;
    if (threadIdx.x == 0) first_bad_s = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t per = ceil_div(N, 1024u);
    const uint32_t lo = min(N, tid * per), hi = min(N, lo + per);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += (uint32_t)rays[i * 3 + 2];
    // inclusive warp scan
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += v;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = warp_tot[lane];
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (uint32_t)o) winc += v;
        }
        warp_tot[lane] = winc - w;  // exclusive
        if (lane == 31) {
            // counter[0] accumulates the grand total, counter[1] the ray count (ref: raymarching.cu:408-409)
            base_s = (uint32_t)counter[0];
            total_s = base_s + winc;
            counter[0] = (int32_t)(base_s + winc);
            counter[1] += (int32_t)N;
        }
    }
    __syncthreads();
    uint32_t off = base_s + warp_tot[wid] + (inc - sum);
    uint32_t bad = 0xffffffffu;
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t c = (uint32_t)rays[i * 3 + 2];
        rays[i * 3 + 1] = (int32_t)off;
        if (c && off + c > M) bad = min(bad, off);          // this ray (and every later one) writes nothing
        off += c;
    }
    if (valid) {
        if (bad != 0xffffffffu) atomicMin(&first_bad_s, bad);
        __syncthreads();
        if (tid == 0) *valid = (int32_t)min(min(first_bad_s, total_s), M);
    }
}

// pass 3: re-march and write, one warp per ray: the lanes of a warp write consecutive samples, so the three
// output arrays receive contiguous 384 B / 384 B / 256 B runs per store instruction.
__global__ void __launch_bounds__(256) k_march_train_write(const float* __restrict__ rays_o,
                                                           const float* __restrict__ rays_d,
                                                           const uint8_t* __restrict__ grid, float bound,
                                                           float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C,
                                                           uint32_t H, uint32_t M, const float* __restrict__ nears,
                                                           const float* __restrict__ fars,
                                                           const float* __restrict__ noises,
                                                           const int32_t* __restrict__ rays, float* __restrict__ xyzs,
                                                           float* __restrict__ dirs, float* __restrict__ deltas) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t num_steps = (uint32_t)rays[n * 3 + 2], offset = (uint32_t)rays[n * 3 + 1];
    if (num_steps == 0 || offset + num_steps > M) return;   // ref: raymarching.cu:418-419
    Marcher m;
    m.init(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, bound, dt_gamma, max_steps, C, H, grid);
    const float t = m.first_t(nears[n], noises[n]);
    warp_walk<true>(m, t, fars[n], num_steps, lane, xyzs + (size_t)offset * 3, dirs + (size_t)offset * 3,
                    deltas + (size_t)offset * 2);
}

// pass 3, list-driven: pass 1 left the ray parameter of every sample in t_list[n][0..count); the samples are produced
// from that list without touching the occupancy grid again. Lane k handles samples k, k+32, ...: position and step from
// Marcher::position (the expressions of the walk), real delta = (t_k + dt_k) - (t_{k-1} + dt_{k-1}) (first sample:
// - first lattice point), exactly what warp_walk<true> writes.
__global__ void __launch_bounds__(256) k_march_train_emit(const float* __restrict__ rays_o,
                                                          const float* __restrict__ rays_d, float bound, float dt_gamma,
                                                          uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                                          uint32_t M, const float* __restrict__ nears,
                                                          const float* __restrict__ noises,
                                                          const int32_t* __restrict__ rays,
                                                          const float* __restrict__ t_list, float* __restrict__ xyzs,
                                                          float* __restrict__ dirs, float* __restrict__ deltas) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (n >= N) return;
    const uint32_t num_steps = (uint32_t)rays[n * 3 + 2], offset = (uint32_t)rays[n * 3 + 1];
    if (num_steps == 0 || offset + num_steps > M) return;   // ref: raymarching.cu:418-419
    Marcher m;
    m.init(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, bound, dt_gamma, max_steps, C, H, nullptr);
    const float t0 = m.first_t(nears[n], noises[n]);
    const float* tl = t_list + (size_t)n * max_steps;
    float* px = xyzs + (size_t)offset * 3;
    float* pd = dirs + (size_t)offset * 3;
    float2* pl = reinterpret_cast<float2*>(deltas) + offset;
    for (uint32_t k = lane; k < num_steps; k += 32) {
        const float t = tl[k];
        float prev_end = t0;
        if (k) { const float tp = tl[k - 1]; prev_end = tp + m.step_size(tp); }
        float x, y, z, dt;
        m.position(t, x, y, z, dt);
        const float t_after = t + dt;
        px[(size_t)k * 3 + 0] = x; px[(size_t)k * 3 + 1] = y; px[(size_t)k * 3 + 2] = z;
        pd[(size_t)k * 3 + 0] = m.dx; pd[(size_t)k * 3 + 1] = m.dy; pd[(size_t)k * 3 + 2] = m.dz;
        pl[k] = make_float2(dt, t_after - prev_end);
    }
}

// world-space bounds of the occupied cells of all cascades, padded by one cell. Two launches: kOccBlocks CTAs scan the
// bitfield (512 KiB at C = 2, H = 128) and leave per-CTA bounds in the scratch part of occ; one CTA reduces them.
// One byte of the bitfield = 8 cells with consecutive Morton codes = one 2x2x2 block; a block with any bit set counts
// as occupied. A side that reaches the scene bound is opened (+-FLT_MAX): positions are clamped to the bound
// (raymarching.cu:366-368), so "beyond the bound" does not imply "outside the cell". occ[0..6) = (lo xyz, hi xyz);
// an empty grid yields lo > hi (every ray misses). occ[6 .. 6 + 6 kOccBlocks) is scratch.
constexpr int kOccBlocks = kNumSMs;

__global__ void __launch_bounds__(256) k_occupied_bounds_scan(const uint8_t* __restrict__ bitfield, uint32_t C, uint32_t H,
                                                               float bound, float* __restrict__ occ) {
    const float big = 3.402823466e+38f;
    float lo[3] = {big, big, big}, hi[3] = {-big, -big, -big};
    const uint32_t bytes_per_level = H * H * H / 8, total = C * bytes_per_level;
    const uint32_t* words = reinterpret_cast<const uint32_t*>(bitfield);
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < total / 4; w += gridDim.x * blockDim.x) {
        const uint32_t v = __ldg(words + w);
        if (!v) continue;
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) {
            if (!((v >> (8 * j)) & 0xffu)) continue;
            const uint32_t byte = w * 4 + j, level = byte / bytes_per_level;
            const uint32_t code = (byte - level * bytes_per_level) * 8;
            const uint32_t c[3] = {compact3(code), compact3(code >> 1), compact3(code >> 2)};
            const float mb = fminf(scalbnf(1.0f, (int)level), bound);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                lo[k] = fminf(lo[k], (((float)c[k] - 1.0f) / (float)H * 2 - 1) * mb);
                hi[k] = fmaxf(hi[k], (((float)c[k] + 3.0f) / (float)H * 2 - 1) * mb);
            }
        }
    }
    __shared__ float red[6][8];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if (lane == 0) { red[k][wid] = lo[k]; red[3 + k][wid] = hi[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const bool is_lo = threadIdx.x < 3;
        float v = is_lo ? big : -big;
        for (uint32_t i = 0; i < blockDim.x / 32; i++) v = is_lo ? fminf(v, red[threadIdx.x][i]) : fmaxf(v, red[threadIdx.x][i]);
        occ[6 + blockIdx.x * 6 + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(192) k_occupied_bounds_finish(uint32_t C, uint32_t H, float bound, uint32_t blocks,
                                                                 float* __restrict__ occ) {
    const float big = 3.402823466e+38f;
    const uint32_t lane = threadIdx.x & 31u, k = threadIdx.x >> 5;     // warp k reduces component k
    const bool is_lo = k < 3;
    float v = is_lo ? big : -big;
    for (uint32_t b = lane; b < blocks; b += 32) {
        const float p = occ[6 + b * 6 + k];
        v = is_lo ? fminf(v, p) : fmaxf(v, p);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float q = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_lo ? fminf(v, q) : fmaxf(v, q);
    }
    if (lane == 0) {
        const float cell = fminf(scalbnf(1.0f, (int)C - 1), bound) * 2 / (float)H;   // coarsest cell
        if (is_lo && v <= -bound + cell) v = -big;
        if (!is_lo && v >= bound - cell) v = big;
        occ[k] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// mark_untrained_grid (ref: NeRFRenderer.mark_untrained_grid, nerf/renderer.py:395-465): a cell that no
// training camera sees, or that lies closer than min_near to a camera that sees it (optionally: closer than min_near to
// any camera centre), gets density -1 and never sets an occupancy bit. The reference walks a 5-deep Python loop of
// batched matmuls over 64^3 blocks x cascades x 64-camera chunks; here one thread owns one (cascade, cell) and loops over
// the cameras, which are staged in shared memory 64 at a time (12 floats each). Same expressions in fp32:
//   world = (2 c / (H-1) - 1) * (bound_k - bound_k / H);  cam = (world - t) R  (row vector times the c2w rotation)
//   seen  = z > 0 && |x| < cx/fx z + 2 half && |y| < cy/fy z + 2 half;  close = seen && z < min_near
// ------------------------------------------------------------------------------------------------
constexpr int kMarkCams = 64;

__global__ void __launch_bounds__(256) k_mark_untrained(const float* __restrict__ poses, uint32_t B, float cxfx, float cyfy,
                                                        uint32_t C, uint32_t H, float bound, float min_near,
                                                        int filter_close_point, float* __restrict__ density_grid,
                                                        unsigned int* __restrict__ n_marked) {
    __shared__ float cam[kMarkCams][12];
    const uint32_t H3 = H * H * H;
    const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x, cas = blockIdx.y;
    const bool live = cell < H3;
    const float b = fminf(scalbnf(1.0f, (int)cas), bound), half = b / (float)H;
    float wx = 0.f, wy = 0.f, wz = 0.f;
    if (live) {
        const uint32_t x = compact3(cell), y = compact3(cell >> 1), z = compact3(cell >> 2);     // density_grid is in Morton order
        const float s = b - half, hm1 = (float)(H - 1);
        wx = (2 * (float)x / hm1 - 1) * s; wy = (2 * (float)y / hm1 - 1) * s; wz = (2 * (float)z / hm1 - 1) * s;
    }
    uint32_t seen = 0, close = 0;
    for (uint32_t c0 = 0; c0 < B; c0 += kMarkCams) {
        const uint32_t nc = min((uint32_t)kMarkCams, B - c0);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nc * 12; i += blockDim.x) {
            const uint32_t k = i / 12, e = i - k * 12;          // rows 0..2 of the 4x4 pose: R[r][0..2], t[r]
            cam[k][e] = poses[(size_t)(c0 + k) * 16 + (e >> 2) * 4 + (e & 3)];
        }
        __syncthreads();
        if (!live) continue;
        for (uint32_t k = 0; k < nc; k++) {
            const float* P = cam[k];
            const float dx = wx - P[3], dy = wy - P[7], dz = wz - P[11];
            const float cx_ = dx * P[0] + dy * P[4] + dz * P[8];
            const float cy_ = dx * P[1] + dy * P[5] + dz * P[9];
            const float cz_ = dx * P[2] + dy * P[6] + dz * P[10];
            const bool vis = cz_ > 0 && fabsf(cx_) < cxfx * cz_ + half * 2 && fabsf(cy_) < cyfy * cz_ + half * 2;
            seen += vis ? 1u : 0u;
            close += (vis && cz_ < min_near) ? 1u : 0u;
            if (filter_close_point) close += (sqrtf(cx_ * cx_ + cy_ * cy_ + cz_ * cz_) < min_near) ? 1u : 0u;
        }
    }
    const bool mark = live && (seen == 0u || close != 0u);
    if (mark) density_grid[(size_t)cas * H3 + cell] = -1.0f;
    if (n_marked) {
        const uint32_t m = __popc(__ballot_sync(0xffffffffu, mark));
        if ((threadIdx.x & 31u) == 0 && m) atomicAdd(n_marked, m);
    }
}

// ------------------------------------------------------------------------------------------------
// inference march (ref: raymarching.cu:907-1011)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_march_rays(uint32_t n_alive, uint32_t n_step,
                                                    const int32_t* __restrict__ rays_alive,
                                                    const float* __restrict__ rays_t,
                                                    const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                    float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                                                    uint32_t H, const uint8_t* __restrict__ grid,
                                                    const float* __restrict__ nears, const float* __restrict__ fars,
                                                    float* __restrict__ xyzs, float* __restrict__ dirs,
                                                    float* __restrict__ deltas, const float* __restrict__ noises) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    Marcher m;
    m.init(rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, bound, dt_gamma, max_steps, C, H, grid);
    const float far = fars[index];
    float t = rays_t[index];
    t += m.step_size(t) * noises[n];
    float last_t = t;
    uint32_t step = 0;
    float* px = xyzs + (size_t)n * n_step * 3;
    float* pd = dirs + (size_t)n * n_step * 3;
    float* pl = deltas + (size_t)n * n_step * 2;
    float x, y, z, dt;
    while (t < far && step < n_step) {
        if (m.probe(t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = m.dx; pd[1] = m.dy; pd[2] = m.dz;
            t += dt;
            pl[0] = dt;
            pl[1] = t - last_t;
            last_t = t;
            px += 3; pd += 3; pl += 2;
            step++;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// spread per-ray values to the ray's samples (ref: raymarching.cu:848-882); warp per ray, coalesced writes
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_spread(const float* __restrict__ input, const int32_t* __restrict__ rays,
                                                uint32_t M, uint32_t N, uint32_t n_channel,
                                                float* __restrict__ output) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (warp >= N) return;
    const uint32_t index = (uint32_t)rays[warp * 3], offset = (uint32_t)rays[warp * 3 + 1];
    uint32_t num_steps = (uint32_t)rays[warp * 3 + 2];
    if (num_steps == 0 || offset >= M) return;
    num_steps = min(num_steps, M - offset);
    const float* in = input + (size_t)index * n_channel;
    float* out = output + (size_t)offset * n_channel;
    const uint32_t total = num_steps * n_channel;
    for (uint32_t i = lane; i < total; i += 32) out[i] = in[i % n_channel];
}

}  // namespace pnerf

// ================================================================================================
// C ABI
// ================================================================================================
using namespace pnerf;

extern "C" {

int pnerf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                             float* nears, float* fars, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(rays_o && rays_d && aabb && nears && fars);
    k_near_far<<<ceil_div(N, 256u), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    return check_launch("near_far_from_aabb");
}

int pnerf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords,
                       void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(rays_o && rays_d && coords);
    k_sph_from_ray<<<ceil_div(N, 256u), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, radius, N, coords);
    return check_launch("sph_from_ray");
}

int pnerf_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(coords && indices);
    k_morton3D<<<ceil_div(N, 256u), 256, 0, (cudaStream_t)stream>>>(coords, N, indices);
    return check_launch("morton3D");
}

int pnerf_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(coords && indices);
    k_morton3D_invert<<<ceil_div(N, 256u), 256, 0, (cudaStream_t)stream>>>(indices, N, coords);
    return check_launch("morton3D_invert");
}

int pnerf_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(grid && bitfield);
    const uint32_t warps = ceil_div(N, 128u);
    k_packbits<<<ceil_div(warps, 8u), 256, 0, (cudaStream_t)stream>>>(grid, N, density_thresh, bitfield);
    return check_launch("packbits");
}

int pnerf_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                           uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                           const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                           const float* noises, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(rays_o && rays_d && grid && nears && fars && xyzs && dirs && deltas && rays && counter && noises);
    PNERF_REQUIRE(C >= 1 && C <= 16 && H >= 1 && max_steps >= 1);
    if (H > 1024) return PNERF_ERR_UNSUPPORTED;  // 10-bit Morton coordinates
    cudaStream_t s = (cudaStream_t)stream;
    k_march_train_count<<<ceil_div(N, 8u), 256, 0, s>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears,
                                                        fars, noises, rays);
    k_march_train_scan<<<1, 1024, 0, s>>>(rays, N, counter);
    k_march_train_write<<<ceil_div(N, 8u), 256, 0, s>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M,
                                                        nears, fars, noises, rays, xyzs, dirs, deltas);
    return check_launch("march_rays_train");
}

uint32_t pnerf_occupied_bounds_floats(void) { return 6u + 6u * (uint32_t)kOccBlocks; }

int pnerf_occupied_bounds(const uint8_t* bitfield, uint32_t C, uint32_t H, float bound, float* occ_aabb, void* stream) {
    PNERF_REQUIRE(bitfield && occ_aabb && C >= 1 && C <= 16 && H >= 2 && bound > 0.f);
    if (H > 1024 || ((uint64_t)H * H * H) % 32 != 0) return PNERF_ERR_UNSUPPORTED;
    k_occupied_bounds_scan<<<kOccBlocks, 256, 0, (cudaStream_t)stream>>>(bitfield, C, H, bound, occ_aabb);
    k_occupied_bounds_finish<<<1, 192, 0, (cudaStream_t)stream>>>(C, H, bound, kOccBlocks, occ_aabb);
    return check_launch("occupied_bounds");
}

int pnerf_march_rays_train_ws(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                              uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                              const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                              const float* noises, float* t_list, const float* occ_aabb, int32_t* valid_rows, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(rays_o && rays_d && grid && nears && fars && xyzs && dirs && deltas && rays && counter && noises && t_list);
    PNERF_REQUIRE(C >= 1 && C <= 16 && H >= 1 && max_steps >= 1);
    if (H > 1024) return PNERF_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    k_march_train_count<<<ceil_div(N, 8u), 256, 0, s>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears,
                                                        fars, noises, rays, t_list, occ_aabb);
    k_march_train_scan<<<1, 1024, 0, s>>>(rays, N, counter, M, valid_rows);
    k_march_train_emit<<<ceil_div(N, 8u), 256, 0, s>>>(rays_o, rays_d, bound, dt_gamma, max_steps, N, C, H, M, nears, noises,
                                                       rays, t_list, xyzs, dirs, deltas);
    return check_launch("march_rays_train_ws");
}

int pnerf_mark_untrained_grid(const float* poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t C, uint32_t H,
                              float bound, float min_near, int filter_close_point, float* density_grid,
                              uint32_t* n_marked, void* stream) {
    PNERF_REQUIRE(density_grid && (B == 0 || poses) && C >= 1 && C <= 16 && H >= 2 && fx != 0.f && fy != 0.f && bound > 0.f);
    if (H > 1024) return PNERF_ERR_UNSUPPORTED;
    const dim3 grid(ceil_div(H * H * H, 256u), C, 1);
    k_mark_untrained<<<grid, 256, 0, (cudaStream_t)stream>>>(poses, B, cx / fx, cy / fy, C, H, bound, min_near,
                                                            filter_close_point, density_grid, n_marked);
    return check_launch("mark_untrained_grid");
}

int pnerf_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                     const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                     uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                     float* dirs, float* deltas, const float* noises, void* stream) {
    if (n_alive == 0 || n_step == 0) return PNERF_OK;
    PNERF_REQUIRE(rays_alive && rays_t && rays_o && rays_d && grid && nears && fars && xyzs && dirs && deltas && noises);
    PNERF_REQUIRE(C >= 1 && C <= 16 && H >= 1 && max_steps >= 1);
    if (H > 1024) return PNERF_ERR_UNSUPPORTED;
    k_march_rays<<<ceil_div(n_alive, 128u), 128, 0, (cudaStream_t)stream>>>(
        n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars, xyzs,
        dirs, deltas, noises);
    return check_launch("march_rays");
}

int pnerf_spread_ray_to_sample(const float* input, const int32_t* rays, uint32_t M, uint32_t N, uint32_t n_channel,
                               float* output, void* stream) {
    if (N == 0 || n_channel == 0) return PNERF_OK;
    PNERF_REQUIRE(input && rays && output);
    if (n_channel > 128) return PNERF_ERR_UNSUPPORTED;
    k_spread<<<ceil_div(N * 32u, 256u), 256, 0, (cudaStream_t)stream>>>(input, rays, M, N, n_channel, output);
    return check_launch("spread_ray_to_sample");
}

}  // extern "C"
