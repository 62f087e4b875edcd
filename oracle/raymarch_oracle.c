/*
 * raymarch_oracle.c — CPU restatement (plain C, fp32) of the reference's ray-marching / compositing algorithms.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the
 * checker. The product path (palettenerf_b200/) never links or calls this file.
 *
 * PARITY STATUS: the reference holds no golden vectors / known-answer tests for this path, so this
 * restatement is pinned against the reference ITSELF: tests/test_golden.py checks it against outputs of the reference's
 * own CUDA kernels (tests/golden/ref_kernels.npz, written on a B200 by tests/golden/make_golden_gpu.py from oracle/_ref,
 * which oracle/build_ref.py compiles from /root/reference), and tests/test_raymarching_gpu.py / test_composite_gpu.py run
 * those kernels side by side with this oracle and the new kernels on the GPU box.
 *
 * Every function cites the reference lines it follows (paths under the reference repo). fp32 arithmetic is written
 * with explicit fmaf() wherever nvcc's default -fmad=true contracts a*b+c in the reference kernels, and this file is
 * compiled with -ffp-contract=off so the host compiler adds none of its own.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/liboracle.so oracle/raymarch_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

/* raymarching.cu:59-74 */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
/* raymarching.cu:76-84 */
static inline uint32_t morton3D_invert(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

void oracle_morton3D(const int32_t* coords, uint32_t N, int32_t* indices) { /* raymarching.cu:217-229 */
    for (uint32_t n = 0; n < N; n++)
        indices[n] = (int32_t)morton3D((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}

void oracle_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords) { /* raymarching.cu:240-257 */
    for (uint32_t n = 0; n < N; n++) {
        const int32_t ind = indices[n];
        coords[n * 3 + 0] = (int32_t)morton3D_invert((uint32_t)(ind >> 0));
        coords[n * 3 + 1] = (int32_t)morton3D_invert((uint32_t)(ind >> 1));
        coords[n * 3 + 2] = (int32_t)morton3D_invert((uint32_t)(ind >> 2));
    }
}

void oracle_packbits(const float* grid, uint32_t N, float thresh, uint8_t* bitfield) { /* raymarching.cu:271-292 */
    for (uint32_t n = 0; n < N; n++) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; i++) bits |= (grid[(size_t)n * 8 + i] > thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* raymarching.cu:95-148 */
void oracle_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                               float* nears, float* fars) {
    const float FMAX = 3.402823466e+38f;
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float rdx = 1 / rays_d[n * 3], rdy = 1 / rays_d[n * 3 + 1], rdz = 1 / rays_d[n * 3 + 2];
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, s;
        if (near > far) { s = near; near = far; far = s; }
        float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
        if (near_y > far_y) { s = near_y; near_y = far_y; far_y = s; }
        if (near > far_y || near_y > far) { nears[n] = fars[n] = FMAX; continue; }
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
        if (near_z > far_z) { s = near_z; near_z = far_z; far_z = s; }
        if (near > far_z || near_z > far) { nears[n] = fars[n] = FMAX; continue; }
        if (near_z > near) near = near_z;
        if (far_z < far) far = far_z;
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* raymarching.cu:45-57: frexpf exponent, clamped to [0, C-1] */
static inline int mip_from_value(float mx, float C) {
    int e;
    frexpf(mx, &e);
    return (int)fminf(C - 1, fmaxf(0, (float)e));
}

typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, bound, dt_gamma, dt_min, dt_max, rH;
    uint32_t C, H;
    const uint8_t* grid;
} ray_ctx;

static void ray_init(ray_ctx* r, const float* o, const float* d, float bound, float dt_gamma, uint32_t max_steps,
                     uint32_t C, uint32_t H, const uint8_t* grid) {
    r->ox = o[0]; r->oy = o[1]; r->oz = o[2];
    r->dx = d[0]; r->dy = d[1]; r->dz = d[2];
    r->rdx = 1 / d[0]; r->rdy = 1 / d[1]; r->rdz = 1 / d[2];
    r->bound = bound; r->dt_gamma = dt_gamma;
    r->dt_min = 2 * 1.7320508075688772f / (float)max_steps;                    /* raymarching.cu:348 */
    r->dt_max = 2 * 1.7320508075688772f * (float)(1u << (C - 1)) / (float)H;   /* raymarching.cu:349 */
    r->rH = 1 / (float)H;
    r->C = C; r->H = H; r->grid = grid;
}

/* One lattice point (raymarching.cu:364-403). Returns 1 if occupied (x,y,z,dt filled; t untouched),
 * else advances *t past the empty voxel and returns 0. */
static int ray_probe(const ray_ctx* r, float* t, float* x, float* y, float* z, float* dt) {
    const float H = (float)r->H;
    *x = clampf(fmaf(*t, r->dx, r->ox), -r->bound, r->bound);
    *y = clampf(fmaf(*t, r->dy, r->oy), -r->bound, r->bound);
    *z = clampf(fmaf(*t, r->dz, r->oz), -r->bound, r->bound);
    *dt = clampf(*t * r->dt_gamma, r->dt_min, r->dt_max);

    const float mx = fmaxf(fabsf(*x), fmaxf(fabsf(*y), fabsf(*z)));
    const int lp = mip_from_value(mx, (float)r->C);
    const int ld = mip_from_value((float)((double)(*dt * H) * 0.5), (float)r->C);
    const int level = lp > ld ? lp : ld;

    const float mip_bound = fminf(scalbnf(1.0f, level), r->bound);
    const float mip_rbound = 1 / mip_bound;

    /* 0.5 * (x * mip_rbound + 1) * H  evaluated in double, then narrowed by clamp()'s float parameter */
    const int nx = (int)clampf((float)(0.5 * (double)fmaf(*x, mip_rbound, 1.0f) * (double)r->H), 0.0f, (float)(r->H - 1));
    const int ny = (int)clampf((float)(0.5 * (double)fmaf(*y, mip_rbound, 1.0f) * (double)r->H), 0.0f, (float)(r->H - 1));
    const int nz = (int)clampf((float)(0.5 * (double)fmaf(*z, mip_rbound, 1.0f) * (double)r->H), 0.0f, (float)(r->H - 1));

    const float H3 = (float)(r->H * r->H * r->H);
    const uint32_t index = (uint32_t)fmaf((float)level, H3, (float)morton3D((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    const int occ = r->grid[index / 8] & (1 << (index % 8));
    if (occ) return 1;

    const float sx = copysignf(1.0f, r->dx), sy = copysignf(1.0f, r->dy), sz = copysignf(1.0f, r->dz);
    const float tx = fmaf(fmaf(fmaf(0.5f, sx, (float)nx + 0.5f) * r->rH, 2.0f, -1.0f), mip_bound, -*x) * r->rdx;
    const float ty = fmaf(fmaf(fmaf(0.5f, sy, (float)ny + 0.5f) * r->rH, 2.0f, -1.0f), mip_bound, -*y) * r->rdy;
    const float tz = fmaf(fmaf(fmaf(0.5f, sz, (float)nz + 0.5f) * r->rH, 2.0f, -1.0f), mip_bound, -*z) * r->rdz;
    const float tt = *t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    do {
        *t += clampf(*t * r->dt_gamma, r->dt_min, r->dt_max);
    } while (*t < tt);
    return 0;
}

/* raymarching.cu:315-483 with the slot race resolved in ray order (offset = running sum, row n = ray n). */
void oracle_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                             uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                             const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                             const float* noises) {
    uint32_t running = (uint32_t)counter[0];
    for (uint32_t n = 0; n < N; n++) {
        ray_ctx r;
        ray_init(&r, rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, bound, dt_gamma, max_steps, C, H, grid);
        const float far = fars[n];
        float t0 = nears[n];
        t0 = fmaf(clampf(t0 * dt_gamma, r.dt_min, r.dt_max), noises[n], t0);
        float t = t0, x, y, z, dt;
        uint32_t num_steps = 0;
        while (t < far && num_steps < max_steps) {
            if (ray_probe(&r, &t, &x, &y, &z, &dt)) { num_steps++; t += dt; }
        }
        const uint32_t point_index = running;
        running += num_steps;
        rays[n * 3] = (int32_t)n;
        rays[n * 3 + 1] = (int32_t)point_index;
        rays[n * 3 + 2] = (int32_t)num_steps;
        if (num_steps == 0) continue;
        if (point_index + num_steps > M) continue;
        float* px = xyzs + (size_t)point_index * 3;
        float* pd = dirs + (size_t)point_index * 3;
        float* pl = deltas + (size_t)point_index * 2;
        t = t0;
        float last_t = t;
        uint32_t step = 0;
        while (t < far && step < num_steps) {
            if (ray_probe(&r, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                t += dt;
                pl[0] = dt;
                pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2;
                step++;
            }
        }
    }
    counter[0] = (int32_t)running;
    counter[1] += (int32_t)N;
}

/* raymarching.cu:907-1011 */
void oracle_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                       const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                       uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                       float* dirs, float* deltas, const float* noises) {
    (void)nears;
    for (uint32_t n = 0; n < n_alive; n++) {
        const int index = rays_alive[n];
        ray_ctx r;
        ray_init(&r, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, bound, dt_gamma, max_steps, C, H, grid);
        const float far = fars[index];
        float t = rays_t[index];
        t = fmaf(clampf(t * dt_gamma, r.dt_min, r.dt_max), noises[n], t);
        float last_t = t, x, y, z, dt;
        uint32_t step = 0;
        float* px = xyzs + (size_t)n * n_step * 3;
        float* pd = dirs + (size_t)n * n_step * 3;
        float* pl = deltas + (size_t)n * n_step * 2;
        while (t < far && step < n_step) {
            if (ray_probe(&r, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                t += dt;
                pl[0] = dt;
                pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2;
                step++;
            }
        }
    }
}

/* raymarching.cu:504-580 (expf stands in for __expf; tolerance in the tests) */
void oracle_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                         uint32_t M, uint32_t N, float T_thresh, float* weights_sum, float* depth,
                                         float* image) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        if (!(num_steps == 0 || offset + num_steps > M)) {
            for (uint32_t step = 0; step < num_steps; step++) {
                const size_t s = (size_t)offset + step;
                const float alpha = 1.0f - expf(-sigmas[s] * deltas[s * 2]);
                const float w = alpha * T;
                r += w * rgbs[s * 3]; g += w * rgbs[s * 3 + 1]; b += w * rgbs[s * 3 + 2];
                t += deltas[s * 2 + 1];
                d += w * t;
                ws += w;
                T *= 1.0f - alpha;
                if (T < T_thresh) break;
            }
        }
        weights_sum[index] = ws; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* raymarching.cu:681-761 */
void oracle_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                          const float* rgbs, const float* deltas, const int32_t* rays,
                                          const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                          float T_thresh, float* grad_sigmas, float* grad_rgbs) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps > M) continue;
        const float* gi = grad_image + (size_t)index * 3;
        const float gws = grad_weights_sum[index];
        const float rf = image[index * 3], gf = image[index * 3 + 1], bf = image[index * 3 + 2], wsf = weights_sum[index];
        float T = 1.0f, r = 0, g = 0, b = 0;
        for (uint32_t step = 0; step < num_steps; step++) {
            const size_t s = (size_t)offset + step;
            const float alpha = 1.0f - expf(-sigmas[s] * deltas[s * 2]);
            const float w = alpha * T;
            r += w * rgbs[s * 3]; g += w * rgbs[s * 3 + 1]; b += w * rgbs[s * 3 + 2];
            T *= 1.0f - alpha;
            grad_rgbs[s * 3] = gi[0] * w; grad_rgbs[s * 3 + 1] = gi[1] * w; grad_rgbs[s * 3 + 2] = gi[2] * w;
            grad_sigmas[s] = deltas[s * 2] * (gi[0] * (T * rgbs[s * 3] - (rf - r)) + gi[1] * (T * rgbs[s * 3 + 1] - (gf - g)) +
                                              gi[2] * (T * rgbs[s * 3 + 2] - (bf - b)) + gws * (1 - wsf));
            if (T < T_thresh) break;
        }
    }
}

/* raymarching.cu:583-645 (note the >= M drop test at :601) */
void oracle_composite_rays_flex_train_forward(const float* sigmas, const float* input, const float* deltas,
                                              const int32_t* rays, uint32_t M, uint32_t N, uint32_t nc, float T_thresh,
                                              float* output) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        float temp[128];
        memset(temp, 0, sizeof(temp));
        if (!(num_steps == 0 || offset + num_steps >= M)) {
            float T = 1.0f;
            for (uint32_t step = 0; step < num_steps; step++) {
                const size_t s = (size_t)offset + step;
                const float alpha = 1.0f - expf(-sigmas[s] * deltas[s * 2]);
                const float w = alpha * T;
                for (uint32_t i = 0; i < nc; i++) temp[i] += w * input[s * nc + i];
                T *= 1.0f - alpha;
                if (T < T_thresh) break;
            }
        }
        for (uint32_t i = 0; i < nc; i++) output[(size_t)index * nc + i] = temp[i];
    }
}

/* raymarching.cu:764-819 (breaks before writing the terminating sample's gradient, :806) */
void oracle_composite_rays_flex_train_backward(const float* grad_output, const float* sigmas, const float* deltas,
                                               const int32_t* rays, uint32_t M, uint32_t N, uint32_t nc, float T_thresh,
                                               float* grad_input) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps >= M) continue;
        float T = 1.0f;
        for (uint32_t step = 0; step < num_steps; step++) {
            const size_t s = (size_t)offset + step;
            const float alpha = 1.0f - expf(-sigmas[s] * deltas[s * 2]);
            const float w = alpha * T;
            T *= 1.0f - alpha;
            if (T < T_thresh) break;
            for (uint32_t i = 0; i < nc; i++) grad_input[s * nc + i] = grad_output[(size_t)index * nc + i] * w;
        }
    }
}

/* raymarching.cu:848-882 */
void oracle_spread_ray_to_sample(const float* input, const int32_t* rays, uint32_t M, uint32_t N, uint32_t nc,
                                 float* output) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        for (uint32_t step = 0; step < num_steps && offset + step < M; step++)
            for (uint32_t i = 0; i < nc; i++) output[((size_t)offset + step) * nc + i] = input[(size_t)index * nc + i];
    }
}

/* raymarching.cu:1025-1111 */
void oracle_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                           const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum, float* depth,
                           float* image) {
    for (uint32_t n = 0; n < n_alive; n++) {
        const int index = rays_alive[n];
        float t = rays_t[index], ws = weights_sum[index], d = depth[index];
        float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
        uint32_t step = 0;
        while (step < n_step) {
            const size_t s = (size_t)n * n_step + step;
            if (deltas[s * 2] == 0) break;
            const float alpha = 1.0f - expf(-sigmas[s] * deltas[s * 2]);
            const float T = 1 - ws;
            const float w = alpha * T;
            ws += w;
            t += deltas[s * 2 + 1];
            d += w * t;
            r += w * rgbs[s * 3]; g += w * rgbs[s * 3 + 1]; b += w * rgbs[s * 3 + 2];
            if (T < T_thresh) break;
            step++;
        }
        if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
        weights_sum[index] = ws; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* raymarching.cu:1114-1185 */
void oracle_composite_rays_flex(uint32_t n_alive, uint32_t n_step, uint32_t nc, float T_thresh, const int32_t* rays_alive,
                                const float* sigmas, const float* input, const float* deltas, const float* weights_sum,
                                float* output) {
    for (uint32_t n = 0; n < n_alive; n++) {
        const int index = rays_alive[n];
        float ws = weights_sum[index];
        for (uint32_t step = 0; step < n_step; step++) {
            const size_t s = (size_t)n * n_step + step;
            if (deltas[s * 2] == 0) break;
            const float alpha = 1.0f - expf(-sigmas[s] * deltas[s * 2]);
            const float T = 1 - ws;
            const float w = alpha * T;
            ws += w;
            for (uint32_t i = 0; i < nc; i++) output[(size_t)index * nc + i] += w * input[s * nc + i];
            if (T < T_thresh) break;
        }
    }
}

/* palette/src/palette.cu:58-85 */
void oracle_rgb_to_hsv(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; i++) {
        const float r = in[i * 3], g = in[i * 3 + 1], b = in[i * 3 + 2];
        const float cmax = fmaxf(fmaxf(r, g), b), cmin = fminf(fminf(r, g), b), diff = cmax - cmin;
        float h, s;
        if (fabsf(diff) < 1e-9f) h = 0;
        else if (fabsf(cmax - r) < 1e-9f) h = fmodf(60 * ((g - b) / diff) + 360, 360);
        else if (fabsf(cmax - g) < 1e-9f) h = fmodf(60 * ((b - r) / diff) + 120, 360);
        else h = fmodf(60 * ((r - g) / diff) + 240, 360);
        s = (fabsf(cmax) < 1e-9f) ? 0 : (diff / cmax) * 100;
        out[i * 3] = h; out[i * 3 + 1] = s; out[i * 3 + 2] = cmax * 100;
    }
}

/* palette/src/palette.cu:101-132 */
void oracle_hsv_to_rgb(uint32_t n, const float* in, float* out) {
    for (uint32_t i = 0; i < n; i++) {
        const float h = in[i * 3], s = in[i * 3 + 1], v = in[i * 3 + 2];
        const float c = s / 100 * v / 100;
        const float x = c * (1 - fabsf(fmodf(h / 60, 2) - 1));
        const float m = v / 100 - c;
        float r = 0, g = 0, b = 0;
        if (h >= 0 && h < 60) { r = c; g = x; }
        else if (h >= 60 && h < 120) { r = x; g = c; }
        else if (h >= 120 && h < 180) { g = c; b = x; }
        else if (h >= 180 && h < 240) { g = x; b = c; }
        else if (h >= 240 && h < 300) { r = x; b = c; }
        else { r = c; b = x; }
        out[i * 3] = r + m; out[i * 3 + 1] = g + m; out[i * 3 + 2] = b + m;
    }
}

/* palette/src/bindings.cpp:40-91 */
void oracle_rgb_histogram(const float* rgb, const float* wgt, uint64_t n, int bpc, double* bin_w, float* bin_c) {
    const uint32_t num_bins = 1u << (bpc * 3);
    for (uint32_t i = 0; i < num_bins; i++) bin_w[i] = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint32_t index = 0;
        for (int k = 0; k < 3; k++) {
            float c = fmaxf(0.0f, fminf(0.999f, rgb[i * 3 + k]));
            index <<= bpc;
            index += (uint32_t)(c * (float)(1 << bpc));
        }
        bin_w[index] += (double)wgt[i];
    }
    for (uint32_t ibin = 0; ibin < num_bins; ibin++) {
        uint32_t code = ibin;
        for (int k = 0; k < 3; k++) {
            const float c = (float)(code & ((1u << bpc) - 1));
            bin_c[ibin * 3 + (2 - k)] = (c + 0.5f) / (float)(1 << bpc);
            code >>= bpc;
        }
    }
}
