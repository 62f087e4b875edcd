// field_tc.cu — kernels built on the tcgen05 field (field_tc.cuh): the batch field evaluation and the warp-per-ray
// persistent renderer with a 128-sample (4 rays x 32 samples) tensor-core tile per warpgroup.
#include "field_tc.cuh"
#include "hsv_common.cuh"

namespace pnerf {

#ifndef PNERF_TC_GROUPS
#define PNERF_TC_GROUPS 4
#endif
constexpr int kTcGroups = PNERF_TC_GROUPS;                       // warpgroups per CTA (one CTA per SM): 16 warps, <= 128 registers/thread
constexpr int kTcThreads = kTcGroups * 128;
constexpr int kTcSharedBytes = (sizeof(TcShared) + 1023) & ~1023;

__host__ __device__ constexpr size_t tc_smem_bytes(bool clip, size_t extra_per_warp = 0) {
    return (size_t)kTcSharedBytes + (size_t)(clip ? kTcWBytesClip : kTcWBytesNoClip) +
           (size_t)kTcGroups * (clip ? kTcGroupBytesClip : kTcGroupBytesNoClip) + extra_per_warp * kTcGroups * 4;
}

// ------------------------------------------------------------------------------------------------
// batch field evaluation (drop-in for PaletteNetwork.forward in eval mode), 128 samples per warpgroup tile
// ------------------------------------------------------------------------------------------------
template <bool CLIP>
__global__ void __launch_bounds__(kTcThreads, 1)
k_field_forward_tc(const float* __restrict__ xyzs, const float* __restrict__ dirs, uint32_t M, pnerf_palette_field f,
                   float* __restrict__ sigma, float* __restrict__ clip, float* __restrict__ omega, float* __restrict__ off_rad,
                   float* __restrict__ view_dep, float* __restrict__ diffuse) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcShared* sm = reinterpret_cast<TcShared*>(smem_raw);
    unsigned char* wts = smem_raw + kTcSharedBytes;
    unsigned char* groups = wts + (CLIP ? kTcWBytesClip : kTcWBytesNoClip);
    constexpr int group_bytes = CLIP ? kTcGroupBytesClip : kTcGroupBytesNoClip;
    tc_prologue<kTcGroups>(f, f.wpack_tc, sm, wts, groups, group_bytes);
    TcGroup g = tc_make_group(sm, wts, groups, group_bytes);
    const int lane = threadIdx.x & 31, gi = threadIdx.x >> 7;
    const uint32_t n_tiles = ceil_div(M, 128u);
    for (uint32_t tile = blockIdx.x * kTcGroups + gi; tile < n_tiles; tile += gridDim.x * kTcGroups) {
        const uint32_t s = tile * 128 + (uint32_t)g.row;
        const bool active = s < M;
        float x = 0, y = 0, z = 0, dx = 0, dy = 0, dz = 1;
        if (active) {
            x = xyzs[(size_t)s * 3]; y = xyzs[(size_t)s * 3 + 1]; z = xyzs[(size_t)s * 3 + 2];
            dx = dirs[(size_t)s * 3]; dy = dirs[(size_t)s * 3 + 1]; dz = dirs[(size_t)s * 3 + 2];
        }
        FieldOut o;
        eval_field_tc<CLIP ? TC_PALETTE_CLIP : TC_PALETTE>(f, *sm, g, x, y, z, dx, dy, dz, active, lane, o);
        if (active) {
            sigma[s] = o.sigma;
#pragma unroll
            for (int i = 0; i < 3; i++) { diffuse[(size_t)s * 3 + i] = o.diffuse[i]; view_dep[(size_t)s * 3 + i] = o.view_dep[i]; }
#pragma unroll
            for (int i = 0; i < 13; i++) off_rad[(size_t)s * 13 + i] = o.off_rad[i];
#pragma unroll
            for (int b = 0; b < kNB; b++) omega[(size_t)s * kNB + b] = o.omega[b];
            if (CLIP && clip) {
                for (uint32_t i = 0; i < f.clip_dim; i++) clip[(size_t)s * f.clip_dim + i] = o.clip[i];
            }
        }
    }
    tc_epilogue_cta<kTcGroups>(sm);
}

}  // namespace pnerf

using namespace pnerf;

static int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            sms = kNumSMs;
    }
    return sms;
}

extern "C" {

uint32_t pnerf_palette_tc_weight_bytes(uint32_t pred_clip) { return pred_clip ? kTcWBytesClip : kTcWBytesNoClip; }

/* tensor-core (tcgen05) version of pnerf_palette_field_forward: same arguments; needs field->wpack_tc (weight image of
 * palettenerf_b200/fused.py::tc_pack_index) and field->table_sigma_palette (interleaved tables). */
int pnerf_palette_field_forward_tc(const float* xyzs, const float* dirs, uint32_t M, const pnerf_palette_field* field,
                                   float* sigma, float* clip, float* omega, float* off_rad, float* view_dep, float* diffuse,
                                   void* stream) {
    if (M == 0) return PNERF_OK;
    PNERF_REQUIRE(xyzs && dirs && field && sigma && omega && off_rad && view_dep && diffuse);
    PNERF_REQUIRE(field->table_sigma_palette && field->offsets && field->wpack_tc && field->head_bias && field->palette);
    if (field->L != 16 || field->clip_dim > (uint32_t)kClipMax) return PNERF_ERR_UNSUPPORTED;
    if (field->pred_clip && !field->table_clip) return PNERF_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool clip_on = field->pred_clip != 0;
    const size_t smem = tc_smem_bytes(clip_on);
    const uint32_t grid = min(ceil_div(ceil_div(M, 128u), (uint32_t)kTcGroups), (uint32_t)sm_count());
    static bool attr_done[2] = {false, false};
    if (!attr_done[clip_on]) {
        cudaError_t e = clip_on ? cudaFuncSetAttribute(k_field_forward_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                : cudaFuncSetAttribute(k_field_forward_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_cuda_error(e, "field_forward_tc attr"); return PNERF_ERR_CUDA; }
        attr_done[clip_on] = true;
    }
    if (clip_on) k_field_forward_tc<true><<<grid, kTcThreads, smem, s>>>(xyzs, dirs, M, *field, sigma, clip, omega, off_rad, view_dep, diffuse);
    else k_field_forward_tc<false><<<grid, kTcThreads, smem, s>>>(xyzs, dirs, M, *field, sigma, nullptr, omega, off_rad, view_dep, diffuse);
    return check_launch("palette_field_forward_tc");
}

}  // extern "C"

// =================================================================================================================
// warp-per-ray persistent renderer on the tensor-core field
// =================================================================================================================
namespace pnerf {

constexpr int kMaxRuns = 6;                         // occupied stretches of a ray recorded by the pre-pass

// Pre-pass record of one candidate ray: the maximal runs of consecutive occupied lattice points of the reference's walk.
// The lattice inside a run is t_{k+1} = fl(t_k + dt): the renderer regenerates it bit for bit (lattice_window), so a run is
// (first ray parameter, number of points). Rays with more than kMaxRuns runs are walked by the renderer itself.
struct RayRuns {
    uint32_t count;                                 // samples of the ray (0: nothing to render)
    uint32_t n_runs;                                // kMaxRuns + 1 = overflow
    float t_start[kMaxRuns];
    uint32_t n[kMaxRuns];
    float t0;                                       // first lattice point of the walk (== nears (+ noise step))
    uint32_t pad;
};

struct RaysTcArgs {
    const float* rays_o; const float* rays_d; const float* nears; const float* fars; const float* noises;
    const uint8_t* bitfield;
    const float* occ;
    uint32_t N, C, Hgrid, max_steps;
    float dt_gamma, T_thresh;
    float* weights_sum; float* depth; float* image;
    float* direct_rgb; float* view_dep_rgb; float* basis_acc; float* basis_rgb; float* unscaled_basis_rgb;
    float* clip_feat;
    unsigned int* queue;                            // [8]: ray cursor, samples shaded, rays with samples, tiles, candidates
    const int32_t* cand;                            // [N] candidate ray ids
    const RayRuns* runs;                            // [N] indexed by candidate slot
    float* t_scratch;                               // [warps, max_steps]
    const int32_t* out_index;                       // optional [N]: row of the output maps that ray n writes (tile-sharded views)
    pnerf_palette_edit edit;                        // GUI-time edit of the blend (mode 0: none)
    uint32_t share_windows;                         // 1: a ray's last window is filled with the next ray's first samples
};

// edit parameters staged in shared memory (behind the group regions)
struct EditShared {
    float delta_hsv[kNB * 3], mean_xyz[3], mean_clip[kClipMax], dI[kNB], dP[kNB * 3], ddelta[kNB * 9];
};

enum { QT_CURSOR = 0, QT_SAMPLES = 1, QT_RAYS = 2, QT_TILES = 3, QT_CAND = 4 };

__global__ void __launch_bounds__(256) k_tc_candidates(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                       const float* __restrict__ nears, const float* __restrict__ fars,
                                                       uint32_t N, const float* __restrict__ occ, int32_t* __restrict__ cand,
                                                       unsigned int* __restrict__ queue) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
    bool keep = false;
    if (n < N) {
        const float near = nears[n], far = fars[n];
        keep = near < far;
        if (keep && occ) {
            Marcher m;
            m.ox = rays_o[(size_t)n * 3]; m.oy = rays_o[(size_t)n * 3 + 1]; m.oz = rays_o[(size_t)n * 3 + 2];
            m.dx = rays_d[(size_t)n * 3]; m.dy = rays_d[(size_t)n * 3 + 1]; m.dz = rays_d[(size_t)n * 3 + 2];
            m.rdx = 1 / m.dx; m.rdy = 1 / m.dy; m.rdz = 1 / m.dz;
            keep = near < m.occupied_exit(occ);
        }
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, keep);
    if (mask) {
        uint32_t base = 0;
        const uint32_t leader = __ffs(mask) - 1;
        if (lane == leader) base = atomicAdd(queue + QT_CAND, (unsigned int)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (keep) cand[base + __popc(mask & ((1u << lane) - 1u))] = (int32_t)n;
    }
}

// thread per candidate ray: the reference's serial walk (raymarching.cu:351-403 with n_step unbounded), recording runs
__global__ void __launch_bounds__(128) k_tc_prepass(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                    const float* __restrict__ nears, const float* __restrict__ fars,
                                                    const float* __restrict__ noises, const uint8_t* __restrict__ bitfield,
                                                    uint32_t C, uint32_t H, uint32_t max_steps, float bound, float dt_gamma,
                                                    const int32_t* __restrict__ cand, const unsigned int* __restrict__ queue,
                                                    const float* __restrict__ occ, RayRuns* __restrict__ runs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= queue[QT_CAND]) return;
    const uint32_t n = (uint32_t)cand[i];
    Marcher m;
    m.init(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, bound, dt_gamma, max_steps, C, H, bitfield);
    float far = fars[n];
    if (occ) far = fminf(far, m.occupied_exit(occ));
    float t = m.first_t(nears[n], noises ? noises[n] : 0.f);
    RayRuns rr;
    rr.t0 = t; rr.count = 0; rr.n_runs = 0; rr.pad = 0;
#pragma unroll
    for (int k = 0; k < kMaxRuns; k++) { rr.t_start[k] = 0.f; rr.n[k] = 0; }
    uint32_t count = 0, n_runs = 0, run_n = 0;
    float run_t = 0.f;
    bool in_run = false;
    auto close_run = [&]() {
        in_run = false;
        if (n_runs < (uint32_t)kMaxRuns) {
#pragma unroll
            for (int k = 0; k < kMaxRuns; k++)
                if (k == (int)n_runs) { rr.t_start[k] = run_t; rr.n[k] = run_n; }
        }
        n_runs++;
    };
    // (classifying several lattice points ahead inside a run — independent bitfield loads in flight — was measured and
    // rejected: 80 instead of 56 registers and wasted probes cost more than the shorter dependency chains save,
    // 0.44 -> 0.8 ms per 800x800 view)
    float x, y, z, dt;
    if (dt_gamma == 0.f) {
        // constant step: an occupied point's voxel is probed ONCE; the lattice points that provably stay inside it
        // (Marcher::probe_point<true>) cost one add and one compare each instead of a probe (~4.6 points per voxel)
        while (t < far && count < max_steps) {
            float tt;
            if (m.probe_point<true>(t, x, y, z, dt, tt)) {
                if (!in_run) { in_run = true; run_t = t; run_n = 0; }
                const float stop = fminf(tt, far);
                do {
                    run_n++;
                    count++;
                    t += dt;
                } while (t < stop && count < max_steps);
            } else {
                t = m.advance_past(t, tt);
                if (in_run) close_run();
            }
        }
    } else {
        while (t < far && count < max_steps) {
            if (m.probe(t, x, y, z, dt)) {
                if (!in_run) { in_run = true; run_t = t; run_n = 0; }
                run_n++;
                count++;
                t += dt;
            } else if (in_run) {          // probe() advanced t past the empty voxel: the run is closed
                close_run();
            }
        }
    }
    if (in_run) close_run();
    rr.count = count;
    rr.n_runs = min(n_runs, (uint32_t)kMaxRuns + 1u);
    runs[i] = rr;
}

__host__ __device__ constexpr size_t tc_render_smem(bool clip) { return tc_smem_bytes(clip) + sizeof(EditShared) + 16; }

// MODE: TC_PALETTE / TC_PALETTE_CLIP / TC_NERF (stage-1 model: colour = the colour net's output, no palette blend)
// EDIT: 0 = plain palette blend, 1 = RegionEdit, 2 = Stylizer (csrc: pnerf_palette_edit)
template <int MODE, bool AUX, int EDIT>
__global__ void __launch_bounds__(kTcThreads, 1) k_render_rays_tc(RaysTcArgs a, pnerf_palette_field f) {
    constexpr bool CLIP = MODE == TC_PALETTE_CLIP;
    static_assert(MODE != TC_NERF || (!AUX && EDIT == 0), "the stage-1 model has no palette maps and no edits");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcShared* sm = reinterpret_cast<TcShared*>(smem_raw);
    unsigned char* wts = smem_raw + kTcSharedBytes;
    unsigned char* groups = wts + (CLIP ? kTcWBytesClip : kTcWBytesNoClip);
    constexpr int group_bytes = CLIP ? kTcGroupBytesClip : kTcGroupBytesNoClip;
    constexpr bool kCoreInScratch = AUX;               // the five core sums as scratch channels: 5.06 -> 4.93 ms per view
    constexpr int kRedCh = kAuxCh + (kCoreInScratch ? 5 : 0);      // + weights_sum, depth, r, g, b
    static_assert(kRedCh <= 4 * (group_bytes / kTcChunk) && kRedCh <= 64 && kClipMax <= 4 * (group_bytes / kTcChunk),
                  "per-warp reduction scratch: 4 channels per k-chunk of the warp's own rows");
    EditShared* ed = reinterpret_cast<EditShared*>(groups + kTcGroups * group_bytes);
    if (EDIT != 0) {
        const int tid = threadIdx.x;
        if (EDIT == 1) {
            if (tid < kNB * 3) ed->delta_hsv[tid] = a.edit.delta_hsv[tid];
            if (tid < 3) ed->mean_xyz[tid] = a.edit.mean_xyz ? a.edit.mean_xyz[tid] : 0.f;
            if (tid < kClipMax) ed->mean_clip[tid] = (a.edit.mean_clip && tid < (int)f.clip_dim) ? a.edit.mean_clip[tid] : 0.f;
        } else {
            if (tid < kNB) ed->dI[tid] = a.edit.dI[tid];
            if (tid < kNB * 3) ed->dP[tid] = a.edit.dP[tid];
            if (tid < kNB * 9) ed->ddelta[tid] = a.edit.ddelta[tid];
        }
    }
    tc_prologue<kTcGroups>(f, f.wpack_tc, sm, wts, groups, group_bytes);
    TcGroup g = tc_make_group(sm, wts, groups, group_bytes);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, gi = wid >> 2, wig = wid & 3;
    // after the field of a tile is complete the group's input regions are dead until the next gather: THIS WARP'S OWN ROWS of
    // them (32 rows x 16 B = 512 B in each of the 10 / 14 k-chunks; nobody else ever writes there) are its scratch for the
    // per-tile channel reduction. Channel c = 128 B in chunk c / 4; sample l of channel c sits in 16-byte segment
    // (l / 4) ^ (c & 7), word l & 3: the row-wise writes (lane = sample) hit 32 banks, the column sums (lane = channel) read
    // whole segments (LDS.128) whose positions differ across the 8 lanes of a quarter-warp — conflict-free both ways.
    unsigned char* const red_base = g.smem + wig * 512;
    auto red_row = [red_base](int c) -> unsigned char* { return red_base + (c >> 2) * kTcChunk + (c & 3) * 128; };
    auto red = [&](int c, int l) -> float& {               // element (channel c, sample l)
        return *reinterpret_cast<float*>(red_row(c) + (((l << 2) ^ ((c & 7) << 4))));
    };
    // sum over the 32 samples of channel c (c = this lane's channel); `split`: samples >= na go to `rest` instead
    auto red_sum = [&](int c, bool split, uint32_t na_, float& rest) -> float {
        const uint32_t y = (uint32_t)(c & 7) << 4;
        unsigned char* const row = red_row(c);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 v = *reinterpret_cast<const float4*>(row + (((uint32_t)j << 4) ^ y));
            if (!split) {
                s += (v.x + v.y) + (v.z + v.w);
            } else {
                const uint32_t l0 = (uint32_t)j << 2;      // first sample of this (logical) segment
                if (l0 + 0 < na_) s += v.x; else rest += v.x;
                if (l0 + 1 < na_) s += v.y; else rest += v.y;
                if (l0 + 2 < na_) s += v.z; else rest += v.z;
                if (l0 + 3 < na_) s += v.w; else rest += v.w;
            }
        }
        return s;
    };
    // two t-lists per warp: the current ray's and the one of the ray that fills the free lanes of the current ray's last window
    float* t_list = a.t_scratch + (size_t)(blockIdx.x * (kTcGroups * 4) + wid) * 2 * a.max_steps;
    float* t_next = t_list + a.max_steps;
    float* const pend = sm->pend[wid];                     // origin / direction of the ray in t_next
    const bool clip_on = CLIP && a.clip_feat != nullptr;
    const uint32_t n_cand = a.queue[QT_CAND];
    const float dt_min = 2 * 1.7320508075688772f / a.max_steps;
    const float dt_max = 2 * 1.7320508075688772f * (1u << (a.C - 1)) / a.Hgrid;
    uint32_t shaded = 0, tiles = 0, hit_rays = 0, vote = 0;

    // per-warp state of the CURRENT ray (warp-uniform unless noted)
    bool has_ray = false, exhausted = false;
    uint32_t ray = 0, count = 0, done = 0;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 1;
    float T_run = 1.f;
    float wsum = 0.f, dep = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;   // ray totals (warp-uniform)
    float acc_lo = 0.f, acc_hi = 0.f, acc_clip = 0.f;      // per-LANE: totals of aux channel `lane`, `lane + 32`, clip channel `lane`

    // retire: one writer per value; lane = auxiliary channel
    // (out_index: this rank renders a shard of a view whose maps live elsewhere — possibly in a peer GPU's memory, mapped over
    // NVLink: the stores below are the gather)
    auto retire = [&](uint32_t r, float s_w, float s_d, float s_r, float s_g, float s_b, float lo, float hi, float cl) {
        if (a.out_index) r = (uint32_t)a.out_index[r];
        if (!kCoreInScratch && lane == 0) {
            a.weights_sum[r] = s_w; a.depth[r] = s_d;
            a.image[(size_t)r * 3] = s_r; a.image[(size_t)r * 3 + 1] = s_g; a.image[(size_t)r * 3 + 2] = s_b;
        }
        if (AUX) {
            auto dst_of = [&](int c) -> float* {           // 0-2 direct, 3-5 view_dep, 6-9 basis_acc, 10-21 basis_rgb, 22-33 unscaled
                if (c < 3) return a.direct_rgb + (size_t)r * 3 + c;
                if (c < 6) return a.view_dep_rgb + (size_t)r * 3 + (c - 3);
                if (c < 6 + kNB) return a.basis_acc + (size_t)r * kNB + (c - 6);
                if (c < 6 + kNB + kNB * 3) return a.basis_rgb + (size_t)r * kNB * 3 + (c - 6 - kNB);
                return a.unscaled_basis_rgb + (size_t)r * kNB * 3 + (c - 6 - kNB - kNB * 3);
            };
            *dst_of(lane) = lo;
            const int c = lane + 32;                        // second channel of this lane: aux 32, 33 (then the five core sums)
            if (c < kAuxCh) *dst_of(c) = hi;
            else if (kCoreInScratch && c == kAuxCh) a.weights_sum[r] = hi;
            else if (kCoreInScratch && c == kAuxCh + 1) a.depth[r] = hi;
            else if (kCoreInScratch && c < kRedCh) a.image[(size_t)r * 3 + (c - kAuxCh - 2)] = hi;
        }
        if (CLIP && clip_on && lane < (int)f.clip_dim) a.clip_feat[(size_t)r * f.clip_dim + lane] = cl;
    };

    for (;;) {
        // ---- this warp's window: the next samples of its ray (lanes < na); when those do not fill the window (the ray's
        //      last one) the free lanes take the first samples of the NEXT ray from the queue (lanes na .. na + nb - 1), so
        //      that the 128-row MMA tiles stay full: tile fill 0.90 -> 0.97 on the 800x800 view ----
        const uint32_t na = has_ray ? min(32u, count - done) : 0u;
        bool has_next = false;
        uint32_t ray_n = 0, count_n = 0;
        while (na < 32u && !has_next && !exhausted && (a.share_windows || !has_ray)) {
            uint32_t slot = 0;
            if (lane == 0) slot = atomicAdd(a.queue + QT_CURSOR, 1u);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot >= n_cand) { exhausted = true; break; }
            const RayRuns* rr = a.runs + slot;
            count_n = rr->count;
            if (count_n == 0) continue;
            ray_n = (uint32_t)a.cand[slot];
            Marcher m;
            m.init(a.rays_o + (size_t)ray_n * 3, a.rays_d + (size_t)ray_n * 3, f.bound, a.dt_gamma, a.max_steps, a.C, a.Hgrid, a.bitfield);
            const uint32_t n_runs = rr->n_runs;
            const float t0 = rr->t0;
            if (n_runs > (uint32_t)kMaxRuns) {          // too many stretches for the record: walk the ray here
                float far = a.fars[ray_n];
                if (a.occ) far = fminf(far, m.occupied_exit(a.occ));
                count_n = warp_walk<false>(m, t0, far, a.max_steps, (uint32_t)lane, nullptr, nullptr, nullptr, t_next);
            } else {
                uint32_t pos = 0;
                for (uint32_t r = 0; r < n_runs; r++) {
                    float ts = rr->t_start[r];
                    uint32_t left = rr->n[r];
                    while (left > 0) {
                        float t;
                        uint32_t nvalid;
                        lattice_window(m, ts, (uint32_t)lane, t, nvalid);
                        const uint32_t nv = min(nvalid, left);
                        if ((uint32_t)lane < nv) t_next[pos + lane] = t;
                        pos += nv;
                        left -= nv;
                        ts = __shfl_sync(0xffffffffu, t + m.step_size(t), nv - 1);
                    }
                }
            }
            if (lane == 0) { pend[0] = m.ox; pend[1] = m.oy; pend[2] = m.oz; pend[3] = m.dx; pend[4] = m.dy; pend[5] = m.dz; }
            __syncwarp();
            if (count_n == 0) continue;
            has_next = true;
            hit_rays++;
        }
        // ---- group vote: the field is evaluated until no warp of the group had a tile. The warps meet here (in phase they run
        //      better: without this barrier 5.22 instead of 5.05 ms per 800x800 view), but the votes are READ after the field:
        //      no shared-memory load and branch between the barrier and the gather (5.22 -> 5.05 ms as well). Price: one empty
        //      tile per group at the very end. Two sets of flags: the next vote is written while a slow warp may still be
        //      reading this one. ----
        if (lane == 0) sm->flags[vote][gi][wig] = (has_ray || has_next) ? 1u : 0u;
        tc::group_bar(g.bar_id, 128);

        const bool in_next = (uint32_t)lane >= na;          // lanes of the next ray (or free)
        const uint32_t k = in_next ? (uint32_t)lane - na : done + (uint32_t)lane;
        const bool active = in_next ? (has_next && k < count_n) : true;
        const float lox = in_next ? pend[0] : ox, loy = in_next ? pend[1] : oy, loz = in_next ? pend[2] : oz;
        const float ldx = in_next ? pend[3] : dx, ldy = in_next ? pend[4] : dy, ldz = in_next ? pend[5] : dz;
        const float t = active ? (in_next ? t_next[k] : t_list[k]) : 0.f;
        const float x = clampf(lox + t * ldx, -f.bound, f.bound);
        const float y = clampf(loy + t * ldy, -f.bound, f.bound);
        const float z = clampf(loz + t * ldz, -f.bound, f.bound);
        const float dt = clampf(t * a.dt_gamma, dt_min, dt_max);
        const float t_end = t + dt;
        FieldOut o;
        eval_field_tc<MODE, true>(f, *sm, g, x, y, z, active ? ldx : 0.f, active ? ldy : 0.f, active ? ldz : 1.f, active, lane, o);   // x, y, z clamped above
        const uint32_t any = sm->flags[vote][gi][0] | sm->flags[vote][gi][1] | sm->flags[vote][gi][2] | sm->flags[vote][gi][3];
        vote ^= 1u;
        if (!any) break;
        if (!has_ray && !has_next) continue;            // (warp-uniform) idle warp of a busy group
        tiles++;

        // ---- front-to-back compositing of the window (ref: raymarching.cu:1051-1110 per sample); the product scan of
        //      (1 - alpha) restarts at lane na, where the next ray begins ----
        const float alpha = active ? 1.0f - fast_exp(-(f.density_scale * o.sigma) * dt) : 0.f;
        float incl = 1.0f - alpha;
        const int seg0 = in_next ? (int)na : 0;
#pragma unroll
        for (int ofs = 1; ofs < 32; ofs <<= 1) {
            const float up = __shfl_up_sync(0xffffffffu, incl, ofs);
            if (lane - ofs >= seg0) incl *= up;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == seg0) excl = 1.f;
        const float T = (in_next ? 1.f : T_run) * excl;
        const uint32_t below = __ballot_sync(0xffffffffu, active && T < a.T_thresh);
        const uint32_t mask_a = na >= 32u ? 0xffffffffu : ((1u << na) - 1u);
        const uint32_t term = below & mask_a, term_n = below & ~mask_a;
        const int last = in_next ? (term_n ? (__ffs(term_n) - 1) : 31) : (term ? (__ffs(term) - 1) : 31);
        const bool use = active && lane <= last;
        const float wgt = use ? alpha * T : 0.f;
        shaded += __popc(__ballot_sync(0xffffffffu, use));
        // window totals of the next ray's lanes (the current ray's go straight into its running totals)
        float n_w = 0.f, n_d = 0.f, n_r = 0.f, n_g = 0.f, n_b = 0.f, n_lo = 0.f, n_hi = 0.f, n_clip = 0.f;
        {
            float rgb[3], basis_rgb[kNB * 3], unscaled[kNB * 3];
            const float sp = MODE == TC_NERF ? 0.f : softplusf_(o.off_rad[12]);
            rgb[0] = rgb[1] = rgb[2] = 0.f;
            if (MODE == TC_NERF) {
#pragma unroll
                for (int c = 0; c < 3; c++) rgb[c] = o.view_dep[c];
            } else if (EDIT == 2) {
                // Stylizer.forward (ref: palette/renderer.py:166-183)
#pragma unroll
                for (int b = 0; b < kNB; b++) {
                    const float gain = fmaxf(sp + ed->dI[b], 0.f);
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        float off = 0.f;
#pragma unroll
                        for (int i = 0; i < 3; i++) off += o.off_rad[b * 3 + i] * ed->ddelta[b * 9 + i * 3 + c];
                        const float col = fminf(fmaxf(gain * (sm->palette[b * 3 + c] + ed->dP[b * 3 + c] + off), 0.f), 1.f);
                        basis_rgb[b * 3 + c] = unscaled[b * 3 + c] = 0.f;
                        rgb[c] += o.omega[b] * col;
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; c++) rgb[c] += o.view_dep[c];
            } else {
                float ew = 1.f;
                if (EDIT == 1) {
                    // RegionEdit.forward's spatial / semantic weight (ref: palette/renderer.py:127-135)
                    if (a.edit.mean_xyz) {
                        const float ex = x - ed->mean_xyz[0], ey = y - ed->mean_xyz[1], ez = z - ed->mean_xyz[2];
                        ew *= __expf(-(ex * ex + ey * ey + ez * ez) / a.edit.std_xyz);
                    }
                    if (a.edit.mean_clip) {
                        float d2 = 0.f;
#pragma unroll
                        for (int i = 0; i < kClipMax; i++) {
                            const float e = (i < (int)f.clip_dim) ? o.clip[i] - ed->mean_clip[i] : 0.f;
                            d2 += e * e;
                        }
                        ew *= __expf(-d2 / a.edit.std_clip);
                    }
                }
#pragma unroll
                for (int b = 0; b < kNB; b++) {
                    float fin[3];
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float off = o.off_rad[b * 3 + c];
                        unscaled[b * 3 + c] = sm->palette[b * 3 + c] + off;
                        fin[c] = sp * (sm->palette[b * 3 + c] + f.offsets_weight * off);
                    }
                    if (EDIT == 1) {
                        if (a.edit.weight_mode) {
                            fin[0] = fin[1] = fin[2] = ew;
                        } else {
                            float h, sa, v, nr, ng, nb2;
                            rgb_to_hsv_dev(fin[0], fin[1], fin[2], h, sa, v);
                            h = fmodf(h + ed->delta_hsv[b * 3] + 360.f, 360.f);
                            sa = fmaxf(sa * ed->delta_hsv[b * 3 + 1], 0.f);
                            v = fmaxf(v * ed->delta_hsv[b * 3 + 2], 0.f);
                            hsv_to_rgb_dev(h, sa, v, nr, ng, nb2);
                            fin[0] += ew * (nr - fin[0]); fin[1] += ew * (ng - fin[1]); fin[2] += ew * (nb2 - fin[2]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        basis_rgb[b * 3 + c] = o.omega[b] * fin[c];
                        rgb[c] += basis_rgb[b * 3 + c];
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; c++) rgb[c] += f.view_dep_weight * o.view_dep[c];
            }
            if (kCoreInScratch) {
                // weights_sum, depth and the colour ride in the scratch as channels kAuxCh .. kAuxCh + 4 (summed by lanes 2-6
                // in the second column pass, which runs anyway): no shuffle reductions
                red(kAuxCh, lane) = wgt;
                red(kAuxCh + 1, lane) = wgt * t_end;
#pragma unroll
                for (int c = 0; c < 3; c++) red(kAuxCh + 2 + c, lane) = wgt * rgb[c];
            } else if (!has_next) {
                wsum += warp_sum(wgt);
                dep += warp_sum(wgt * t_end);
                cr += warp_sum(wgt * rgb[0]); cg += warp_sum(wgt * rgb[1]); cb += warp_sum(wgt * rgb[2]);
            } else {
                const float wa = in_next ? 0.f : wgt, wn = in_next ? wgt : 0.f;
                wsum += warp_sum(wa); n_w = warp_sum(wn);
                dep += warp_sum(wa * t_end); n_d = warp_sum(wn * t_end);
                cr += warp_sum(wa * rgb[0]); n_r = warp_sum(wn * rgb[0]);
                cg += warp_sum(wa * rgb[1]); n_g = warp_sum(wn * rgb[1]);
                cb += warp_sum(wa * rgb[2]); n_b = warp_sum(wn * rgb[2]);
            }
            if (AUX) {
                // channel-major scratch red[c][lane]; the column sums below run with lane = channel
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    red(c, lane) = wgt * (o.diffuse[c] + o.view_dep[c]);
                    red(3 + c, lane) = wgt * o.view_dep[c];
                }
#pragma unroll
                for (int q = 0; q < kNB; q++) red(6 + q, lane) = wgt * o.omega[q];
#pragma unroll
                for (int q = 0; q < kNB * 3; q++) {
                    red(6 + kNB + q, lane) = wgt * basis_rgb[q];
                    red(6 + kNB + kNB * 3 + q, lane) = wgt * unscaled[q];
                }
            }
        }
        if (AUX) {
            __syncwarp();
            float dummy = 0.f;
            if (!has_next) {
                acc_lo += red_sum(lane, false, 32u, dummy);
                if (lane + 32 < kRedCh) acc_hi += red_sum(lane + 32, false, 32u, dummy);
            } else {                                        // column sums split at lane na
                acc_lo += red_sum(lane, true, na, n_lo);
                if (lane + 32 < kRedCh) acc_hi += red_sum(lane + 32, true, na, n_hi);
            }
        }
        if (CLIP && clip_on) {
            __syncwarp();
#pragma unroll
            for (int q = 0; q < kClipMax; q++) red(q, lane) = wgt * o.clip[q];
            __syncwarp();
            if (lane < kClipMax) acc_clip += red_sum(lane, true, na, n_clip);
        }
        __syncwarp();
        // ---- the current ray: finished when it ran out of samples (always the case when the window was shared) or below
        //      the transmittance threshold ----
        if (has_ray) {
            done += na;
            if (term || done >= count) {
                retire(ray, wsum, dep, cr, cg, cb, acc_lo, acc_hi, acc_clip);
                has_ray = false;
            } else {
                T_run *= __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        // ---- the next ray becomes the current one ----
        if (has_next) {
            const float T_next = __shfl_sync(0xffffffffu, incl, 31);
            done = min(32u - na, count_n);
            if (term_n || done >= count_n) {
                retire(ray_n, n_w, n_d, n_r, n_g, n_b, n_lo, n_hi, n_clip);
            } else {
                has_ray = true;
                ray = ray_n; count = count_n;
                ox = pend[0]; oy = pend[1]; oz = pend[2]; dx = pend[3]; dy = pend[4]; dz = pend[5];
                T_run = T_next;
                wsum = n_w; dep = n_d; cr = n_r; cg = n_g; cb = n_b;
                acc_lo = n_lo; acc_hi = n_hi; acc_clip = n_clip;
                float* const tmp = t_list; t_list = t_next; t_next = tmp;
            }
            __syncwarp();                                   // pend / t_next are rewritten by the next fetch
        }
    }
    if (lane == 0 && tiles) {
        atomicAdd(a.queue + QT_SAMPLES, shaded);
        atomicAdd(a.queue + QT_RAYS, hit_rays);
        atomicAdd(a.queue + QT_TILES, tiles);
    }
    tc_epilogue_cta<kTcGroups>(sm);
}

}  // namespace pnerf

static bool g_tc_timing = false, g_tc_timed = false;
static cudaEvent_t g_tc_ev[2] = {nullptr, nullptr};

extern "C" {

/* rows of max_steps floats the renderer needs in t_scratch: two per resident warp (current ray + the ray sharing its last window) */
uint32_t pnerf_palette_render_tc_warps(void) { return 2u * (uint32_t)sm_count() * (uint32_t)(kTcGroups * 4); }
uint32_t pnerf_palette_render_tc_runs_bytes(void) { return (uint32_t)sizeof(RayRuns); }

/* Tensor-core warp-per-ray renderer (csrc/field_tc.cu): like pnerf_palette_render_rays, with the field on tcgen05 / TMEM and
 * a thread-per-ray pre-pass that records each ray's occupied stretches.
 *   queue [8] u32 zero on entry; cand [N] int32; runs [N * pnerf_palette_render_tc_runs_bytes()] bytes;
 *   t_scratch [pnerf_palette_render_tc_warps() * max_steps] fp32. Needs field->wpack_tc and field->table_sigma_palette.
 *   out_index (optional, [N] int32): ray n writes row out_index[n] of the output maps instead of row n — the maps may be a
 *   peer GPU's memory (one view sharded over several GPUs: every rank stores its rays straight into the owner's image).
 *   flags: PNERF_RENDER_REPRODUCIBLE — see include/pnerf_b200.h */
int pnerf_palette_render_tc(const float* rays_o, const float* rays_d, const float* nears, const float* fars, const float* noises,
                            const uint8_t* bitfield, uint32_t N, uint32_t C, uint32_t Hgrid, uint32_t max_steps, float dt_gamma,
                            float T_thresh, const pnerf_palette_field* field, float* weights_sum, float* depth, float* image,
                            float* direct_rgb, float* view_dep_rgb, float* basis_acc, float* basis_rgb, float* unscaled_basis_rgb,
                            float* clip_feat, uint32_t* queue, int32_t* cand, void* runs, float* t_scratch, const float* occ_aabb,
                            const int32_t* out_index, const pnerf_palette_edit* edit, uint32_t flags, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(rays_o && rays_d && nears && fars && bitfield && field && weights_sum && depth && image && queue);
    PNERF_REQUIRE(cand && runs && t_scratch);
    const bool nerf = field->model_kind == 1;
    PNERF_REQUIRE((nerf ? field->table_sigma : field->table_sigma_palette) && field->offsets && field->wpack_tc &&
                  field->head_bias && field->palette);
    PNERF_REQUIRE(C >= 1 && C <= 16 && Hgrid >= 1 && max_steps >= 1);
    if (field->L != 16 || field->clip_dim > (uint32_t)kClipMax || Hgrid > 1024) return PNERF_ERR_UNSUPPORTED;
    if (field->pred_clip && !field->table_clip) return PNERF_ERR_INVALID_ARG;
    const bool aux = direct_rgb != nullptr;
    if (aux) PNERF_REQUIRE(view_dep_rgb && basis_acc && basis_rgb && unscaled_basis_rgb);
    if (nerf) PNERF_REQUIRE(!aux && !clip_feat && !field->pred_clip && (!edit || edit->mode == 0));
    RaysTcArgs a;
    a.rays_o = rays_o; a.rays_d = rays_d; a.nears = nears; a.fars = fars; a.noises = noises; a.bitfield = bitfield; a.occ = occ_aabb;
    a.N = N; a.C = C; a.Hgrid = Hgrid; a.max_steps = max_steps; a.dt_gamma = dt_gamma; a.T_thresh = T_thresh;
    a.weights_sum = weights_sum; a.depth = depth; a.image = image;
    a.direct_rgb = direct_rgb; a.view_dep_rgb = view_dep_rgb; a.basis_acc = basis_acc; a.basis_rgb = basis_rgb;
    a.unscaled_basis_rgb = unscaled_basis_rgb; a.clip_feat = clip_feat; a.queue = queue; a.cand = cand;
    a.runs = (const RayRuns*)runs; a.t_scratch = t_scratch; a.out_index = out_index;
    a.share_windows = (flags & PNERF_RENDER_REPRODUCIBLE) ? 0u : 1u;
    pnerf_palette_edit none = {};
    a.edit = edit ? *edit : none;
    const int emode = (int)a.edit.mode;
    PNERF_REQUIRE(emode >= 0 && emode <= 2);
    if (emode == 1) PNERF_REQUIRE(a.edit.delta_hsv != nullptr);
    if (emode == 2) PNERF_REQUIRE(a.edit.dI && a.edit.dP && a.edit.ddelta && !aux);   // the Stylizer produces no debug maps
    cudaStream_t s = (cudaStream_t)stream;
    k_tc_candidates<<<ceil_div(N, 256u), 256, 0, s>>>(rays_o, rays_d, nears, fars, N, occ_aabb, cand, queue);
    k_tc_prepass<<<ceil_div(N, 128u), 128, 0, s>>>(rays_o, rays_d, nears, fars, noises, bitfield, C, Hgrid, max_steps, field->bound,
                                                   dt_gamma, cand, queue, occ_aabb, (RayRuns*)runs);
    const bool clip_on = field->pred_clip != 0;
    const size_t smem = tc_render_smem(clip_on);
    const uint32_t grid = (uint32_t)sm_count();
    static bool attr_done[2][2][3] = {};
#define PNERF_LAUNCH_TC(CL, AX, ED)                                                                                      \
    do {                                                                                                                \
        if (!attr_done[CL][AX][ED]) {                                                                                   \
            cudaError_t e = cudaFuncSetAttribute(k_render_rays_tc<(CL) ? TC_PALETTE_CLIP : TC_PALETTE, AX, ED>,              \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_render_smem(CL)); \
            if (e != cudaSuccess) { set_last_cuda_error(e, "render_tc attr"); return PNERF_ERR_CUDA; }                  \
            attr_done[CL][AX][ED] = true;                                                                               \
        }                                                                                                               \
        k_render_rays_tc<(CL) ? TC_PALETTE_CLIP : TC_PALETTE, AX, ED><<<grid, kTcThreads, smem, s>>>(a, *field);         \
    } while (0)
#define PNERF_LAUNCH_TC_ED(CL, AX)                                                                                       \
    do {                                                                                                                \
        if (emode == 0) PNERF_LAUNCH_TC(CL, AX, 0); else PNERF_LAUNCH_TC(CL, AX, 1);                                    \
    } while (0)
    if (g_tc_timing) {
        if (!g_tc_ev[0]) { cudaEventCreate(&g_tc_ev[0]); cudaEventCreate(&g_tc_ev[1]); }
        cudaEventRecord(g_tc_ev[0], s);
    }
    if (nerf) {
        static bool nerf_attr = false;
        if (!nerf_attr) {
            cudaError_t e = cudaFuncSetAttribute(k_render_rays_tc<TC_NERF, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)tc_render_smem(false));
            if (e != cudaSuccess) { set_last_cuda_error(e, "render_tc attr"); return PNERF_ERR_CUDA; }
            nerf_attr = true;
        }
        k_render_rays_tc<TC_NERF, false, 0><<<grid, kTcThreads, smem, s>>>(a, *field);
    }
    else if (emode == 2) { if (clip_on) PNERF_LAUNCH_TC(true, false, 2); else PNERF_LAUNCH_TC(false, false, 2); }
    else if (clip_on) { if (aux) PNERF_LAUNCH_TC_ED(true, true); else PNERF_LAUNCH_TC_ED(true, false); }
    else { if (aux) PNERF_LAUNCH_TC_ED(false, true); else PNERF_LAUNCH_TC_ED(false, false); }
#undef PNERF_LAUNCH_TC_ED
#undef PNERF_LAUNCH_TC
    if (g_tc_timing) { cudaEventRecord(g_tc_ev[1], s); g_tc_timed = true; }
    return check_launch("palette_render_tc");
}

void pnerf_render_tc_timing(int enable) { g_tc_timing = enable != 0; g_tc_timed = false; }

float pnerf_render_tc_last_ms(void) {
    if (!g_tc_timed) return -1.0f;
    float ms = -1.0f;
    if (cudaEventSynchronize(g_tc_ev[1]) != cudaSuccess || cudaEventElapsedTime(&ms, g_tc_ev[0], g_tc_ev[1]) != cudaSuccess) return -1.0f;
    return ms;
}

}  // extern "C"
