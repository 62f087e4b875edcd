// fused.cu — the fused per-sample field of PaletteNeRF (hash grids -> sigma / diffuse / view-dependent / palette-basis /
// semantic MLPs -> palette heads) and the persistent fused renderer built on it, for B200 (sm_100a).
//
// The reference evaluates this field as ~14 cuBLAS GEMMs + ~30 elementwise kernels per call, round-tripping every
// [M,64] activation through HBM (palette/network.py:156-280), inside a host loop of march / shade / composite /
// compact launches (palette/renderer.py:430-523). Here:
//
//   * one warp owns 32 samples. Lanes gather the 16-level hash-grid features of their own sample (8 corners x 2
//     levels in flight, fp16x2 entries, fp32 interpolation), park them as fp16 rows in a per-warp shared-memory
//     tile, and the warp then runs the whole MLP chain on tensor cores: mma.sync.m16n8k16 (f16 x f16 -> f32) with
//     the A operand chained layer to layer IN REGISTERS (an m16n8 accumulator pair is exactly an m16k16 A fragment)
//     and the B operand (all ~20 k weights, 40-46 KB) resident in shared memory in fragment order, so a layer is
//     one conflict-free LDS.64 + one HMMA per (k-step, n-tile). No activation ever touches HBM.
//     Why mma.sync and not tcgen05: per sample the field needs ~10 KB of L2 gather traffic against 21 k MACs, so the
//     kernel is bound by the gathers (L2 sector throughput), not by the tensor pipe; the 64-wide layers would fill
//     at most a 128x64 UMMA tile per warpgroup, and the TMEM round trip (tcgen05.ld -> activation -> st.shared ->
//     fence -> next MMA) per layer costs more than the register-chained HMMA path saves. DESIGN.md has the numbers.
//   * the concatenations of the reference ([SH16 | geo15], [grid32 | diffuse3], geo = h[1:16]) cost nothing: the
//     weight columns are permuted / zero-padded when the fragments are packed on the host (fused.py), so the
//     register fragments of one layer feed the next as they are.
//   * the renderer is ONE persistent kernel: each lane owns a ray (origin, direction, march position, compositing
//     accumulators in registers), marches to its next occupied sample, the warp evaluates the field for its 32
//     samples, lanes composite, finished rays are replaced from a global queue (one warp-aggregated atomic). There is
//     no host loop, no alive-list compaction, no xyzs/dirs/deltas/sigmas/rgbs buffer and no host synchronisation.
#include "fused_field.cuh"

namespace pnerf {

// ------------------------------------------------------------------------------------------------
// kernel 1: field evaluation for a batch of samples (drop-in for PaletteNetwork.forward in eval mode)
// ------------------------------------------------------------------------------------------------
template <bool CLIP>
__global__ void __launch_bounds__(kFusedWarps * 32, 1)
k_field_forward(const float* __restrict__ xyzs, const float* __restrict__ dirs, uint32_t M, pnerf_palette_field f,
                float* __restrict__ sigma, float* __restrict__ clip, float* __restrict__ omega,
                float* __restrict__ off_rad, float* __restrict__ view_dep, float* __restrict__ diffuse) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem* sm = reinterpret_cast<FusedSmem*>(smem_raw);
    uint2* wts = reinterpret_cast<uint2*>(smem_raw + ((sizeof(FusedSmem) + 15) & ~(size_t)15));
    WarpScratch* scratch = reinterpret_cast<WarpScratch*>(wts + (f.pred_clip ? kWUnitsClip : kWUnitsNoClip));
    fused_prologue(f, sm, wts);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpScratch& ws = scratch[wid];
    const uint32_t n_tiles = ceil_div(M, 32u);
    for (uint32_t tile = blockIdx.x * kFusedWarps + wid; tile < n_tiles; tile += gridDim.x * kFusedWarps) {
        const uint32_t s = tile * 32 + lane;
        const bool active = s < M;
        float x = 0, y = 0, z = 0, dx = 0, dy = 0, dz = 1;
        if (active) {
            x = xyzs[(size_t)s * 3]; y = xyzs[(size_t)s * 3 + 1]; z = xyzs[(size_t)s * 3 + 2];
            dx = dirs[(size_t)s * 3]; dy = dirs[(size_t)s * 3 + 1]; dz = dirs[(size_t)s * 3 + 2];
        }
        FieldOut o;
        eval_field<CLIP>(f, *sm, wts, ws, x, y, z, dx, dy, dz, active, lane, o);
        if (active) {
            sigma[s] = o.sigma;
#pragma unroll
            for (int i = 0; i < 3; i++) { diffuse[(size_t)s * 3 + i] = o.diffuse[i]; view_dep[(size_t)s * 3 + i] = o.view_dep[i]; }
#pragma unroll
            for (int i = 0; i < 13; i++) off_rad[(size_t)s * 13 + i] = o.off_rad[i];
#pragma unroll
            for (int b = 0; b < kNB; b++) omega[(size_t)s * kNB + b] = o.omega[b];
            if (clip) {
                for (uint32_t i = 0; i < f.clip_dim; i++) clip[(size_t)s * f.clip_dim + i] = o.clip[i];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// kernel 2: persistent fused renderer (march -> field -> blend -> composite), one ray per lane
// ------------------------------------------------------------------------------------------------
struct RenderArgs {
    const float* rays_o; const float* rays_d; const float* nears; const float* fars; const float* noises;  // noises may be NULL
    const uint8_t* bitfield;
    uint32_t N, C, Hgrid, max_steps;
    float dt_gamma, T_thresh;
    float* weights_sum; float* depth; float* image;          // [N], [N], [N,3]   (written once per ray)
    float* direct_rgb; float* view_dep_rgb; float* basis_acc; float* basis_rgb; float* unscaled_basis_rgb;  // aux (NULL in gui mode)
    float* clip_feat;                                         // [N, clip_dim] or NULL
    unsigned int* queue;                                      // [4]: next hit-list slot, samples shaded, rays with samples, 32-sample tiles evaluated
    const int32_t* hit_list;                                  // [N] ids of the rays that own at least one sample
    const float* t_first; const float* t_last;                // [N] lattice t of each ray's first / last occupied point
};

// ------------------------------------------------------------------------------------------------
// pre-pass: one thread per ray walks the occupancy grid once (same lattice walk as the reference's march) and
// records the first and last occupied lattice point. Rays without samples (79% of an 800x800 lego view) never
// enter the persistent kernel, and rays inside it never march the empty tail behind their last sample.
// ------------------------------------------------------------------------------------------------
constexpr int kLptBuckets = 32;          // buckets of 16 samples; rays with > 496 samples share the last one
constexpr int kQueueHist = 4, kQueueCursor = 4 + kLptBuckets;
__device__ __forceinline__ uint32_t lpt_bucket(uint32_t count) { return min((uint32_t)kLptBuckets - 1u, (count - 1u) >> 4); }

// candidates: rays that hit the scene box and (when the occupied bounds are known) the bounds of the occupied cells —
// 21 % of an object-centred 800x800 view. They are compacted (warp-aggregated append, cursor in queue[0]) so that the
// thread-per-ray walk of the pre-pass runs on full warps; every other ray gets ray_count = 0 here.
__global__ void __launch_bounds__(256) k_ray_candidates(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                        const float* __restrict__ nears, const float* __restrict__ fars,
                                                        uint32_t N, const float* __restrict__ occ,
                                                        int32_t* __restrict__ cand, int32_t* __restrict__ ray_count,
                                                        unsigned int* __restrict__ queue) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
    bool keep = false;
    if (n < N) {
        ray_count[n] = 0;
        const float near = nears[n], far = fars[n];
        keep = near < far;
        if (keep && occ) {
            const float ox = rays_o[(size_t)n * 3], oy = rays_o[(size_t)n * 3 + 1], oz = rays_o[(size_t)n * 3 + 2];
            const float dx = rays_d[(size_t)n * 3], dy = rays_d[(size_t)n * 3 + 1], dz = rays_d[(size_t)n * 3 + 2];
            Marcher m;
            m.ox = ox; m.oy = oy; m.oz = oz; m.dx = dx; m.dy = dy; m.dz = dz;
            m.rdx = 1 / dx; m.rdy = 1 / dy; m.rdz = 1 / dz;
            keep = near < m.occupied_exit(occ);
        }
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, keep);
    if (mask) {
        uint32_t base = 0;
        const uint32_t leader = __ffs(mask) - 1;
        if (lane == leader) base = atomicAdd(queue, (unsigned int)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (keep) cand[base + __popc(mask & ((1u << lane) - 1u))] = (int32_t)n;
    }
}

__global__ void __launch_bounds__(128) k_ray_prepass(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                     const float* __restrict__ nears, const float* __restrict__ fars,
                                                     const float* __restrict__ noises, const uint8_t* __restrict__ bitfield,
                                                     uint32_t N, uint32_t C, uint32_t H, uint32_t max_steps, float bound,
                                                     float dt_gamma, int32_t* __restrict__ hit_list,
                                                     int32_t* __restrict__ ray_count, float* __restrict__ t_first,
                                                     float* __restrict__ t_last, unsigned int* __restrict__ queue,
                                                     const float* __restrict__ occ) {
    // thread i walks candidate i (k_ray_candidates; the list occupies the front of hit_list until k_lpt_scatter
    // overwrites it with the ordered hit list)
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    const bool work = i < min(N, queue[0]);
    const uint32_t n = work ? (uint32_t)hit_list[i] : 0u;
    bool hit = false;
    float tf = 0.f, tl = 0.f;
    uint32_t count = 0;
    if (work) {
        Marcher m;
        m.init(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, bound, dt_gamma, max_steps, C, H, bitfield);
        // The walk below is the reference's inference schedule with n_step = 1: after every sample the march
        // restarts from the compositor's ray parameter tc = tc + (t_end_of_sample - last_t) (raymarching.cu:1073,
        // 984-986), which is NOT always bit-identical to the marcher's own t (the subtraction rounds when the sample
        // lies beyond 2x the restart point). The persistent kernel performs the same walk, so t_first / t_last are
        // exact lattice points of it.
        // occ (optional): bounds of the occupied cells; behind the ray's exit from them no lattice point is occupied, so
        // the walk stops there (a ray that misses them — most of an object-centred view — does not walk at all)
        float far = fars[n];
        if (occ) far = fminf(far, m.occupied_exit(occ));
        float tc = nears[n];
        float t = tc;
        if (noises) t += m.step_size(t) * noises[n];
        float t_mark = t;
        float x, y, z, dt;
        while (t < far && count < max_steps) {
            if (m.probe(t, x, y, z, dt)) {
                if (count == 0) tf = t;
                tl = t;
                count++;
                t += dt;
                tc += t - t_mark;
                t = tc;
                t_mark = tc;
            }
        }
        hit = count > 0;
    }
    // longest-processing-time-first order: rays are bucketed by their sample count (16 samples per bucket) and the hit
    // list is written bucket by bucket, longest first. A lane of the persistent kernel keeps a ray until it ends, so with
    // pixel order the last rays pulled from the queue decide when a warp finishes (tile fill 0.80 at 800x800); with the
    // longest rays first the tail consists of the shortest ones.
    // (warp-aggregated global atomics, no CTA barrier: a warp retires as soon as its own rays are walked)
    if (work) {
        ray_count[n] = (int32_t)count;
        if (hit) {
            t_first[n] = tf;
            t_last[n] = tl;
        }
    }
    const uint32_t key = hit ? lpt_bucket(count) : (uint32_t)kLptBuckets + lane;     // misses: a key nobody shares
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (hit && lane == (uint32_t)__ffs(peers) - 1u) atomicAdd(queue + kQueueHist + key, (unsigned int)__popc(peers));
}

// one warp: descending exclusive scan of the bucket histogram -> per-bucket write cursors; total -> queue[2]
__global__ void k_lpt_offsets(unsigned int* __restrict__ queue) {
    const uint32_t lane = threadIdx.x;
    const unsigned int h = queue[kQueueHist + lane];
    unsigned int suffix = h;   // inclusive suffix sum over buckets >= lane
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_down_sync(0xffffffffu, suffix, o);
        if (lane + o < 32) suffix += v;
    }
    queue[kQueueCursor + lane] = suffix - h;      // rays in longer buckets come first
    if (lane == 0) {
        queue[2] = suffix;                        // number of rays with at least one sample
        queue[0] = 0u;                            // was the candidate cursor of the pre-pass; now the hit-list cursor
    }
}

__global__ void __launch_bounds__(256) k_lpt_scatter(const int32_t* __restrict__ ray_count, uint32_t N,
                                                     int32_t* __restrict__ hit_list, unsigned int* __restrict__ queue) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
    const uint32_t c = n < N ? (uint32_t)ray_count[n] : 0u;
    const uint32_t key = c ? lpt_bucket(c) : (kLptBuckets + lane);     // misses: a key nobody shares
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (c && lane == leader) base = atomicAdd(queue + kQueueCursor + key, (unsigned int)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (c) hit_list[base + __popc(peers & ((1u << lane) - 1u))] = (int32_t)n;
}

template <bool CLIP, bool AUX>
__global__ void __launch_bounds__(kFusedWarps * 32, 1) k_render_fused(RenderArgs a, pnerf_palette_field f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem* sm = reinterpret_cast<FusedSmem*>(smem_raw);
    uint2* wts = reinterpret_cast<uint2*>(smem_raw + ((sizeof(FusedSmem) + 15) & ~(size_t)15));
    WarpScratch* scratch = reinterpret_cast<WarpScratch*>(wts + (f.pred_clip ? kWUnitsClip : kWUnitsNoClip));
    fused_prologue(f, sm, wts);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpScratch& ws = scratch[wid];
    WarpAux* auxs = reinterpret_cast<WarpAux*>(scratch + kFusedWarps);
    float (*aux)[32] = AUX ? auxs[wid].acc : nullptr;
    // semantic-feature accumulators of the lane's ray: shared memory (written to clip_feat once, when the ray retires)
    WarpClip* clips = reinterpret_cast<WarpClip*>(reinterpret_cast<unsigned char*>(auxs) + (AUX ? sizeof(WarpAux) * kFusedWarps : 0));
    const bool clip_on = CLIP && a.clip_feat != nullptr;
    float (*cacc)[32] = clip_on ? clips[wid].acc : nullptr;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // retire a ray: its accumulators go to global memory once
    auto retire = [&](uint32_t ray_, float wsum_, float dep_, float r_, float g_, float b_) {
        if (CLIP && clip_on) {
            float* pc = a.clip_feat + (size_t)ray_ * f.clip_dim;
#pragma unroll
            for (int k = 0; k < kClipMax; k++)
                if (k < (int)f.clip_dim) pc[k] = cacc[k][lane];
        }
        a.weights_sum[ray_] = wsum_; a.depth[ray_] = dep_;
        a.image[(size_t)ray_ * 3] = r_; a.image[(size_t)ray_ * 3 + 1] = g_; a.image[(size_t)ray_ * 3 + 2] = b_;
        if (AUX) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                a.direct_rgb[(size_t)ray_ * 3 + c] = aux[c][lane];
                a.view_dep_rgb[(size_t)ray_ * 3 + c] = aux[3 + c][lane];
            }
#pragma unroll
            for (int k = 0; k < kNB; k++) a.basis_acc[(size_t)ray_ * kNB + k] = aux[6 + k][lane];
#pragma unroll
            for (int k = 0; k < kNB * 3; k++) {
                a.basis_rgb[(size_t)ray_ * kNB * 3 + k] = aux[6 + kNB + k][lane];
                a.unscaled_basis_rgb[(size_t)ray_ * kNB * 3 + k] = aux[6 + kNB + kNB * 3 + k][lane];
            }
        }
    };

    // per-lane ray state (the Marcher is rebuilt from rays_o/rays_d at each march call to keep registers free)
    bool has_ray = false;
    float ddx = 0.f, ddy = 0.f, ddz = 1.f;
    uint32_t ray = 0, count = 0;
    float tc = 0.f;        // the compositor's ray parameter (== rays_t of the reference)
    float t_cur = 0.f;     // march position: next lattice point to probe
    float t_mark = 0.f;    // "last_t" of the reference's march call (start of the current real-delta interval)
    float t_end = 0.f;     // lattice t of the ray's last occupied point (from the pre-pass)
    float wsum = 0.f, dep = 0.f, r = 0.f, g = 0.f, b = 0.f;
    bool exhausted = false;                    // warp-uniform: the hit list is used up
    uint32_t shaded = 0, rounds = 0;
    const uint32_t n_hit = a.queue[2];
    constexpr int kProbesPerRound = 6;         // bounds SIMT divergence: a lane crossing a gap resumes next round

    for (;;) {
        // ---- every lane finds its next sample; lanes without a ray pull one from the hit list ----
        bool sample = false;
        float x = 0.f, y = 0.f, z = 0.f, dt = 0.f, rdt = 0.f;
#pragma unroll 1
        for (int attempt = 0; attempt < 2; attempt++) {
            const uint32_t need = __ballot_sync(0xffffffffu, !has_ray);
            if (need && !exhausted) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(a.queue, (unsigned int)__popc(need));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base + __popc(need) >= n_hit) exhausted = true;
                if (!has_ray) {
                    const uint32_t slot = base + __popc(need & lt_mask);
                    if (slot < n_hit) {
                        ray = (uint32_t)a.hit_list[slot];
                        tc = a.nears[ray];
                        // first march call of the reference: t = rays_t (+ noise); last_t = t; the walk to the first
                        // occupied lattice point was done by the pre-pass
                        t_mark = tc;
                        if (a.noises) {
                            const float dt_min = 2 * 1.7320508075688772f / a.max_steps;
                            const float dt_max = 2 * 1.7320508075688772f * (1u << (a.C - 1)) / a.Hgrid;
                            t_mark += clampf(tc * a.dt_gamma, dt_min, dt_max) * a.noises[ray];
                        }
                        t_cur = a.t_first[ray];
                        t_end = a.t_last[ray];
                        has_ray = true; count = 0;
                        wsum = dep = r = g = b = 0.f;
                        if (AUX) {
#pragma unroll
                            for (int c = 0; c < kAuxCh; c++) aux[c][lane] = 0.f;
                        }
                        if (CLIP && clip_on) {
#pragma unroll
                            for (int k = 0; k < kClipMax; k++) cacc[k][lane] = 0.f;
                        }
                    }
                }
            }
            if (has_ray && !sample) {
                Marcher m;
                m.init(a.rays_o + (size_t)ray * 3, a.rays_d + (size_t)ray * 3, f.bound, a.dt_gamma, a.max_steps, a.C, a.Hgrid,
                       a.bitfield);
                ddx = m.dx; ddy = m.dy; ddz = m.dz;
                int probes = 0;
                while (t_cur <= t_end && probes < kProbesPerRound) {
                    probes++;
                    if (m.probe(t_cur, x, y, z, dt)) {
                        t_cur += dt;
                        rdt = t_cur - t_mark;
                        sample = true;
                        break;
                    }
                }
                if (!sample && !(t_cur <= t_end)) {   // no occupied lattice point left: retire the ray
                    retire(ray, wsum, dep, r, g, b);
                    has_ray = false;
                }
            }
            if (__ballot_sync(0xffffffffu, !has_ray) == 0u || exhausted) break;
        }
        const uint32_t smask = __ballot_sync(0xffffffffu, sample);
        if (smask == 0u) {
            if (exhausted && __ballot_sync(0xffffffffu, has_ray) == 0u) break;
            continue;
        }

        // ---- shade the warp's 32 samples ----
        rounds++;
        FieldOut o;
        eval_field<CLIP>(f, *sm, wts, ws, x, y, z, sample ? ddx : 0.f, sample ? ddy : 0.f, sample ? ddz : 1.f, sample, lane, o);

        // ---- composite (ref: raymarching.cu:1051-1110, one sample) ----
        if (sample) {
            shaded++;
            count++;
            float rgb[3], basis_rgb[kNB * 3], unscaled[kNB * 3];
            blend(f, *sm, o, rgb, basis_rgb, unscaled);
            const float alpha = 1.0f - __expf(-(f.density_scale * o.sigma) * dt);
            const float T = 1 - wsum;
            const float wgt = alpha * T;
            wsum += wgt;
            tc += rdt;
            dep += wgt * tc;
            r += wgt * rgb[0]; g += wgt * rgb[1]; b += wgt * rgb[2];
            // the next march call of the reference restarts from rays_t (= tc) with last_t = rays_t
            t_cur = tc;
            t_mark = tc;
            if (AUX) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    aux[c][lane] += wgt * (o.diffuse[c] + o.view_dep[c]);
                    aux[3 + c][lane] += wgt * o.view_dep[c];
                }
#pragma unroll
                for (int k = 0; k < kNB; k++) aux[6 + k][lane] += wgt * o.omega[k];
#pragma unroll
                for (int k = 0; k < kNB * 3; k++) {
                    aux[6 + kNB + k][lane] += wgt * basis_rgb[k];
                    aux[6 + kNB + kNB * 3 + k][lane] += wgt * unscaled[k];
                }
            }
            if (CLIP && clip_on) {
#pragma unroll
                for (int k = 0; k < kClipMax; k++) cacc[k][lane] += wgt * o.clip[k];      // o.clip is zero beyond clip_dim
            }
            // early termination (the terminating sample is accumulated, like the reference) or sample budget used up
            if (T < a.T_thresh || count >= a.max_steps) {
                retire(ray, wsum, dep, r, g, b);
                has_ray = false;
            }
        }
    }
    // sample statistics (one atomic per warp)
    shaded = (uint32_t)warp_sum((float)shaded);  // exact below 2^24 per warp
    if (lane == 0 && shaded) { atomicAdd(a.queue + 1, shaded); atomicAdd(a.queue + 3, rounds); }
}

// (A warp-per-ray variant of this kernel — lane = one of 32 consecutive samples of ONE ray, so that the corner loads of
// the coarse levels share cache lines — was drafted in round 1 and is in the history (commit 15cc783); it is not part
// of the build until it is finished and measured.)

}  // namespace pnerf

using namespace pnerf;

// measurement hook (off by default; not thread-safe, meant for the single-threaded benchmark)
static bool g_render_timing = false, g_render_timed = false;
static cudaEvent_t g_render_ev[2] = {nullptr, nullptr};

extern "C" {

/* fused field: xyzs, dirs [M,3] fp32 -> sigma [M], clip [M,clip_dim] (may be NULL), omega [M,4], off_rad [M,13],
 * view_dep [M,3], diffuse [M,3], all fp32. Replaces PaletteNetwork.forward (palette/network.py:156-185) in eval mode. */
int pnerf_palette_field_forward(const float* xyzs, const float* dirs, uint32_t M,
                                          const pnerf_palette_field* field, float* sigma, float* clip, float* omega,
                                          float* off_rad, float* view_dep, float* diffuse, void* stream) {
    if (M == 0) return PNERF_OK;
    PNERF_REQUIRE(xyzs && dirs && field && sigma && omega && off_rad && view_dep && diffuse);
    PNERF_REQUIRE(field->table_sigma && field->table_palette && field->offsets && field->wpack && field->head_bias &&
                  field->palette);
    if (field->L != 16 || field->clip_dim > (uint32_t)kClipMax) return PNERF_ERR_UNSUPPORTED;
    if (field->pred_clip && !field->table_clip) return PNERF_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const bool clip_on = field->pred_clip != 0;
    const size_t smem = fused_smem_bytes(clip_on, false);
    const uint32_t tiles = ceil_div(M, 32u);
    const uint32_t grid = min(ceil_div(tiles, (uint32_t)kFusedWarps), (uint32_t)kNumSMs);
    cudaError_t e;
    if (clip_on) {
        e = cudaFuncSetAttribute(k_field_forward<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_cuda_error(e, "field_forward attr"); return PNERF_ERR_CUDA; }
        k_field_forward<true><<<grid, kFusedWarps * 32, smem, s>>>(xyzs, dirs, M, *field, sigma, clip, omega, off_rad, view_dep, diffuse);
    } else {
        e = cudaFuncSetAttribute(k_field_forward<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_cuda_error(e, "field_forward attr"); return PNERF_ERR_CUDA; }
        k_field_forward<false><<<grid, kFusedWarps * 32, smem, s>>>(xyzs, dirs, M, *field, sigma, nullptr, omega, off_rad, view_dep, diffuse);
    }
    return check_launch("palette_field_forward");
}

/* persistent fused renderer: replaces the inference loop of PaletteRenderer.run_cuda (palette/renderer.py:430-523).
 * All outputs must be zero-initialised; queue[68] must be zero; hit_list is [2N], t_first/t_last are [N] scratch.
 * Aux maps may all be NULL (gui_mode). Launches a thread-per-ray pre-pass, the two small ordering kernels and the persistent kernel. */
int pnerf_palette_render_fused(const float* rays_o, const float* rays_d, const float* nears, const float* fars,
                                         const float* noises, const uint8_t* bitfield, uint32_t N, uint32_t C,
                                         uint32_t Hgrid, uint32_t max_steps, float dt_gamma, float T_thresh,
                                         const pnerf_palette_field* field, float* weights_sum, float* depth, float* image,
                                         float* direct_rgb, float* view_dep_rgb, float* basis_acc, float* basis_rgb,
                                         float* unscaled_basis_rgb, float* clip_feat, uint32_t* queue,
                                         int32_t* hit_list, float* t_first, float* t_last, const float* occ_aabb, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(rays_o && rays_d && nears && fars && bitfield && field && weights_sum && depth && image && queue);
    PNERF_REQUIRE(hit_list && t_first && t_last);
    PNERF_REQUIRE(field->table_sigma && field->table_palette && field->offsets && field->wpack && field->head_bias &&
                  field->palette);
    PNERF_REQUIRE(C >= 1 && C <= 16 && Hgrid >= 1 && max_steps >= 1);
    if (field->L != 16 || field->clip_dim > (uint32_t)kClipMax || Hgrid > 1024) return PNERF_ERR_UNSUPPORTED;
    if (field->pred_clip && !field->table_clip) return PNERF_ERR_INVALID_ARG;
    const bool aux = direct_rgb != nullptr;
    if (aux) PNERF_REQUIRE(view_dep_rgb && basis_acc && basis_rgb && unscaled_basis_rgb);
    RenderArgs a;
    a.rays_o = rays_o; a.rays_d = rays_d; a.nears = nears; a.fars = fars; a.noises = noises; a.bitfield = bitfield;
    a.N = N; a.C = C; a.Hgrid = Hgrid; a.max_steps = max_steps; a.dt_gamma = dt_gamma; a.T_thresh = T_thresh;
    a.weights_sum = weights_sum; a.depth = depth; a.image = image;
    a.direct_rgb = direct_rgb; a.view_dep_rgb = view_dep_rgb; a.basis_acc = basis_acc; a.basis_rgb = basis_rgb;
    a.unscaled_basis_rgb = unscaled_basis_rgb; a.clip_feat = clip_feat; a.queue = queue;
    a.hit_list = hit_list; a.t_first = t_first; a.t_last = t_last;
    cudaStream_t s = (cudaStream_t)stream;
    int32_t* ray_count = hit_list + N;       // second half of the scratch: samples per ray
    k_ray_candidates<<<ceil_div(N, 256u), 256, 0, s>>>(rays_o, rays_d, nears, fars, N, occ_aabb, hit_list, ray_count, queue);
    k_ray_prepass<<<ceil_div(N, 128u), 128, 0, s>>>(rays_o, rays_d, nears, fars, noises, bitfield, N, C, Hgrid, max_steps,
                                                    field->bound, dt_gamma, hit_list, ray_count, t_first, t_last, queue, occ_aabb);
    k_lpt_offsets<<<1, 32, 0, s>>>(queue);
    k_lpt_scatter<<<ceil_div(N, 256u), 256, 0, s>>>(ray_count, N, hit_list, queue);
    const bool clip_on = field->pred_clip != 0;
    const size_t smem = fused_smem_bytes(clip_on, aux, clip_on && clip_feat != nullptr);
    const uint32_t warps_needed = ceil_div(N, 32u);
    const uint32_t grid = min(ceil_div(warps_needed, (uint32_t)kFusedWarps), (uint32_t)kNumSMs);  // persistent: one CTA per SM
#define PNERF_LAUNCH_RENDER(CL, AX)                                                                                   \
    do {                                                                                                              \
        cudaError_t e = cudaFuncSetAttribute(k_render_fused<CL, AX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) { set_last_cuda_error(e, "render_fused attr"); return PNERF_ERR_CUDA; }                 \
        k_render_fused<CL, AX><<<grid, kFusedWarps * 32, smem, s>>>(a, *field);                                       \
    } while (0)
    // optional timing of the persistent kernel alone (bench.py's roofline): an event pair on the launch stream
    if (g_render_timing) {
        if (!g_render_ev[0]) { cudaEventCreate(&g_render_ev[0]); cudaEventCreate(&g_render_ev[1]); }
        cudaEventRecord(g_render_ev[0], s);
    }
    if (clip_on) { if (aux) PNERF_LAUNCH_RENDER(true, true); else PNERF_LAUNCH_RENDER(true, false); }
    else { if (aux) PNERF_LAUNCH_RENDER(false, true); else PNERF_LAUNCH_RENDER(false, false); }
#undef PNERF_LAUNCH_RENDER
    if (g_render_timing) { cudaEventRecord(g_render_ev[1], s); g_render_timed = true; }
    return check_launch("palette_render_fused");
}

void pnerf_render_kernel_timing(int enable) { g_render_timing = enable != 0; g_render_timed = false; }

float pnerf_render_kernel_last_ms(void) {
    if (!g_render_timed) return -1.0f;
    float ms = -1.0f;
    if (cudaEventSynchronize(g_render_ev[1]) != cudaSuccess || cudaEventElapsedTime(&ms, g_render_ev[0], g_render_ev[1]) != cudaSuccess)
        return -1.0f;
    return ms;
}

}  // extern "C"
