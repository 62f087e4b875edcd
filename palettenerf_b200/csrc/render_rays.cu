// render_rays.cu — the WARP-PER-RAY persistent palette renderer (round 2), for B200 (sm_100a).
//
// Replaces the inference loop of PaletteRenderer.run_cuda (palette/renderer.py:430-523: march n_step samples of every
// alive ray -> field -> blend -> 6+1 compositing launches -> compact the alive list, up to 1024 host iterations) by ONE
// persistent kernel in which a warp owns ONE ray at a time:
//
//   walk     the warp classifies 32 lattice points of its ray per step (warp_walk, march_common.cuh: the reference's
//            serial lattice, bit for bit) and leaves the ray parameters of the ray's samples in a per-warp scratch list;
//   shade    32 CONSECUTIVE samples of the ray form a tile (lane = sample): hash-grid gather with lane pairs
//            (gather_coop), the whole MLP chain on tensor cores, palette blend — eval_field<CLIP, COOP = true>;
//   blend    front-to-back compositing of the tile with a warp product scan of (1 - alpha); a ballot finds the sample at
//            which the ray terminates (T < T_thresh; that sample is accumulated, like raymarching.cu:1084-1098);
//   retire   per-lane partial sums are reduced once per ray and written to the output maps (one writer per ray).
//
// Why this shape (measured motivation, profiles/README.md): round 1's lane-per-ray kernel put 32 UNRELATED rays into
// every gather instruction — 32 different 128-byte lines per load, L1/TEX pipe 81 % busy, 0.064 of the tensor peak. Here
// the 32 lanes of a load hold 16 consecutive samples of one ray x the two x-neighbour corners: the coarse levels collapse
// to a few lines per instruction and the fine levels to ~17 instead of 32. It also removes the pre-pass (thread-per-ray
// walk, longest-first ordering, 4 launches): a ray is a few tiles, so a dynamic queue of rays balances by itself.
//
// Sample positions: the walk is the reference's march with the ray restarted nowhere (n_step -> infinity); the reference
// restarts every ray from its compositor parameter after n_step in [1, 8] samples, which perturbs the lattice in the last
// bit when a sample lies beyond twice the restart point (raymarching.cu:984-986, 1073). Both are the same lattice up to
// that rounding; the parity tests bound the effect (tests/test_golden_palette_gpu.py: <= 1e-3 on every map).
#include "fused_field.cuh"

namespace pnerf {

struct RaysArgs {
    const float* rays_o; const float* rays_d; const float* nears; const float* fars; const float* noises;  // noises may be NULL
    const uint8_t* bitfield;
    const float* occ;                                          // [6] bounds of the occupied cells or NULL
    uint32_t N, C, Hgrid, max_steps;
    float dt_gamma, T_thresh;
    float* weights_sum; float* depth; float* image;          // [N], [N], [N,3]   (zero-initialised by the caller)
    float* direct_rgb; float* view_dep_rgb; float* basis_acc; float* basis_rgb; float* unscaled_basis_rgb;  // aux (NULL in gui mode)
    float* clip_feat;                                         // [N, clip_dim] or NULL
    unsigned int* queue;                                      // [8]: ray cursor, samples shaded, rays with samples, tiles, candidates
    const int32_t* cand;                                      // [N] ids of the rays that can have samples
    float* t_scratch;                                         // [gridDim.x * kFusedWarps, max_steps]
};

enum { Q_CURSOR = 0, Q_SAMPLES = 1, Q_RAYS = 2, Q_TILES = 3, Q_CAND = 4 };

// per-warp partial sums of the auxiliary maps: [channel][lane], row stride 33 -> the per-ray column sums (lane = channel)
// are conflict-free
constexpr int kAccStride = 33;
struct RayAux { float acc[kAuxCh][kAccStride]; };
struct RayClip { float acc[kClipMax][kAccStride]; };

// candidates: rays that hit the scene box and (when known) the bounds of the occupied cells — 21 % of an object-centred
// 800x800 view. Warp-aggregated append keeps pixel order within a warp, so neighbouring rays stay neighbours in the queue.
__global__ void __launch_bounds__(256) k_rays_candidates(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
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
        if (lane == leader) base = atomicAdd(queue + Q_CAND, (unsigned int)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (keep) cand[base + __popc(mask & ((1u << lane) - 1u))] = (int32_t)n;
    }
}

__host__ __device__ constexpr size_t rays_smem_bytes(bool clip, bool aux, bool clip_acc) {
    return ((sizeof(FusedSmem) + 15) & ~(size_t)15) + (size_t)(clip ? kWUnitsClip : kWUnitsNoClip) * sizeof(uint2) +
           sizeof(WarpScratch) * kFusedWarps + (aux ? sizeof(RayAux) * kFusedWarps : 0) +
           (clip_acc ? sizeof(RayClip) * kFusedWarps : 0) + 16;
}

template <bool CLIP, bool AUX>
__global__ void __launch_bounds__(kFusedWarps * 32, 1) k_render_rays(RaysArgs a, pnerf_palette_field f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem* sm = reinterpret_cast<FusedSmem*>(smem_raw);
    uint2* wts = reinterpret_cast<uint2*>(smem_raw + ((sizeof(FusedSmem) + 15) & ~(size_t)15));
    WarpScratch* scratch = reinterpret_cast<WarpScratch*>(wts + (f.pred_clip ? kWUnitsClip : kWUnitsNoClip));
    fused_prologue(f, sm, wts);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpScratch& ws = scratch[wid];
    RayAux* auxs = reinterpret_cast<RayAux*>(scratch + kFusedWarps);
    float (*aux)[kAccStride] = AUX ? auxs[wid].acc : nullptr;
    RayClip* clips = reinterpret_cast<RayClip*>(reinterpret_cast<unsigned char*>(auxs) + (AUX ? sizeof(RayAux) * kFusedWarps : 0));
    const bool clip_on = CLIP && a.clip_feat != nullptr;
    float (*cacc)[kAccStride] = clip_on ? clips[wid].acc : nullptr;
    float* const t_list = a.t_scratch + (size_t)(blockIdx.x * kFusedWarps + wid) * a.max_steps;
    const uint32_t n_cand = a.queue[Q_CAND];
    uint32_t shaded = 0, tiles = 0, hit_rays = 0;

    for (;;) {
        // ---- next ray of this warp ----
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(a.queue + Q_CURSOR, 1u);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= n_cand) break;
        const uint32_t ray = (uint32_t)a.cand[slot];

        // ---- walk: ray parameters of all its samples -> t_list (the reference's lattice, bit for bit) ----
        uint32_t count;
        float t0;
        {
            Marcher m;
            m.init(a.rays_o + (size_t)ray * 3, a.rays_d + (size_t)ray * 3, f.bound, a.dt_gamma, a.max_steps, a.C, a.Hgrid,
                   a.bitfield);
            float far = a.fars[ray];
            if (a.occ) far = fminf(far, m.occupied_exit(a.occ));
            t0 = m.first_t(a.nears[ray], a.noises ? a.noises[ray] : 0.f);
            count = warp_walk<false>(m, t0, far, a.max_steps, (uint32_t)lane, nullptr, nullptr, nullptr, t_list);
        }
        if (count == 0) continue;
        __syncwarp();
        hit_rays++;

        // ---- per-ray state ----
        const float ox = a.rays_o[(size_t)ray * 3], oy = a.rays_o[(size_t)ray * 3 + 1], oz = a.rays_o[(size_t)ray * 3 + 2];
        const float dx = a.rays_d[(size_t)ray * 3], dy = a.rays_d[(size_t)ray * 3 + 1], dz = a.rays_d[(size_t)ray * 3 + 2];
        const float dt_min = 2 * 1.7320508075688772f / a.max_steps;
        const float dt_max = 2 * 1.7320508075688772f * (1u << (a.C - 1)) / a.Hgrid;
        float T_run = 1.f;                               // transmittance in front of the current tile (warp-uniform)
        float wsum = 0.f, dep = 0.f, r = 0.f, g = 0.f, b = 0.f;   // per-lane partial sums
        if (AUX) {
#pragma unroll
            for (int c = 0; c < kAuxCh; c++) aux[c][lane] = 0.f;
        }
        if (CLIP && clip_on) {
#pragma unroll
            for (int k = 0; k < kClipMax; k++) cacc[k][lane] = 0.f;
        }

        // ---- tiles of 32 consecutive samples ----
#pragma unroll 1
        for (uint32_t base = 0; base < count; base += 32) {
            const uint32_t k = base + (uint32_t)lane;
            const bool active = k < count;
            const float t = active ? t_list[k] : t0;
            const float x = clampf(ox + t * dx, -f.bound, f.bound);
            const float y = clampf(oy + t * dy, -f.bound, f.bound);
            const float z = clampf(oz + t * dz, -f.bound, f.bound);
            const float dt = clampf(t * a.dt_gamma, dt_min, dt_max);
            const float t_end = t + dt;                  // == the compositor's ray parameter after this sample
            tiles++;
            FieldOut o;
            eval_field<CLIP, true>(f, *sm, wts, ws, x, y, z, active ? dx : 0.f, active ? dy : 0.f, active ? dz : 1.f, active, lane, o);

            // front-to-back compositing of the tile (ref: raymarching.cu:1051-1110 per sample)
            const float alpha = active ? 1.0f - __expf(-(f.density_scale * o.sigma) * dt) : 0.f;
            float incl = 1.0f - alpha;                   // inclusive product scan of (1 - alpha)
#pragma unroll
            for (int ofs = 1; ofs < 32; ofs <<= 1) {
                const float up = __shfl_up_sync(0xffffffffu, incl, ofs);
                if (lane >= ofs) incl *= up;
            }
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.f;
            const float T = T_run * excl;                // transmittance in front of this sample
            // the sample whose T falls below the threshold is still accumulated; everything behind it is not
            const uint32_t term = __ballot_sync(0xffffffffu, active && T < a.T_thresh);
            const int last = term ? (__ffs(term) - 1) : 31;
            const bool use = active && lane <= last;
            const float wgt = use ? alpha * T : 0.f;
            shaded += __popc(__ballot_sync(0xffffffffu, use));
            float rgb[3], basis_rgb[kNB * 3], unscaled[kNB * 3];
            blend(f, *sm, o, rgb, basis_rgb, unscaled);
            wsum += wgt;
            dep += wgt * t_end;
            r += wgt * rgb[0]; g += wgt * rgb[1]; b += wgt * rgb[2];
            if (AUX) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    aux[c][lane] += wgt * (o.diffuse[c] + o.view_dep[c]);
                    aux[3 + c][lane] += wgt * o.view_dep[c];
                }
#pragma unroll
                for (int q = 0; q < kNB; q++) aux[6 + q][lane] += wgt * o.omega[q];
#pragma unroll
                for (int q = 0; q < kNB * 3; q++) {
                    aux[6 + kNB + q][lane] += wgt * basis_rgb[q];
                    aux[6 + kNB + kNB * 3 + q][lane] += wgt * unscaled[q];
                }
            }
            if (CLIP && clip_on) {
#pragma unroll
                for (int q = 0; q < kClipMax; q++) cacc[q][lane] += wgt * o.clip[q];      // o.clip is zero beyond clip_dim
            }
            if (term) break;
            T_run *= __shfl_sync(0xffffffffu, incl, 31);
        }

        // ---- retire: reduce the per-lane partial sums, one writer per value ----
        wsum = warp_sum(wsum); dep = warp_sum(dep); r = warp_sum(r); g = warp_sum(g); b = warp_sum(b);
        if (lane == 0) {
            a.weights_sum[ray] = wsum; a.depth[ray] = dep;
            a.image[(size_t)ray * 3] = r; a.image[(size_t)ray * 3 + 1] = g; a.image[(size_t)ray * 3 + 2] = b;
        }
        __syncwarp();
        if (AUX) {
#pragma unroll 1
            for (int c = lane; c < kAuxCh; c += 32) {
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < 32; q++) s += aux[c][q];
                float* dst;
                if (c < 3) dst = a.direct_rgb + (size_t)ray * 3 + c;
                else if (c < 6) dst = a.view_dep_rgb + (size_t)ray * 3 + (c - 3);
                else if (c < 6 + kNB) dst = a.basis_acc + (size_t)ray * kNB + (c - 6);
                else if (c < 6 + kNB + kNB * 3) dst = a.basis_rgb + (size_t)ray * kNB * 3 + (c - 6 - kNB);
                else dst = a.unscaled_basis_rgb + (size_t)ray * kNB * 3 + (c - 6 - kNB - kNB * 3);
                *dst = s;
            }
        }
        if (CLIP && clip_on) {
            if (lane < (int)f.clip_dim) {
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < 32; q++) s += cacc[lane][q];
                a.clip_feat[(size_t)ray * f.clip_dim + lane] = s;
            }
        }
        __syncwarp();
    }
    // statistics (one atomic per warp and counter)
    if (lane == 0 && tiles) {
        atomicAdd(a.queue + Q_SAMPLES, shaded);
        atomicAdd(a.queue + Q_RAYS, hit_rays);
        atomicAdd(a.queue + Q_TILES, tiles);
    }
}

}  // namespace pnerf

using namespace pnerf;

static bool g_rays_timing = false, g_rays_timed = false;
static cudaEvent_t g_rays_ev[2] = {nullptr, nullptr};

extern "C" {

uint32_t pnerf_palette_render_rays_warps(void) {
    int dev = 0, sms = kNumSMs;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (uint32_t)sms * (uint32_t)kFusedWarps;
}

/* Warp-per-ray persistent renderer: replaces the inference loop of PaletteRenderer.run_cuda (palette/renderer.py:430-523).
 * Outputs and queue[8] must be zero-initialised; cand is [N] int32 scratch, t_scratch [pnerf_palette_render_rays_warps() *
 * max_steps] fp32 scratch. Aux maps may all be NULL (gui_mode); occ_aabb [6] from pnerf_occupied_bounds or NULL.
 * On return (stream order) queue = {ray cursor, samples shaded, rays with samples, 32-sample tiles, candidate rays}. */
int pnerf_palette_render_rays(const float* rays_o, const float* rays_d, const float* nears, const float* fars,
                              const float* noises, const uint8_t* bitfield, uint32_t N, uint32_t C, uint32_t Hgrid,
                              uint32_t max_steps, float dt_gamma, float T_thresh, const pnerf_palette_field* field,
                              float* weights_sum, float* depth, float* image, float* direct_rgb, float* view_dep_rgb,
                              float* basis_acc, float* basis_rgb, float* unscaled_basis_rgb, float* clip_feat, uint32_t* queue,
                              int32_t* cand, float* t_scratch, const float* occ_aabb, void* stream) {
    if (N == 0) return PNERF_OK;
    PNERF_REQUIRE(rays_o && rays_d && nears && fars && bitfield && field && weights_sum && depth && image && queue);
    PNERF_REQUIRE(cand && t_scratch);
    PNERF_REQUIRE(field->table_sigma_palette && field->offsets && field->wpack && field->head_bias && field->palette);
    PNERF_REQUIRE(C >= 1 && C <= 16 && Hgrid >= 1 && max_steps >= 1);
    if (field->L != 16 || field->clip_dim > (uint32_t)kClipMax || Hgrid > 1024) return PNERF_ERR_UNSUPPORTED;
    if (field->pred_clip && !field->table_clip) return PNERF_ERR_INVALID_ARG;
    const bool aux = direct_rgb != nullptr;
    if (aux) PNERF_REQUIRE(view_dep_rgb && basis_acc && basis_rgb && unscaled_basis_rgb);
    RaysArgs a;
    a.rays_o = rays_o; a.rays_d = rays_d; a.nears = nears; a.fars = fars; a.noises = noises; a.bitfield = bitfield;
    a.occ = occ_aabb;
    a.N = N; a.C = C; a.Hgrid = Hgrid; a.max_steps = max_steps; a.dt_gamma = dt_gamma; a.T_thresh = T_thresh;
    a.weights_sum = weights_sum; a.depth = depth; a.image = image;
    a.direct_rgb = direct_rgb; a.view_dep_rgb = view_dep_rgb; a.basis_acc = basis_acc; a.basis_rgb = basis_rgb;
    a.unscaled_basis_rgb = unscaled_basis_rgb; a.clip_feat = clip_feat; a.queue = queue; a.cand = cand; a.t_scratch = t_scratch;
    cudaStream_t s = (cudaStream_t)stream;
    k_rays_candidates<<<ceil_div(N, 256u), 256, 0, s>>>(rays_o, rays_d, nears, fars, N, occ_aabb, cand, queue);
    const bool clip_on = field->pred_clip != 0;
    const bool clip_acc = clip_on && clip_feat != nullptr;
    const size_t smem = rays_smem_bytes(clip_on, aux, clip_acc);
    const uint32_t grid = pnerf_palette_render_rays_warps() / (uint32_t)kFusedWarps;      // persistent: one CTA per SM
    static bool attr_done[2][2] = {{false, false}, {false, false}};
#define PNERF_LAUNCH_RAYS(CL, AX)                                                                                        \
    do {                                                                                                                \
        if (!attr_done[CL][AX]) {                                                                                       \
            cudaError_t e = cudaFuncSetAttribute(k_render_rays<CL, AX>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                                 (int)rays_smem_bytes(CL, AX, CL));                                     \
            if (e != cudaSuccess) { set_last_cuda_error(e, "render_rays attr"); return PNERF_ERR_CUDA; }                \
            attr_done[CL][AX] = true;                                                                                   \
        }                                                                                                               \
        k_render_rays<CL, AX><<<grid, kFusedWarps * 32, smem, s>>>(a, *field);                                          \
    } while (0)
    if (g_rays_timing) {
        if (!g_rays_ev[0]) { cudaEventCreate(&g_rays_ev[0]); cudaEventCreate(&g_rays_ev[1]); }
        cudaEventRecord(g_rays_ev[0], s);
    }
    if (clip_on) { if (aux) PNERF_LAUNCH_RAYS(true, true); else PNERF_LAUNCH_RAYS(true, false); }
    else { if (aux) PNERF_LAUNCH_RAYS(false, true); else PNERF_LAUNCH_RAYS(false, false); }
#undef PNERF_LAUNCH_RAYS
    if (g_rays_timing) { cudaEventRecord(g_rays_ev[1], s); g_rays_timed = true; }
    return check_launch("palette_render_rays");
}

/* measurement hook of bench.py's roofline (off by default; single-threaded use): an event pair around the persistent
 * kernel of the LAST pnerf_palette_render_rays call */
void pnerf_render_rays_timing(int enable) { g_rays_timing = enable != 0; g_rays_timed = false; }

float pnerf_render_rays_last_ms(void) {
    if (!g_rays_timed) return -1.0f;
    float ms = -1.0f;
    if (cudaEventSynchronize(g_rays_ev[1]) != cudaSuccess || cudaEventElapsedTime(&ms, g_rays_ev[0], g_rays_ev[1]) != cudaSuccess)
        return -1.0f;
    return ms;
}

}  // extern "C"
