// density_tc.cu — density-grid refresh (NeRFRenderer.update_extra_state, nerf/renderer.py:467-561) as kernels.
//
// The reference builds the cell coordinates with meshgrid / randint / nonzero / cat, evaluates density() in chunks, scatters
// into a temporary grid, and reads two scalars back to the host. Here:
//   pnerf_density_occupied_list   (partial refresh only) ascending list of the cells with density > 0 per cascade (count per
//                                 chunk, scan, ordered write: == torch.nonzero order), the count stays on the device;
//   pnerf_density_grid_sweep      ONE kernel: cell selection (full sweep, or N uniform + N occupied cells per cascade), jitter
//                                 from a counter-based generator, hash-grid gather + sigma net on tcgen05 (eval_field_tc<
//                                 TC_DENSITY>), atomicMax into the temporary grid (duplicates: the reference keeps "one of
//                                 the written values" by a write race; the maximum is one of them, deterministically);
//   pnerf_density_grid_finalize   EMA-max of the grid + per-block partial sums of clamp(density, 0);
//   pnerf_packbits_mean           every CTA folds the partial sums in the same fixed order -> mean density ->
//                                 threshold = min(mean, density_thresh) -> packbits. No host round trip.
// Data parallel: rank r takes the tiles t with t % world == r (same seed on all ranks), the temporary grids are merged with
// one all-reduce(max) before the finalize step (palettenerf_b200/distributed.py::merge_density).
#include "field_tc.cuh"

namespace pnerf {

constexpr int kDenGroups = 4;
constexpr int kDenThreads = kDenGroups * 128;
constexpr int kDenSharedBytes = (sizeof(TcShared) + 1023) & ~1023;

struct DensityArgs {
    float* tmp_grid;                 // [C, H^3], initialised to -1 by the caller
    uint32_t C, H;
    float bound, density_scale;
    uint32_t partial;                // 0: every cell once; 1: n_random uniform cells + n_random occupied cells per cascade
    uint32_t n_random;
    const int32_t* occ_list;         // [C, H^3] (partial)
    const uint32_t* occ_count;       // [C]
    uint64_t seed;
    uint32_t rank, world;
    const float* jitter;             // optional [C * points_per_cascade, 3] U[0,1) (tests); NULL = counter-based generator
};

__device__ __forceinline__ uint32_t mix32(uint64_t k) {      // splitmix64 finaliser, upper half
    k += 0x9E3779B97F4A7C15ull;
    k = (k ^ (k >> 30)) * 0xBF58476D1CE4E5B9ull;
    k = (k ^ (k >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((k ^ (k >> 31)) >> 32);
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

__global__ void __launch_bounds__(kDenThreads, 1) k_density_grid_sweep(DensityArgs a, pnerf_palette_field f) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcShared* sm = reinterpret_cast<TcShared*>(smem_raw);
    unsigned char* wts = smem_raw + kDenSharedBytes;
    unsigned char* groups = wts + kTcWBytesNoClip;
    constexpr int group_bytes = kTcGroupBytesNoClip;
    tc_prologue<kDenGroups>(f, f.wpack_tc, sm, wts, groups, group_bytes);
    TcGroup g = tc_make_group(sm, wts, groups, group_bytes);
    const int lane = threadIdx.x & 31, gi = threadIdx.x >> 7;
    const uint32_t H = a.H, H3 = H * H * H;
    const uint32_t per_cas = a.partial ? 2u * a.n_random : H3;
    const uint32_t tiles_per_cas = ceil_div(per_cas, 128u);
    const uint32_t n_tiles = tiles_per_cas * a.C;
    for (uint32_t tile = (blockIdx.x * kDenGroups + gi) * a.world + a.rank; tile < n_tiles; tile += gridDim.x * kDenGroups * a.world) {
        const uint32_t cas = tile / tiles_per_cas;
        const uint32_t p = (tile % tiles_per_cas) * 128 + (uint32_t)g.row;
        bool active = p < per_cas;
        uint32_t cx = 0, cy = 0, cz = 0;
        const uint64_t key = a.seed * 0x100000001B3ull + ((uint64_t)cas << 40) + p;
        if (active) {
            if (!a.partial) {                   // x-major order of the reference's meshgrid (nerf/renderer.py:483-491)
                cx = p / (H * H); cy = (p / H) % H; cz = p % H;
            } else if (p < a.n_random) {        // uniform random cells (:515-517)
                cx = mix32(key * 4 + 1) % H; cy = mix32(key * 4 + 2) % H; cz = mix32(key * 4 + 3) % H;
            } else {                            // random occupied cells (:518-523)
                const uint32_t cnt = a.occ_count[cas];
                if (cnt == 0) active = false;
                else {
                    const uint32_t m = (uint32_t)a.occ_list[(size_t)cas * H3 + mix32(key * 4 + 1) % cnt];
                    cx = compact3(m); cy = compact3(m >> 1); cz = compact3(m >> 2);
                }
            }
        }
        const float b = fminf((float)(1u << cas), a.bound), half = b / (float)H;     // (:494-496)
        float jx, jy, jz;
        if (a.jitter) {
            const size_t q = ((size_t)cas * per_cas + p) * 3;
            jx = active ? a.jitter[q] : 0.f; jy = active ? a.jitter[q + 1] : 0.f; jz = active ? a.jitter[q + 2] : 0.f;
        } else {
            jx = u01(mix32(key * 4 + 0x51)); jy = u01(mix32(key * 4 + 0x52)); jz = u01(mix32(key * 4 + 0x53));
        }
        const float sc = b - half;
        const float x = (2.f * (float)cx / (float)(H - 1) - 1.f) * sc + (jx * 2.f - 1.f) * half;
        const float y = (2.f * (float)cy / (float)(H - 1) - 1.f) * sc + (jy * 2.f - 1.f) * half;
        const float z = (2.f * (float)cz / (float)(H - 1) - 1.f) * sc + (jz * 2.f - 1.f) * half;
        FieldOut o;
        eval_field_tc<TC_DENSITY>(f, *sm, g, x, y, z, 0.f, 0.f, 1.f, active, lane, o);
        if (active) {
            const float sig = o.sigma * a.density_scale;           // > 0: its bit pattern orders like the value
            atomicMax(reinterpret_cast<int*>(a.tmp_grid) + (size_t)cas * H3 + morton_encode(cx, cy, cz), __float_as_int(sig));
        }
    }
    tc_epilogue_cta<kDenGroups>(sm);
}

// batch density: points [M,3] -> sigma [M] (PaletteNetwork.density / NeRFNetwork.density in eval mode)
__global__ void __launch_bounds__(kDenThreads, 1) k_density_tc(const float* __restrict__ xyzs, uint32_t M, pnerf_palette_field f,
                                                               float* __restrict__ sigma) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcShared* sm = reinterpret_cast<TcShared*>(smem_raw);
    unsigned char* wts = smem_raw + kDenSharedBytes;
    unsigned char* groups = wts + kTcWBytesNoClip;
    constexpr int group_bytes = kTcGroupBytesNoClip;
    tc_prologue<kDenGroups>(f, f.wpack_tc, sm, wts, groups, group_bytes);
    TcGroup g = tc_make_group(sm, wts, groups, group_bytes);
    const int lane = threadIdx.x & 31, gi = threadIdx.x >> 7;
    const uint32_t n_tiles = ceil_div(M, 128u);
    for (uint32_t tile = blockIdx.x * kDenGroups + gi; tile < n_tiles; tile += gridDim.x * kDenGroups) {
        const uint32_t s = tile * 128 + (uint32_t)g.row;
        const bool active = s < M;
        const float x = active ? xyzs[(size_t)s * 3] : 0.f, y = active ? xyzs[(size_t)s * 3 + 1] : 0.f,
                    z = active ? xyzs[(size_t)s * 3 + 2] : 0.f;
        FieldOut o;
        eval_field_tc<TC_DENSITY>(f, *sm, g, x, y, z, 0.f, 0.f, 1.f, active, lane, o);
        if (active) sigma[s] = o.sigma;
    }
    tc_epilogue_cta<kDenGroups>(sm);
}

// Ascending list (torch.nonzero order) of the cells with density > 0 per cascade, in three small launches: per-chunk counts,
// an exclusive scan of the chunk counts (one CTA per cascade), ordered write. A chunk = 1024 threads x 16 consecutive cells.
constexpr uint32_t kOccPerThread = 16, kOccChunk = 1024 * kOccPerThread;

__device__ __forceinline__ uint32_t occ_thread_mask(const float* __restrict__ gcas, uint32_t first, uint32_t H3) {
    uint32_t m = 0;
#pragma unroll
    for (uint32_t k = 0; k < kOccPerThread; k++) {
        const uint32_t i = first + k;
        if (i < H3 && gcas[i] > 0.f) m |= 1u << k;
    }
    return m;
}

// block-wide exclusive scan of one value per thread (1024 threads); returns the exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_tot, uint32_t* total) {
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += up;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = warp_tot[lane];
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (uint32_t)o) winc += up;
        }
        warp_tot[lane] = winc - w;
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    return warp_tot[wid] + inc - v;
}

__global__ void __launch_bounds__(1024) k_occ_count(const float* __restrict__ grid, uint32_t H3, uint32_t* __restrict__ chunk_counts) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t total;
    const uint32_t cas = blockIdx.y, first = blockIdx.x * kOccChunk + threadIdx.x * kOccPerThread;
    const uint32_t c = __popc(occ_thread_mask(grid + (size_t)cas * H3, first, H3));
    block_exclusive_scan(c, warp_tot, &total);
    if (threadIdx.x == 0) chunk_counts[cas * gridDim.x + blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_occ_scan(uint32_t* __restrict__ chunk_counts, uint32_t n_chunks, uint32_t* __restrict__ occ_count) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t total;
    const uint32_t cas = blockIdx.x;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_chunks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_chunks ? chunk_counts[cas * n_chunks + i] : 0u;
        const uint32_t ex = block_exclusive_scan(v, warp_tot, &total);
        if (i < n_chunks) chunk_counts[cas * n_chunks + i] = carry + ex;        // counts -> offsets, in place
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) occ_count[cas] = carry;
}

__global__ void __launch_bounds__(1024) k_occ_write(const float* __restrict__ grid, uint32_t H3, const uint32_t* __restrict__ chunk_offsets,
                                                    int32_t* __restrict__ occ_list) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t total;
    const uint32_t cas = blockIdx.y, first = blockIdx.x * kOccChunk + threadIdx.x * kOccPerThread;
    uint32_t m = occ_thread_mask(grid + (size_t)cas * H3, first, H3);
    uint32_t pos = chunk_offsets[cas * gridDim.x + blockIdx.x] + block_exclusive_scan(__popc(m), warp_tot, &total);
    int32_t* out = occ_list + (size_t)cas * H3;
    while (m) {
        const uint32_t k = __ffs(m) - 1;
        out[pos++] = (int32_t)(first + k);
        m &= m - 1;
    }
}

constexpr int kFinThreads = 256, kFinPerThread = 16;
// EMA-max (nerf/renderer.py:543-545) + per-CTA partial sums of clamp(density, 0) in a fixed order
__global__ void __launch_bounds__(kFinThreads) k_density_finalize(float* __restrict__ grid, float* __restrict__ tmp, uint64_t n,
                                                                  float decay, float* __restrict__ partials) {
    __shared__ float red[kFinThreads];
    const uint64_t base = (uint64_t)blockIdx.x * kFinThreads * kFinPerThread;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kFinPerThread; k++) {
        const uint64_t i = base + (uint64_t)k * kFinThreads + threadIdx.x;
        if (i < n) {
            float old = grid[i];
            const float t = tmp[i];
            if (old >= 0.f && t >= 0.f) { old = fmaxf(old * decay, t); grid[i] = old; }
            tmp[i] = -1.f;                                  // ready for the next refresh
            s += fmaxf(old, 0.f);
        }
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = kFinThreads / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
}

// every CTA: mean = sum(partials) / n in one fixed order (identical in all CTAs), thresh = min(mean, density_thresh);
// then packbits of its slice (raymarching.cu:271-303). stats[0] = mean, stats[1] = threshold
__global__ void __launch_bounds__(256) k_packbits_mean(const float* __restrict__ grid, uint32_t n_bytes, const float* __restrict__ partials,
                                                       uint32_t n_partials, uint64_t n_cells, float density_thresh,
                                                       uint8_t* __restrict__ bitfield, float* __restrict__ stats) {
    __shared__ float red[256];
    float s = 0.f;
    for (uint32_t i = threadIdx.x; i < n_partials; i += 256) s += partials[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    const float mean = red[0] / (float)n_cells;
    const float thresh = fminf(mean, density_thresh);
    if (blockIdx.x == 0 && threadIdx.x == 0) { stats[0] = mean; stats[1] = thresh; }
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bytes) return;
    const float4 lo = reinterpret_cast<const float4*>(grid)[(size_t)b * 2], hi = reinterpret_cast<const float4*>(grid)[(size_t)b * 2 + 1];
    uint32_t bits = 0;
    bits |= (lo.x > thresh) ? 1u : 0u; bits |= (lo.y > thresh) ? 2u : 0u; bits |= (lo.z > thresh) ? 4u : 0u; bits |= (lo.w > thresh) ? 8u : 0u;
    bits |= (hi.x > thresh) ? 16u : 0u; bits |= (hi.y > thresh) ? 32u : 0u; bits |= (hi.z > thresh) ? 64u : 0u; bits |= (hi.w > thresh) ? 128u : 0u;
    bitfield[b] = (uint8_t)bits;
}

}  // namespace pnerf

using namespace pnerf;

static int den_sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            sms = kNumSMs;
    }
    return sms;
}

static int density_field_ok(const pnerf_palette_field* field) {
    if (!field || !field->table_sigma || !field->offsets || !field->wpack_tc || !field->head_bias || !field->palette)
        return PNERF_ERR_INVALID_ARG;
    if (field->L != 16) return PNERF_ERR_UNSUPPORTED;
    return PNERF_OK;
}

static size_t den_smem() { return (size_t)kDenSharedBytes + kTcWBytesNoClip + (size_t)kDenGroups * kTcGroupBytesNoClip; }

extern "C" {

/* sigma [M] = density(xyzs [M,3]) on the tensor-core field (NOT scaled by density_scale). field->table_sigma must be the
 * fp16 [n,2] table of the density grid. Replaces {NeRF,Palette}Network.density in eval mode (nerf/network.py:126-156). */
int pnerf_density_tc(const float* xyzs, uint32_t M, const pnerf_palette_field* field, float* sigma, void* stream) {
    if (M == 0) return PNERF_OK;
    PNERF_REQUIRE(xyzs && sigma);
    if (int st = density_field_ok(field)) return st;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_density_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)den_smem());
        if (e != cudaSuccess) { set_last_cuda_error(e, "density_tc attr"); return PNERF_ERR_CUDA; }
        attr = true;
    }
    const uint32_t grid = min(ceil_div(ceil_div(M, 128u), (uint32_t)kDenGroups), (uint32_t)den_sm_count());
    k_density_tc<<<grid, kDenThreads, den_smem(), (cudaStream_t)stream>>>(xyzs, M, *field, sigma);
    return check_launch("density_tc");
}

uint32_t pnerf_density_occupied_chunks(uint32_t H) { return ceil_div(H * H * H, kOccChunk); }

/* ascending list of the cells with density_grid > 0 of every cascade: occ_list [C, H^3] int32, occ_count [C] u32;
 * chunk_scratch [C * pnerf_density_occupied_chunks(H)] u32 */
int pnerf_density_occupied_list(const float* density_grid, uint32_t C, uint32_t H, int32_t* occ_list, uint32_t* occ_count,
                                uint32_t* chunk_scratch, void* stream) {
    PNERF_REQUIRE(density_grid && occ_list && occ_count && chunk_scratch && C >= 1 && C <= 16 && H >= 2 && H <= 1024);
    const uint32_t H3 = H * H * H, nb = pnerf_density_occupied_chunks(H);
    cudaStream_t s = (cudaStream_t)stream;
    k_occ_count<<<dim3(nb, C), 1024, 0, s>>>(density_grid, H3, chunk_scratch);
    k_occ_scan<<<C, 1024, 0, s>>>(chunk_scratch, nb, occ_count);
    k_occ_write<<<dim3(nb, C), 1024, 0, s>>>(density_grid, H3, chunk_scratch, occ_list);
    return check_launch("density_occupied_list");
}

/* density-grid sweep: tmp_grid [C, H^3] (pre-filled with -1) receives density * density_scale of one jittered point per
 * selected cell (max over duplicates). partial = 0: every cell; partial = 1: n_random uniform + n_random occupied cells per
 * cascade (occ_list / occ_count from pnerf_density_occupied_list). rank / world: this rank evaluates every world-th tile.
 * jitter: optional explicit U[0,1) numbers [C * points_per_cascade, 3] (tests). Replaces nerf/renderer.py:476-541. */
int pnerf_density_grid_sweep(float* tmp_grid, uint32_t C, uint32_t H, float bound, float density_scale, uint32_t partial,
                             uint32_t n_random, const int32_t* occ_list, const uint32_t* occ_count, uint64_t seed,
                             uint32_t rank, uint32_t world, const float* jitter, const pnerf_palette_field* field, void* stream) {
    PNERF_REQUIRE(tmp_grid && C >= 1 && C <= 16 && H >= 2 && H <= 1024 && world >= 1 && rank < world);
    if (partial) PNERF_REQUIRE(occ_list && occ_count && n_random >= 1);
    if (int st = density_field_ok(field)) return st;
    DensityArgs a;
    a.tmp_grid = tmp_grid; a.C = C; a.H = H; a.bound = bound; a.density_scale = density_scale; a.partial = partial;
    a.n_random = n_random; a.occ_list = occ_list; a.occ_count = occ_count; a.seed = seed; a.rank = rank; a.world = world;
    a.jitter = jitter;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_density_grid_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)den_smem());
        if (e != cudaSuccess) { set_last_cuda_error(e, "density_grid_sweep attr"); return PNERF_ERR_CUDA; }
        attr = true;
    }
    k_density_grid_sweep<<<(uint32_t)den_sm_count(), kDenThreads, den_smem(), (cudaStream_t)stream>>>(a, *field);
    return check_launch("density_grid_sweep");
}

uint32_t pnerf_density_finalize_partials(uint64_t n_cells) { return (uint32_t)ceil_div<uint64_t>(n_cells, (uint64_t)kFinThreads * kFinPerThread); }

/* density_grid <- max(density_grid * decay, tmp) where both >= 0 (nerf/renderer.py:543-545); tmp is reset to -1; then
 * mean = mean(clamp(density_grid, 0)), threshold = min(mean, density_thresh), bitfield = packbits(density_grid, threshold)
 * (:546-553). partials: scratch [pnerf_density_finalize_partials(C*H^3)]; stats [2] receives (mean, threshold). */
int pnerf_density_grid_finalize(float* density_grid, float* tmp_grid, uint32_t C, uint32_t H, float decay, float density_thresh,
                                float* partials, uint8_t* bitfield, float* stats, void* stream) {
    PNERF_REQUIRE(density_grid && tmp_grid && partials && bitfield && stats && C >= 1 && C <= 16 && H >= 2 && H <= 1024);
    const uint64_t n = (uint64_t)C * H * H * H;
    if (n % 8 != 0) return PNERF_ERR_UNSUPPORTED;
    const uint32_t nb = pnerf_density_finalize_partials(n);
    cudaStream_t s = (cudaStream_t)stream;
    k_density_finalize<<<nb, kFinThreads, 0, s>>>(density_grid, tmp_grid, n, decay, partials);
    const uint32_t n_bytes = (uint32_t)(n / 8);
    k_packbits_mean<<<ceil_div(n_bytes, 256u), 256, 0, s>>>(density_grid, n_bytes, partials, nb, n, density_thresh, bitfield, stats);
    return check_launch("density_grid_finalize");
}

}  // extern "C"
