// field_tc.cuh — the fused PaletteNeRF field on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// A WARPGROUP (4 warps, 128 threads) evaluates a tile of 128 samples: warp j of the group owns rows 32 j .. 32 j + 31 —
// which are exactly the TMEM lanes a warp with (warp id % 4) == j may read — so a THREAD owns one sample from the gather to
// the final palette heads. Per tile:
//   gather   lane-pair cooperative hash-grid gather (fused_common.cuh::gather_coop) straight into the A operands of the
//            first layers: fp16 feature rows in the canonical K-major UMMA layout [k-chunk][row][8 halfs] (tc_common.cuh);
//            SH(4) of the view direction the same way;
//   layers   11 (13 with the semantic branch) dense layers, each ONE tcgen05.mma chain issued by one thread of the group:
//            A = the gathered inputs in shared memory (SS form) and/or the previous layer's activations in TMEM (TS form),
//            B = the layer's weights resident in shared memory (same K-major layout), D = a 64-column fp32 accumulator in
//            TMEM; completion arrives on the group's mbarrier (tcgen05.commit);
//   epilogue every thread reads ITS row of the accumulator (tcgen05.ld 32x32b), applies ReLU / ELU / sigmoid, and writes the
//            next layer's A operand back to TMEM as packed fp16 pairs (tcgen05.st; the next tcgen05.mma takes A from TMEM)
//            — activations never touch shared memory — or keeps the value in registers when it is a result (sigma logit,
//            diffuse, view-dependent colour, palette heads, semantic feature).
// Per layer the group synchronises twice: a named barrier (A operand complete + fence.proxy.async) before the MMA, the
// mbarrier wait after it. The reference's concatenations cost nothing: [SH16 | logit, geo15] and [palette grid 32 | diffuse 3]
// are successive K = 16 steps of one MMA chain (the first from shared memory, the second from TMEM), the weight columns that
// must not see the logit are zero.
//
// Versus the mma.sync chain of fused_field.cuh (measured, profiles/README.md round 2): per 32 samples it replaces 310 HMMA
// + 310 LDS.64 of B fragments (every 16-row tile re-reads all weights from shared memory: 27 % of the L1 data-pipe
// wavefronts of the kernel) + the fragment repacking by ~25 tcgen05.ld and ~60 16-byte stores.
#pragma once
#include "fused_common.cuh"
#include "tc_common.cuh"

namespace pnerf {

// ---- weights: per layer [k-chunk][n][8 halfs], layers in this order (palettenerf_b200/fused_train.py::tc_pack_index) ----
// Two layers of the image are PRODUCTS of reference layers that have no activation between them (csrc/field_cache.cu::
// k_cache_merge computes them in fp32 whenever the weight cache is refreshed):
//   TD0 = diff_net.0 [64 x 15] x sigma_net.1[1:16] [15 x 64]: the geo features are a linear function of the sigma net's hidden
//         activations (nerf/network.py:101-107: no activation on the last sigma layer), so the first diffuse layer reads those
//         activations directly and runs in the SAME round as sigma_net.1 instead of waiting for it;
//   TB1 = [offsets_radiance_net ; omega_net.0] [17 x 15] x basis_net.1 [15 x 64] (palette/network.py:262-268: the heads are
//         applied to the basis net's raw output): one layer instead of two.
// 9 MMA rounds per tile instead of 11. (TH is kept in the table as 512 unused bytes so that offsets stay a pure function.)
enum TcLayer { TS0, TS1, TD0, TD1, TD2, TV0, TV1, TV2, TB0, TB1, TH, TC0, TC1, kTcLayers };
__host__ __device__ constexpr int tc_n(int l) {   // padded output width (multiple of 16: M = 128 needs N % 16 == 0)
    return l == TS0 ? 64 : l == TS1 ? 16 : l == TD0 ? 64 : l == TD1 ? 64 : l == TD2 ? 16 : l == TV0 ? 64 : l == TV1 ? 64
         : l == TV2 ? 16 : l == TB0 ? 64 : l == TB1 ? 32 : l == TH ? 16 : l == TC0 ? 64 : 16;
}
__host__ __device__ constexpr int tc_k(int l) {   // padded input width (multiple of 16)
    return l == TS0 ? 32 : l == TS1 ? 64 : l == TD0 ? 64 : l == TD1 ? 64 : l == TD2 ? 64 : l == TV0 ? 32 : l == TV1 ? 64
         : l == TV2 ? 64 : l == TB0 ? 48 : l == TB1 ? 64 : l == TH ? 16 : l == TC0 ? 32 : 64;
}
__host__ __device__ constexpr int tc_woff(int l) {   // byte offset of layer l in the weight image
    int o = 0;
    for (int i = 0; i < l; i++) o += tc_n(i) * tc_k(i) * 2;
    return o;
}
constexpr int kTcWBytesNoClip = tc_woff(TC0);
constexpr int kTcWBytesClip = tc_woff(kTcLayers);

// ---- shared memory of one warpgroup (bytes): only the gathered INPUTS, each region [k-chunk][128 rows][16 B] ----
constexpr int kTcChunk = 128 * 16;                 // one k-chunk (8 halfs) of 128 rows
constexpr int kTcRS = 0;                           // F_sigma   (4 chunks)
constexpr int kTcRP = kTcRS + 4 * kTcChunk;        // F_palette (4 chunks)
constexpr int kTcRH = kTcRP + 4 * kTcChunk;        // SH(4) of the view direction (2 chunks)
constexpr int kTcRC = kTcRH + 2 * kTcChunk;        // F_clip (4 chunks), only with the semantic branch
constexpr int kTcGroupBytesNoClip = kTcRC;
constexpr int kTcGroupBytesClip = kTcRC + 4 * kTcChunk;

// ---- TMEM columns of one warpgroup (128 of the SM's 512): accumulator + the activations that feed the next layers ----
// fp16 A operands live in TMEM packed two per 32-bit column (K = 16 per instruction = 8 columns)
constexpr uint32_t kTcColD = 0;                    // fp32 accumulator, up to 64 columns
constexpr uint32_t kTcColH = 64;                   // hidden activations, 64 halfs = 32 columns
constexpr uint32_t kTcColG = 96;                   // [sigma logit | geo 15]: 8 columns
constexpr uint32_t kTcColX = 104;                  // [diffuse 3 | 0 ...]: 8 columns
constexpr uint32_t kTcColS = 112;                  // second fp32 accumulator (16 columns): sigma_net.1 next to the first diffuse layer
constexpr uint32_t kTcColsPerGroup = 128;

struct TcShared {                                  // per CTA, in front of the weights
    LevelParams lp[16];
    float head_bias[16];
    float palette[kNB * 3];
    uint64_t mbar[4];                              // one per group
    uint32_t tmem_base;
    uint32_t flags[2][4][4];                       // per group: per-warp "has a tile" votes of the renderer (two alternating sets)
    float pend[16][8];                             // per warp: origin / direction of the ray that shares the current window
};

struct TcGroup {                                   // per-thread view of its group
    unsigned char* smem;                           // the group's input regions
    uint32_t smem_addr;                            // shared-window address of `smem`
    uint32_t w_addr;                               // shared-window address of the weight image
    uint32_t tmem;                                 // TMEM address of the group's columns at THIS warp's lanes
    uint32_t tmem0;                                // the same columns at lane 0 (what the MMA instruction takes)
    uint64_t* mbar;
    uint32_t phase;                                // parity of the next mbarrier completion
    uint32_t bar_id;                               // named barrier of the group
    int row;                                       // this thread's row in the tile (0..127)
    bool leader;                                   // the thread that issues the MMAs
};

__device__ __forceinline__ uint4* tc_row_ptr(unsigned char* region, int chunk, int row) {
    return reinterpret_cast<uint4*>(region + chunk * kTcChunk + row * 16);
}

// Source of one K = 16 step of a layer's A operand: shared memory (byte offset in the group's regions) or TMEM (column)
struct TcSrc { bool tmem; uint32_t at; };
__host__ __device__ constexpr TcSrc tc_smem(uint32_t byte_off) { return {false, byte_off}; }
__host__ __device__ constexpr TcSrc tc_tmem(uint32_t col) { return {true, col}; }

// one dense layer: D (TMEM column kTcColD..) = A * W_L^T with K/16 steps whose A halves come from `src[ks]`.
// Called by all 128 threads after they have written their part of A; returns when the accumulator is complete.
template <int L, int NS>
__device__ __forceinline__ void tc_layer(TcGroup& g, const TcSrc (&src)[NS], bool smem_written = false) {
    constexpr int N = tc_n(L), K = tc_k(L);
    static_assert(NS == K / 16, "one source per K = 16 step");
    tc::tmem_st_wait();             // this thread's tcgen05.st of the A operand have completed
    tc::tc_fence_before();          // ... and, with its tcgen05.ld of the previous accumulator, are ordered before the barrier
    if (smem_written) tc::fence_async_smem();   // st.shared of the gathered inputs -> visible to the tensor core
    tc::group_bar(g.bar_id, 128);
    if (g.leader) {
        tc::tc_fence_after();
        constexpr uint32_t idesc = tc::make_idesc_f16(128, N);
        const uint32_t b_addr = g.w_addr + tc_woff(L);
#pragma unroll
        for (int ks = 0; ks < NS; ks++) {
            const uint64_t bd = tc::make_smem_desc(b_addr + ks * 2 * (N * 16), N * 16, 128);
            if (src[ks].tmem) {
                tc::umma_f16_ts(g.tmem0 + kTcColD, g.tmem0 + src[ks].at, bd, idesc, ks > 0);
            } else {
                const uint64_t ad = tc::make_smem_desc(g.smem_addr + src[ks].at, kTcChunk, 128);
                tc::umma_f16(g.tmem0 + kTcColD, ad, bd, idesc, ks > 0);
            }
        }
        tc::umma_commit(g.mbar);
    }
    tc::mbar_wait(g.mbar, g.phase);
    g.phase ^= 1u;
    tc::tc_fence_after();
}

// two layers that read the SAME hidden activations (H columns) in one round: layer LA accumulates at column kTcColS (16 wide),
// layer LB at kTcColD; one commit, one wait
template <int LA, int LB>
__device__ __forceinline__ void tc_layer_pair_from_hidden(TcGroup& g) {
    static_assert(tc_k(LA) == 64 && tc_k(LB) == 64 && tc_n(LA) == 16, "both layers take the 64 hidden activations");
    tc::tmem_st_wait();
    tc::tc_fence_before();
    tc::group_bar(g.bar_id, 128);
    if (g.leader) {
        tc::tc_fence_after();
        constexpr uint32_t ia = tc::make_idesc_f16(128, tc_n(LA)), ib = tc::make_idesc_f16(128, tc_n(LB));
        const uint32_t wa = g.w_addr + tc_woff(LA), wb = g.w_addr + tc_woff(LB);
#pragma unroll
        for (int ks = 0; ks < 4; ks++)
            tc::umma_f16_ts(g.tmem0 + kTcColS, g.tmem0 + kTcColH + 8 * ks, tc::make_smem_desc(wa + ks * 2 * (tc_n(LA) * 16), tc_n(LA) * 16, 128),
                            ia, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ks++)
            tc::umma_f16_ts(g.tmem0 + kTcColD, g.tmem0 + kTcColH + 8 * ks, tc::make_smem_desc(wb + ks * 2 * (tc_n(LB) * 16), tc_n(LB) * 16, 128),
                            ib, ks > 0);
        tc::umma_commit(g.mbar);
    }
    tc::mbar_wait(g.mbar, g.phase);
    g.phase ^= 1u;
    tc::tc_fence_after();
}

// epilogue of a 64-wide hidden layer: activation, fp16 pairs, into the H columns of this thread's TMEM lane
template <int ACT>
__device__ __forceinline__ void tc_epilogue_hidden(const TcGroup& g) {
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
        uint32_t r0[16], r1[16], h[16];
        tc::tmem_ld16(g.tmem + kTcColD + c, r0);
        tc::tmem_ld16(g.tmem + kTcColD + c + 16, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; i++) {
            h[i] = pack_act<ACT>(__uint_as_float(r0[2 * i]), __uint_as_float(r0[2 * i + 1]));
            h[8 + i] = pack_act<ACT>(__uint_as_float(r1[2 * i]), __uint_as_float(r1[2 * i + 1]));
        }
        tc::tmem_st16(g.tmem + kTcColH + c / 2, h);
    }
}

// epilogue of a 16-wide layer: the 16 accumulators of this row (fp32) -> out, and as fp16 pairs into 8 TMEM columns
__device__ __forceinline__ void tc_epilogue_16(const TcGroup& g, float (&out)[16], uint32_t dst_col, bool store,
                                               uint32_t src_col = kTcColD) {
    uint32_t r[16];
    tc::tmem_ld16(g.tmem + src_col, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] = __uint_as_float(r[i]);
    if (store) {
        uint32_t h[8];
#pragma unroll
        for (int i = 0; i < 8; i++) h[i] = pack_h2(out[2 * i], out[2 * i + 1]);
        tc::tmem_st8(g.tmem + dst_col, h);
    }
}

// ------------------------------------------------------------------------------------------------
// the field for one 128-sample tile; called by all 128 threads of a group (thread = sample)
// ------------------------------------------------------------------------------------------------
// MODE selects the sub-network (all share one weight-image layout; unused layers are simply never issued):
//   TC_PALETTE / TC_PALETTE_CLIP  the palette field (PaletteNetwork.forward, palette/network.py:156-280)
//   TC_NERF     stage-1 field (NeRFNetwork.forward, nerf/network.py:95-124): sigma net + colour net (SH ++ geo), ONE hash grid;
//               the colour lands in o.view_dep
//   TC_DENSITY  sigma net only (NeRFNetwork.density / PaletteNetwork.density, used by the density-grid refresh)
enum TcMode { TC_PALETTE = 0, TC_PALETTE_CLIP = 1, TC_NERF = 2, TC_DENSITY = 3 };

// CLAMPED: the caller guarantees -bound <= x, y, z <= bound for every lane (the renderer clamps its samples): the gather then
// skips the out-of-range zeroing of the features (gridencoder.cu:118-130), one select per level
template <int MODE, bool CLAMPED = false>
__device__ __forceinline__ void eval_field_tc(const pnerf_palette_field& f, const TcShared& sm, TcGroup& g, float x, float y,
                                              float z, float dx, float dy, float dz, bool active, int lane, FieldOut& o) {
    static_assert(PNERF_COOP_LV == 4, "one gather batch = the 4 levels of one 16-byte k-chunk");
    constexpr bool CLIP = MODE == TC_PALETTE_CLIP;
    constexpr bool PAL = MODE == TC_PALETTE || MODE == TC_PALETTE_CLIP;
    const float u = (x + f.bound) / (2 * f.bound), v = (y + f.bound) / (2 * f.bound), w = (z + f.bound) / (2 * f.bound);
    const bool in_range = active && !((u < 0 || u > 1) || (v < 0 || v > 1) || (w < 0 || w > 1));
    unsigned char* const RS = g.smem + kTcRS;
    unsigned char* const RP = g.smem + kTcRP;
    unsigned char* const RH = g.smem + kTcRH;
    const int row0 = g.row - lane;                         // first row of this warp

    // ---- gather: both grids (one interleaved table) -> F_sigma, F_palette; SH of the view direction ----
    // (4 levels = 8 halfs = one k-chunk: a lane pair finishes a chunk per batch -> ONE conflict-free 16-byte store)
    if (PAL) {
        auto st = [RS, RP, row0](int e, int s, int l0, const uint32_t (&wd)[4]) {
            *tc_row_ptr(e == 0 ? RS : RP, l0 >> 2, row0 + s) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        };
        gather_coop<2, 4, PNERF_GATHER_HACC != 0, !CLAMPED>(f.table_sigma_palette, sm.lp, u, v, w, in_range, lane, st);
    } else {
        auto st = [RS, row0](int, int s, int l0, const uint32_t (&wd)[4]) {
            *tc_row_ptr(RS, l0 >> 2, row0 + s) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        };
        gather_coop<1, 4, PNERF_GATHER_HACC != 0, !CLAMPED>(f.table_sigma, sm.lp, u, v, w, in_range, lane, st);
    }
    if (CLIP) {
        unsigned char* const RC = g.smem + kTcRC;
        auto st = [RC, row0](int, int s, int l0, const uint32_t (&wd)[4]) {
            *tc_row_ptr(RC, l0 >> 2, row0 + s) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        };
        gather_coop<1, 4, PNERF_GATHER_HACC != 0, !CLAMPED>(f.table_clip, sm.lp, u, v, w, in_range, lane, st);
    }
    if (MODE != TC_DENSITY) {
        float sh[16];
        sh_eval<4, false>(dx, dy, dz, sh, nullptr, nullptr, nullptr);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint4 q;
            q.x = pack_h2(sh[8 * h + 0], sh[8 * h + 1]); q.y = pack_h2(sh[8 * h + 2], sh[8 * h + 3]);
            q.z = pack_h2(sh[8 * h + 4], sh[8 * h + 5]); q.w = pack_h2(sh[8 * h + 6], sh[8 * h + 7]);
            *tc_row_ptr(RH, h, g.row) = q;
        }
    }
    float t16[16];
    constexpr TcSrc kH4[4] = {tc_tmem(kTcColH), tc_tmem(kTcColH + 8), tc_tmem(kTcColH + 16), tc_tmem(kTcColH + 24)};

    // ---- sigma net 32 -> 64 -> 16: logit (col 0) + geo features (cols 1..15) ----
    {
        constexpr TcSrc a[2] = {tc_smem(kTcRS), tc_smem(kTcRS + 2 * kTcChunk)};
        tc_layer<TS0>(g, a, true);
    }
    tc_epilogue_hidden<ACT_RELU>(g);
    if (PAL) {
        // sigma_net.1 and the first diffuse layer (folded through sigma_net.1, see the layer table) in ONE round
        tc_layer_pair_from_hidden<TS1, TD0>(g);
        tc_epilogue_16(g, t16, kTcColG, true, kTcColS);
        o.sigma = fast_exp(t16[0]);
        // ---- diffuse net (15 ->) 64 -> 64 -> 3 ----
        tc_epilogue_hidden<ACT_RELU>(g);
        tc_layer<TD1>(g, kH4);
        tc_epilogue_hidden<ACT_RELU>(g);
        tc_layer<TD2>(g, kH4);
        {
            uint32_t r[8], h[8];
            tc::tmem_ld8(g.tmem + kTcColD, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 3; i++) o.diffuse[i] = sigmoidf_(__uint_as_float(r[i]));
            h[0] = pack_h2(o.diffuse[0], o.diffuse[1]); h[1] = pack_h2(o.diffuse[2], 0.f);
#pragma unroll
            for (int i = 2; i < 8; i++) h[i] = 0u;
            tc::tmem_st8(g.tmem + kTcColX, h);                 // third k-step of the basis net: [diffuse 3 | 0 ...]
        }
    } else {
        tc_layer<TS1>(g, kH4);
        tc_epilogue_16(g, t16, kTcColG, MODE != TC_DENSITY);
        o.sigma = fast_exp(t16[0]);
        if (MODE == TC_DENSITY) return;
    }

    // ---- view-dependent colour net (SH16 ++ geo15) -> 64 -> 64 -> 3 ----
    {
        constexpr TcSrc a[2] = {tc_smem(kTcRH), tc_tmem(kTcColG)};
        tc_layer<TV0>(g, a);
    }
    tc_epilogue_hidden<ACT_RELU>(g);
    tc_layer<TV1>(g, kH4);
    tc_epilogue_hidden<ACT_RELU>(g);
    tc_layer<TV2>(g, kH4);
    {
        uint32_t r[8];
        tc::tmem_ld8(g.tmem + kTcColD, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 3; i++) o.view_dep[i] = sigmoidf_(__uint_as_float(r[i]));
    }

    if (!PAL) return;

    // ---- basis net (palette grid 32 ++ diffuse 3) -> 64 (ELU) -> [15 ->] offsets / radiance / omega heads ----
    {
        constexpr TcSrc a[3] = {tc_smem(kTcRP), tc_smem(kTcRP + 2 * kTcChunk), tc_tmem(kTcColX)};
        tc_layer<TB0>(g, a);
    }
    tc_epilogue_hidden<ACT_ELU>(g);
    tc_layer<TB1>(g, kH4);                                  // basis_net.1 and the heads as one layer (see the layer table)
    {
        uint32_t r0[16], r1[8];
        tc::tmem_ld16(g.tmem + kTcColD, r0);
        tc::tmem_ld8(g.tmem + kTcColD + 16, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 13; i++) o.off_rad[i] = __uint_as_float(r0[i]) + sm.head_bias[i];
        const float lg[kNB] = {__uint_as_float(r0[13]), __uint_as_float(r0[14]), __uint_as_float(r0[15]), __uint_as_float(r1[0])};
        float osum = 0.f;
#pragma unroll
        for (int b = 0; b < kNB; b++) { o.omega[b] = softplusf_(lg[b]) + 0.05f; osum += o.omega[b]; }
        const float rinv = 1.0f / osum;
#pragma unroll
        for (int b = 0; b < kNB; b++) o.omega[b] *= rinv;
    }

    // ---- semantic branch: clip grid 32 -> 64 -> clip_dim ----
#pragma unroll
    for (int i = 0; i < kClipMax; i++) o.clip[i] = 0.f;
    if (CLIP) {
        {
            constexpr TcSrc a[2] = {tc_smem(kTcRC), tc_smem(kTcRC + 2 * kTcChunk)};
            tc_layer<TC0>(g, a);
        }
        tc_epilogue_hidden<ACT_RELU>(g);
        tc_layer<TC1>(g, kH4);
        tc_epilogue_16(g, t16, 0, false);
#pragma unroll
        for (int i = 0; i < kClipMax; i++) o.clip[i] = i < (int)f.clip_dim ? t16[i] : 0.f;
    }
}

// CTA prologue: level table, head bias, palette, weight image -> shared memory; TMEM; mbarriers; zero the constant chunk
__host__ __device__ constexpr uint32_t tc_tmem_cols(int groups) {      // power of two >= 32
    return groups * kTcColsPerGroup <= 128 ? 128u : groups * kTcColsPerGroup <= 256 ? 256u : 512u;
}

template <int GROUPS>
__device__ __forceinline__ void tc_prologue(const pnerf_palette_field& f, const void* wimage, TcShared* sm, unsigned char* wts,
                                            unsigned char* groups, int group_bytes) {
    const int tid = threadIdx.x;
    if (tid < 16) {
        make_level(sm->lp[tid], tid, f.offsets, f.S, f.H, 3, 0, false);
        sm->head_bias[tid] = f.head_bias[tid];
    }
    if (tid < kNB * 3) sm->palette[tid] = f.palette[tid];
    const int wbytes = f.pred_clip ? kTcWBytesClip : kTcWBytesNoClip;
    const uint4* src = reinterpret_cast<const uint4*>(wimage);
    uint4* dst = reinterpret_cast<uint4*>(wts);
    for (int i = tid; i < wbytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    if (tid < 32) tc::tmem_alloc(&sm->tmem_base, tc_tmem_cols(GROUPS));
    if (tid == 32) {
        for (int i = 0; i < GROUPS; i++) tc::mbar_init(&sm->mbar[i], 1);
        tc::fence_mbar_init();
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
}

template <int GROUPS>
__device__ __forceinline__ void tc_epilogue_cta(TcShared* sm) {
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(sm->tmem_base, tc_tmem_cols(GROUPS));
}

__device__ __forceinline__ TcGroup tc_make_group(TcShared* sm, unsigned char* wts, unsigned char* groups, int group_bytes) {
    TcGroup g;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, gi = wid >> 2, wig = wid & 3;
    g.smem = groups + gi * group_bytes;
    g.smem_addr = tc::smem_u32(g.smem);
    g.w_addr = tc::smem_u32(wts);
    g.tmem0 = sm->tmem_base + (uint32_t)gi * kTcColsPerGroup;
    g.tmem = g.tmem0 + ((uint32_t)(wig * 32) << 16);
    g.mbar = &sm->mbar[gi];
    g.phase = 0u;
    g.bar_id = 1u + (uint32_t)gi;
    g.row = wig * 32 + lane;
    g.leader = (wig == 0 && lane == 0);
    return g;
}

}  // namespace pnerf
