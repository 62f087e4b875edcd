// peer.cu — gradient all-reduce of the data-parallel training step over NVLink peer memory.
//
// The reference has no multi-GPU path; north_star asks for ONE all-reduce of the hash-table + MLP gradients per step
// (50.7 MB fp32 in the palette stage). NCCL's all-reduce of that bucket costs ~0.3 ms at 8 ranks inside the captured
// step (profiles/README.md) — as much as the forward and backward field kernels together. On one NVSwitch box every GPU
// can load and store every peer's memory at NVLink speed, so the exchange is written as a plain kernel over the peers'
// buffers (two-shot all-reduce):
//     rank r owns slice r of the bucket: it reads slice r from ALL ranks' buffers (world - 1 of them remote), adds them
//     in rank order (one owner per element: every rank receives bit-identical sums), scales by 1/world, and stores the
//     result into slice r of ALL ranks' buffers.
// Per rank and direction that is (world-1)/world of the bucket over NVLink, with 128-bit accesses; the two cross-GPU
// barriers around the kernel (all gradients packed / all slices delivered) come from the symmetric-memory handle that
// also provides the peer pointers (palettenerf_b200/distributed.py). No slice is read and written by different ranks,
// so the kernel needs no synchronisation of its own.
#include "common.cuh"

namespace pnerf {

constexpr int kPeerMax = PNERF_PEER_MAX;

struct PeerArgs {
    float* buf[kPeerMax];
    uint32_t world, rank;
    uint64_t n4_per_rank;     // float4 elements per slice (the bucket is padded to world * 4 floats)
    float scale;
};

__device__ __forceinline__ float4 ld_sys4(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys4(float4* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(256) k_peer_allreduce(const __grid_constant__ PeerArgs a) {
    const uint64_t base = (uint64_t)a.rank * a.n4_per_rank;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4_per_rank; i += stride) {
        float4 v[kPeerMax];
#pragma unroll
        for (int p = 0; p < kPeerMax; p++)
            if (p < (int)a.world) v[p] = ld_sys4(reinterpret_cast<const float4*>(a.buf[p]) + base + i);    // all loads in flight
        float4 s = v[0];
#pragma unroll
        for (int p = 1; p < kPeerMax; p++)
            if (p < (int)a.world) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }       // rank order: deterministic
        s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
#pragma unroll
        for (int p = 0; p < kPeerMax; p++)
            if (p < (int)a.world) st_sys4(reinterpret_cast<float4*>(a.buf[p]) + base + i, s);
    }
}

// ------------------------------------------------------------------------------------------------
// The same all-reduce with the reduction done INSIDE THE NVSWITCH (NVLS): the bucket is also mapped at a multicast address;
// multimem.ld_reduce on it returns the sum over all ranks' copies of that address (the switch reads the peers and adds),
// multimem.st writes one value into every rank's copy. Rank r still owns slice r — one owner per element, bit-identical
// results everywhere — but per GPU only 1/world of the bucket crosses its NVLink ports in each direction (instead of
// (world-1)/world with peer loads and stores), and the SMs issue one load and one store per 16 bytes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_peer_allreduce_mc(float* __restrict__ mc, uint32_t rank, uint64_t n4_per_rank, float scale) {
    const uint64_t base = (uint64_t)rank * n4_per_rank;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4_per_rank; i += stride) {
        float4* p = reinterpret_cast<float4*>(mc) + base + i;
        float4 s;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(s.x), "=f"(s.y), "=f"(s.z), "=f"(s.w)
                     : "l"(p)
                     : "memory");
        s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(s.x), "f"(s.y), "f"(s.z), "f"(s.w)
                     : "memory");
    }
}

}  // namespace pnerf

using namespace pnerf;

extern "C" {

/* all-reduce through the multicast mapping of the bucket (in-switch reduction, NVLS): mc_ptr = multicast address of the
 * symmetric buffer (torch symmetric memory: handle.multicast_ptr); same contract as pnerf_peer_allreduce otherwise */
int pnerf_peer_allreduce_mc(uint64_t mc_ptr, uint32_t world, uint32_t rank, uint64_t n, float scale, void* stream) {
    PNERF_REQUIRE(mc_ptr != 0 && (mc_ptr & 15ull) == 0 && world >= 1 && world <= (uint32_t)kPeerMax && rank < world);
    PNERF_REQUIRE(n % (4ull * world) == 0);
    if (n == 0) return PNERF_OK;
    const uint64_t n4 = n / 4 / world;
    const uint64_t blocks = ceil_div<uint64_t>(n4, 256);
    const uint32_t grid = (uint32_t)(blocks < 4ull * kNumSMs ? blocks : 4ull * kNumSMs);
    k_peer_allreduce_mc<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float*>(mc_ptr), rank, n4, scale);
    return check_launch("peer_allreduce_mc");
}

int pnerf_peer_allreduce(const uint64_t* peer_ptrs, uint32_t world, uint32_t rank, uint64_t n, float scale, void* stream) {
    PNERF_REQUIRE(peer_ptrs != nullptr && world >= 1 && world <= (uint32_t)kPeerMax && rank < world);
    PNERF_REQUIRE(n % (4ull * world) == 0);                 // callers pad the bucket
    if (n == 0) return PNERF_OK;
    PeerArgs a;
    for (uint32_t p = 0; p < world; p++) {
        PNERF_REQUIRE(peer_ptrs[p] != 0 && (peer_ptrs[p] & 15ull) == 0);
        a.buf[p] = reinterpret_cast<float*>(peer_ptrs[p]);
    }
    for (uint32_t p = world; p < (uint32_t)kPeerMax; p++) a.buf[p] = nullptr;
    a.world = world; a.rank = rank; a.n4_per_rank = n / 4 / world; a.scale = scale;
    const uint64_t blocks = ceil_div<uint64_t>(a.n4_per_rank, 256);
    const uint32_t grid = (uint32_t)(blocks < 4ull * kNumSMs ? blocks : 4ull * kNumSMs);
    k_peer_allreduce<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("peer_allreduce");
}

}  // extern "C"
