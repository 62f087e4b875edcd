// rays.cu — camera-ray generation for B200 (sm_100a): pixel index -> (rays_o, rays_d [, near, far]).
//
// Replaces the tensor program of get_rays (ref: nerf/utils.py:52-151, the part after the pixel indices are drawn,
// :132-149): the reference builds two full [B, H*W] float meshgrids per call (even to pick 4096 training pixels),
// gathers them, stacks / normalises / matmuls the directions (~15 torch kernels, each round-tripping [B,N,3]
// through HBM) and returns rays_o as an expanded view that the marcher then has to materialise. Here one kernel
// reads 8 B per ray (the pixel index; nothing for a full image) and writes the 24 B the marcher needs, plus,
// optionally, the slab test of near_far_from_aabb (ref: raymarching.cu:95-148) on the direction still in registers
// (saves re-reading 24 B per ray and a launch).
//
// Arithmetic, expression by expression (fp32, round-to-nearest, no contraction where torch has none):
//   i = col + 0.5, j = row + 0.5           (linspace(0, W-1, W) is exact in fp32; nerf/utils.py:69-71)
//   x = (i - cx) / fx, y = (j - cy) / fy, z = 1                                   (:134-136)
//   d = (x, y, z) / sqrt(x*x + y*y + z*z)  (torch.norm: sum of squares in index order, then sqrt; :137-138)
//   rays_d[k] = d . R[k, :]                (directions @ R^T, K = 3 accumulated in index order; :139)
//   rays_o    = poses[:, :3, 3]            (:141-142)
// The matmul's accumulation order is cuBLAS's (not specified); parity with the reference is to 1 ulp-level
// tolerance (tests/test_rays_gpu.py), not bit-exact.
#include "common.cuh"
#include "ray_common.cuh"

namespace pnerf {

constexpr int kRayBlock = 256;

// rays of one block are staged in shared memory and written as float4 (768 floats = 192 float4 per array)
__global__ void __launch_bounds__(kRayBlock) k_get_rays(const float* __restrict__ poses, float fx, float fy, float cx,
                                                        float cy, uint32_t H, uint32_t W, const int64_t* __restrict__ inds,
                                                        uint64_t inds_batch_stride, uint32_t N, uint32_t B,
                                                        float* __restrict__ rays_o, float* __restrict__ rays_d,
                                                        const float* __restrict__ aabb, float min_near,
                                                        float* __restrict__ nears, float* __restrict__ fars,
                                                        const float* __restrict__ img0, uint32_t c0, float* __restrict__ out0,
                                                        const float* __restrict__ img1, uint32_t c1, float* __restrict__ out1) {
    __shared__ __align__(16) float so[kRayBlock * 3];
    __shared__ __align__(16) float sd[kRayBlock * 3];
    const uint32_t b = blockIdx.y;
    const uint32_t n0 = blockIdx.x * kRayBlock, n = n0 + threadIdx.x;
    const float* P = poses + (size_t)b * 16;
    const float ox = __ldg(P + 3), oy = __ldg(P + 7), oz = __ldg(P + 11);
    if (n < N) {
        const uint64_t pix = inds ? (uint64_t)__ldg(inds + (size_t)b * inds_batch_stride + n) : (uint64_t)n;
        const uint32_t row = (uint32_t)(pix / W), col = (uint32_t)(pix - (uint64_t)row * W);
        const float i = (float)col + 0.5f, j = (float)row + 0.5f;
        const float x = __fdiv_rn(__fsub_rn(i, cx), fx), y = __fdiv_rn(__fsub_rn(j, cy), fy), z = 1.0f;
        const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), 1.0f));
        const float ux = __fdiv_rn(x, nrm), uy = __fdiv_rn(y, nrm), uz = __fdiv_rn(z, nrm);
        float d[3];
#pragma unroll
        for (int k = 0; k < 3; k++)
            d[k] = fmaf(uz, __ldg(P + 4 * k + 2), fmaf(uy, __ldg(P + 4 * k + 1), __fmul_rn(ux, __ldg(P + 4 * k))));
        so[threadIdx.x * 3 + 0] = ox; so[threadIdx.x * 3 + 1] = oy; so[threadIdx.x * 3 + 2] = oz;
        sd[threadIdx.x * 3 + 0] = d[0]; sd[threadIdx.x * 3 + 1] = d[1]; sd[threadIdx.x * 3 + 2] = d[2];
        if (nears) {
            float near, far;
            slab_near_far(ox, oy, oz, d[0], d[1], d[2], aabb, min_near, near, far);
            nears[(size_t)b * N + n] = near;
            fars[(size_t)b * N + n] = far;
        }
        // training-pixel gather of the data loader's collate (ref: palette/provider.py:387-399: torch.gather of the ground-truth
        // image and of the semantic feature image at the sampled pixels), on the pixel index already in registers
        if (img0) {
            const float* src = img0 + ((size_t)b * H * W + pix) * c0;
            float* dst = out0 + ((size_t)b * N + n) * c0;
            for (uint32_t k = 0; k < c0; k++) dst[k] = __ldg(src + k);
        }
        if (img1) {
            const float* src = img1 + ((size_t)b * H * W + pix) * c1;
            float* dst = out1 + ((size_t)b * N + n) * c1;
            for (uint32_t k = 0; k < c1; k++) dst[k] = __ldg(src + k);
        }
    }
    __syncthreads();
    const uint32_t cnt = min((uint32_t)kRayBlock, N - n0) * 3;           // floats of this block
    const size_t base = ((size_t)b * N + n0) * 3;                         // float offset; 16-byte aligned iff base % 4 == 0
    float* go = rays_o + base;
    float* gd = rays_d + base;
    if ((base & 3) == 0 && ((reinterpret_cast<uintptr_t>(rays_o) | reinterpret_cast<uintptr_t>(rays_d)) & 15) == 0) {
        const uint32_t n4 = cnt / 4;
        for (uint32_t t = threadIdx.x; t < n4; t += kRayBlock) {
            st_stream4(reinterpret_cast<float4*>(go) + t, reinterpret_cast<const float4*>(so)[t]);
            st_stream4(reinterpret_cast<float4*>(gd) + t, reinterpret_cast<const float4*>(sd)[t]);
        }
        for (uint32_t t = n4 * 4 + threadIdx.x; t < cnt; t += kRayBlock) { go[t] = so[t]; gd[t] = sd[t]; }
    } else {
        for (uint32_t t = threadIdx.x; t < cnt; t += kRayBlock) { go[t] = so[t]; gd[t] = sd[t]; }
    }
}

}  // namespace pnerf

using namespace pnerf;

extern "C" {

int pnerf_get_rays(const float* poses, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W, const int64_t* inds,
                   uint64_t inds_batch_stride, uint32_t N, uint32_t B, float* rays_o, float* rays_d, const float* aabb,
                   float min_near, float* nears, float* fars, void* stream) {
    if (N == 0 || B == 0) return PNERF_OK;
    PNERF_REQUIRE(poses && rays_o && rays_d && H >= 1 && W >= 1);
    PNERF_REQUIRE(inds || (uint64_t)N == (uint64_t)H * W);               // without indices: the full image, in pixel order
    PNERF_REQUIRE((nears == nullptr) == (fars == nullptr));
    PNERF_REQUIRE(nears == nullptr || aabb != nullptr);
    if (B > 65535u) return PNERF_ERR_UNSUPPORTED;
    const dim3 grid(ceil_div(N, (uint32_t)kRayBlock), B, 1);
    k_get_rays<<<grid, kRayBlock, 0, (cudaStream_t)stream>>>(poses, fx, fy, cx, cy, H, W, inds, inds_batch_stride, N, B, rays_o,
                                                            rays_d, aabb, min_near, nears, fars, nullptr, 0, nullptr, nullptr, 0,
                                                            nullptr);
    return check_launch("get_rays");
}

/* pnerf_get_rays + the pixel gathers of the training data loader's collate in the same launch: images [B, H*W, c_img] ->
 * out_images [B, N, c_img] and feat_images [B, H*W, c_feat] -> out_feat [B, N, c_feat] at the same pixel indices (either pair
 * may be NULL). Replaces get_rays + the two torch.gather calls of palette/provider.py:377-399. */
int pnerf_get_rays_collate(const float* poses, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W, const int64_t* inds,
                           uint64_t inds_batch_stride, uint32_t N, uint32_t B, float* rays_o, float* rays_d, const float* aabb,
                           float min_near, float* nears, float* fars, const float* images, uint32_t c_img, float* out_images,
                           const float* feat_images, uint32_t c_feat, float* out_feat, void* stream) {
    if (N == 0 || B == 0) return PNERF_OK;
    PNERF_REQUIRE(poses && rays_o && rays_d && H >= 1 && W >= 1);
    PNERF_REQUIRE(inds || (uint64_t)N == (uint64_t)H * W);
    PNERF_REQUIRE((nears == nullptr) == (fars == nullptr));
    PNERF_REQUIRE(nears == nullptr || aabb != nullptr);
    PNERF_REQUIRE((images == nullptr) == (out_images == nullptr) && (feat_images == nullptr) == (out_feat == nullptr));
    PNERF_REQUIRE((!images || c_img >= 1) && (!feat_images || c_feat >= 1));
    if (B > 65535u) return PNERF_ERR_UNSUPPORTED;
    const dim3 grid(ceil_div(N, (uint32_t)kRayBlock), B, 1);
    k_get_rays<<<grid, kRayBlock, 0, (cudaStream_t)stream>>>(poses, fx, fy, cx, cy, H, W, inds, inds_batch_stride, N, B, rays_o,
                                                            rays_d, aabb, min_near, nears, fars, images, c_img, out_images,
                                                            feat_images, c_feat, out_feat);
    return check_launch("get_rays_collate");
}

}  // extern "C"
