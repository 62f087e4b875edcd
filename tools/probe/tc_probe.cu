// tc_probe.cu — stand-alone probe of the tcgen05 building blocks used by csrc/field_tc.cuh: TMEM allocation, K-major
// no-swizzle shared-memory descriptors ([k-chunk][row][8 halfs] layout: LBO = rows * 16 B, SBO = 128 B), the f16 instruction
// descriptor, tcgen05.commit -> mbarrier, tcgen05.ld 32x32b. Computes D[128 x N] = A[128 x K] * B[N x K]^T for a few
// (N, K) and compares with the host. Build + run (GPU box): nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tc_probe
// tools/probe/tc_probe.cu && /tmp/tc_probe
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "../../palettenerf_b200/csrc/tc_common.cuh"

using namespace pnerf::tc;

template <int N, int K>
__global__ void __launch_bounds__(128) k_probe(const __half* A, const __half* B, float* D) {
    extern __shared__ __align__(128) unsigned char smem[];
    __half* sA = reinterpret_cast<__half*>(smem);
    __half* sB = sA + 128 * K;
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * K; i += 128) { int r = i / K, k = i % K; sA[(k / 8) * 128 * 8 + r * 8 + (k % 8)] = A[i]; }
    for (int i = tid; i < N * K; i += 128) { int r = i / K, k = i % K; sB[(k / 8) * N * 8 + r * 8 + (k % 8)] = B[i]; }
    if (warp == 0) tmem_alloc(&tbase, 64);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t t0 = tbase;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_f16(128, N);
        for (int ks = 0; ks < K / 16; ks++) {
            const uint64_t ad = make_smem_desc(smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
            const uint64_t bd = make_smem_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
            umma_f16(t0, ad, bd, idesc, ks > 0);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t r[16];
    for (int c = 0; c < N; c += 16) {
        tmem_ld16(t0 + ((uint32_t)(warp * 32) << 16) + c, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; j++) D[(warp * 32 + lane) * N + c + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(t0, 64);
}

// A operand in TMEM: every thread packs its row of A to fp16 pairs and writes it with tcgen05.st; mixed chain: k-step 0 from
// shared memory (SS form), the remaining k-steps from TMEM (TS form), all into one accumulator
template <int N, int K>
__global__ void __launch_bounds__(128) k_probe_ts(const __half* A, const __half* B, float* D) {
    extern __shared__ __align__(128) unsigned char smem[];
    __half* sA = reinterpret_cast<__half*>(smem);
    __half* sB = sA + 128 * K;
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * K; i += 128) { int r = i / K, k = i % K; sA[(k / 8) * 128 * 8 + r * 8 + (k % 8)] = A[i]; }
    for (int i = tid; i < N * K; i += 128) { int r = i / K, k = i % K; sB[(k / 8) * N * 8 + r * 8 + (k % 8)] = B[i]; }
    if (warp == 0) tmem_alloc(&tbase, 128);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t t0 = tbase;
    const uint32_t mine = t0 + ((uint32_t)(warp * 32) << 16);
    // this thread's row of A -> TMEM columns [64, 64 + K/2)
    const int row = warp * 32 + lane;
    for (int c = 0; c < K / 2; c += 8) {
        uint32_t w[8];
        for (int j = 0; j < 8; j++) {
            const __half2 h = __halves2half2(A[row * K + 2 * (c + j)], A[row * K + 2 * (c + j) + 1]);
            w[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        tmem_st8(mine + 64 + c, w);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16(128, N);
        for (int ks = 0; ks < K / 16; ks++) {
            const uint64_t bd = make_smem_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
            if (ks == 0) {
                const uint64_t ad = make_smem_desc(smem_u32(sA), 128 * 16, 128);
                umma_f16(t0, ad, bd, idesc, 0);
            } else {
                umma_f16_ts(t0, t0 + 64 + ks * 8, bd, idesc, 1);
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t r[16];
    for (int c = 0; c < N; c += 16) {
        tmem_ld16(mine + c, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; j++) D[row * N + c + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(t0, 128);
}

template <int N, int K, bool TS = false>
static int run() {
    std::vector<__half> hA(128 * K), hB(N * K);
    std::vector<float> fA(128 * K), fB(N * K), hD(128 * N);
    for (int i = 0; i < 128 * K; i++) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < N * K; i++) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, hD.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    if (TS) k_probe_ts<N, K><<<1, 128, (128 + N) * K * 2>>>(dA, dB, dD);
    else k_probe<N, K><<<1, 128, (128 + N) * K * 2>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d K=%d: CUDA error %s\n", N, K, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int r = 0; r < 128; r++)
        for (int n = 0; n < N; n++) {
            double s = 0;
            for (int k = 0; k < K; k++) s += (double)fA[r * K + k] * fB[n * K + k];
            maxerr = fmax(maxerr, fabs(s - hD[r * N + n]));
        }
    printf("%sN=%d K=%d: max |D - ref| = %.3e %s\n", TS ? "A in TMEM: " : "", N, K, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return maxerr < 1e-3 ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run<64, 16>();
    bad += run<64, 64>();
    bad += run<16, 64>();
    bad += run<32, 16>();
    bad += run<64, 48>();
    bad += run<64, 32>();
    bad += run<64, 64, true>();
    bad += run<16, 64, true>();
    bad += run<64, 32, true>();
    bad += run<32, 16, true>();
    printf(bad ? "tc_probe: FAILED\n" : "tc_probe: all OK\n");
    return bad;
}
