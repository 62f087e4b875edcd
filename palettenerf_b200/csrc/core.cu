// core.cu — status / error plumbing of the C ABI.
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace pnerf {
static thread_local char g_last_err[256] = "";
void set_last_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_last_err, sizeof(g_last_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}
}  // namespace pnerf

extern "C" {

const char* pnerf_status_string(int status) {
    switch (status) {
        case PNERF_OK: return "ok";
        case PNERF_ERR_INVALID_ARG: return "invalid argument (null pointer / bad size / bad enum)";
        case PNERF_ERR_UNSUPPORTED: return "unsupported configuration";
        case PNERF_ERR_CUDA: return "CUDA error";
        default: return "unknown status";
    }
}

const char* pnerf_last_cuda_error(void) { return pnerf::g_last_err; }

int pnerf_abi_version(void) { return 2; }   // 2: occ_aabb argument of pnerf_palette_render_fused, workspace entry points, loss / Adam / get_rays

const char* pnerf_build_arch(void) { return "sm_100a"; }

// zero-fill of a caller-allocated gradient buffer as a memset node (cudaMemsetAsync: graph-capturable, runs at the HBM write
// rate) — the reference zero-fills its gradient buffers with torch.zeros / zeros_like (gridencoder/grid.py:72,
// raymarching/raymarching.py:283-284), whose elementwise fill kernel reaches a third of that on a 50 MB table gradient
int pnerf_zero_fill(void* dst, uint64_t bytes, void* stream) {
    if (bytes == 0) return PNERF_OK;
    if (!dst) return PNERF_ERR_INVALID_ARG;
    cudaError_t e = cudaMemsetAsync(dst, 0, (size_t)bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) { pnerf::set_last_cuda_error(e, "zero_fill"); return PNERF_ERR_CUDA; }
    return PNERF_OK;
}

}  // extern "C"
