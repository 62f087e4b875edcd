// sh_common.cuh — structured evaluation of the real SH basis (see shenc.cu / tools/gen_sh.py), shared with fused.cu
#pragma once
#include "common.cuh"
#include "sh_tables.h"

namespace pnerf {

// out[DEG*DEG]; with GRAD also d/dx, d/dy, d/dz tables
template <int DEG, bool GRAD>
__device__ __forceinline__ void sh_eval(float x, float y, float z, float* __restrict__ out, float* __restrict__ gx,
                                        float* __restrict__ gy, float* __restrict__ gz) {
    float A[DEG], Bm[DEG];
    A[0] = 1.f; Bm[0] = 0.f;
#pragma unroll
    for (int m = 1; m < DEG; m++) {
        A[m] = x * A[m - 1] - y * Bm[m - 1];
        Bm[m] = x * Bm[m - 1] + y * A[m - 1];
    }
#pragma unroll
    for (int l = 0; l < DEG; l++) {
#pragma unroll
        for (int m = 0; m <= l; m++) {
            float q = 0.f, dq = 0.f;
#pragma unroll
            for (int k = l - m; k >= 0; k--) q = q * z + kShCoef[l][m][k];
            if (GRAD) {
#pragma unroll
                for (int k = l - m; k >= 1; k--) dq = dq * z + (float)k * kShCoef[l][m][k];
            }
            const int ip = l * l + l + m, im = l * l + l - m;
            if (m == 0) {
                out[ip] = q;
                if (GRAD) { gx[ip] = 0.f; gy[ip] = 0.f; gz[ip] = dq; }
            } else {
                out[ip] = q * A[m];
                out[im] = q * Bm[m];
                if (GRAD) {
                    const float qm = q * (float)m;
                    gx[ip] = qm * A[m - 1];  gy[ip] = -qm * Bm[m - 1]; gz[ip] = dq * A[m];
                    gx[im] = qm * Bm[m - 1]; gy[im] = qm * A[m - 1];   gz[im] = dq * Bm[m];
                }
            }
        }
    }
}

}  // namespace pnerf
