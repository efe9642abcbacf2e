// score_common.cuh -- the residual test shared by the SIMT and the tcgen05 scoring kernels.
// Spelled identically in the oracle (oracle/erp_oracle.c: is_inlier), so inlier counts are bit-exact.
#pragma once
#include "common.cuh"

namespace erp {

template <int METRIC>
__device__ __forceinline__ bool inlier(const float* __restrict__ E, const float k[9], float4 l, float4 r,
                                       float tau, float tau2, float sin2)
{
    float res = __fmul_rn(E[0], k[0]);
#pragma unroll
    for (int i = 1; i < 9; i++) res = __fmaf_rn(E[i], k[i], res);
    if (METRIC == ERP_METRIC_ALGEBRAIC) return fabsf(res) < tau;
    float n0 = __fmaf_rn(E[2], r.z, __fmaf_rn(E[1], r.y, __fmul_rn(E[0], r.x)));
    float n1 = __fmaf_rn(E[5], r.z, __fmaf_rn(E[4], r.y, __fmul_rn(E[3], r.x)));
    float n2 = __fmaf_rn(E[8], r.z, __fmaf_rn(E[7], r.y, __fmul_rn(E[6], r.x)));
    float nn = __fmaf_rn(n2, n2, __fmaf_rn(n1, n1, __fmul_rn(n0, n0)));
    float rr = __fmul_rn(res, res);
    if (METRIC == ERP_METRIC_ANGULAR) return rr < __fmul_rn(sin2, nn);
    float m0 = __fmaf_rn(E[6], l.z, __fmaf_rn(E[3], l.y, __fmul_rn(E[0], l.x)));
    float m1 = __fmaf_rn(E[7], l.z, __fmaf_rn(E[4], l.y, __fmul_rn(E[1], l.x)));
    float m2 = __fmaf_rn(E[8], l.z, __fmaf_rn(E[5], l.y, __fmul_rn(E[2], l.x)));
    float mm = __fmaf_rn(m2, m2, __fmaf_rn(m1, m1, __fmul_rn(m0, m0)));
    return rr < __fmul_rn(tau2, __fadd_rn(nn, mm));
}

__device__ __forceinline__ void kron9(float4 l, float4 r, float k[9])
{
    k[0] = __fmul_rn(l.x, r.x); k[1] = __fmul_rn(l.x, r.y); k[2] = __fmul_rn(l.x, r.z);
    k[3] = __fmul_rn(l.y, r.x); k[4] = __fmul_rn(l.y, r.y); k[5] = __fmul_rn(l.y, r.z);
    k[6] = __fmul_rn(l.z, r.x); k[7] = __fmul_rn(l.z, r.y); k[8] = __fmul_rn(l.z, r.z);
}

__device__ __forceinline__ void scale_E(const double* __restrict__ E, float* __restrict__ Eh)
{
    double n = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) n += E[i] * E[i];
    double s = n > 0 ? sqrt(2.0) / sqrt(n) : 0.0;
#pragma unroll
    for (int i = 0; i < 9; i++) Eh[i] = (float)(E[i] * s);
}

} // namespace erp
