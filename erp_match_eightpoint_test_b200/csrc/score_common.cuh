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

// ---- operands and device words of the tensor-core best-hypothesis search (score_tc.cu) -----------------------
// One 128-byte row (32 floats) per hypothesis / correspondence carries the three 3xTF32 products of the 9-term dot:
//     hypothesis     [ Eh_hi(9) | Eh_hi(9) | Eh_lo(9) | 0 x 5 ]   (times a power of two, see score_tc.cu)
//     correspondence [ K_hi(9)  | K_lo(9)  | K_hi(9)  | 0 x 5 ]
__device__ __forceinline__ void write_e_row(const float e[9], float big, float* __restrict__ row)
{
    float v[32];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const float hi = tf32_rna(e[i]), lo = tf32_rna(__fsub_rn(e[i], hi));
        v[i] = hi * big; v[9 + i] = hi * big; v[18 + i] = lo * big;       // exact: power of two
    }
#pragma unroll
    for (int i = 27; i < 32; i++) v[i] = 0.f;
    float4* o = reinterpret_cast<float4*>(row);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// returns |l||r| (the scale of the certificate's band)
__device__ __forceinline__ float write_k_row(float4 l, float4 r, float* __restrict__ row)
{
    float k[9], v[32];
    kron9(l, r, k);
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const float hi = tf32_rna(k[i]), lo = tf32_rna(__fsub_rn(k[i], hi));
        v[i] = hi; v[9 + i] = lo; v[18 + i] = hi;
    }
#pragma unroll
    for (int i = 27; i < 32; i++) v[i] = 0.f;
    float4* o = reinterpret_cast<float4*>(row);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    return sqrtf((l.x * l.x + l.y * l.y + l.z * l.z) * (r.x * r.x + r.y * r.y + r.z * r.z));
}
__device__ __forceinline__ void zero_row(float* __restrict__ row)
{
    float4* o = reinterpret_cast<float4*>(row);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

constexpr int SC_TILE = 256;      // correspondences per tile of the search (UMMA N)
constexpr int SC_N0 = 8;          // tiles of pass A

// int32 words every pass reads from device memory (no length of the search ever visits the host)
enum ScoreWords {
    W_KMAX = 0, W_M = 1, W_NCT = 2, W_N0 = 3,                // set once per call by the correspondence prep
    W_AMAX = 4 /* 2 words */, W_DONE = 6, W_LEN1 = 7, W_LENF = 8, W_LSTAR = 9, W_REMAIN = 10, W_FIRST = 11,
    W_DYN_A = 12, W_DYN_B = 16, W_DYN_C = 20,                // { rows of the A matrix, first tile, end tile, m }
    W_WORDS = 24, W_CHUNK0 = 4                               // words [W_CHUNK0, W_WORDS) are reset per hypothesis chunk
};
static_assert(W_WORDS * sizeof(int32_t) == W_WORDS_BYTES, "common.cuh: W_WORDS_BYTES");

// the correspondence side of the search, one thread per correspondence slot c < capacity: operand row (zeros up to the
// tile boundary past m: those columns have res = 0 exactly and are subtracted by the kernel), |l||r| maximum, and the
// m-dependent words (by the thread that owns slot 0)
__device__ __forceinline__ void prep_k_slot(int c, int m, float4 l, float4 r, float* __restrict__ Ks, int32_t* __restrict__ w)
{
    if (c == 0) {
        const int nct = (m + SC_TILE - 1) / SC_TILE;
        w[W_M] = m; w[W_NCT] = nct; w[W_N0] = nct < SC_N0 ? nct : SC_N0;
    }
    float nrm = 0.f;
    if (c < m) nrm = write_k_row(l, r, Ks + (size_t)c * 32);
    else if (c < ((m + SC_TILE - 1) / SC_TILE) * SC_TILE) zero_row(Ks + (size_t)c * 32);
    // non-negative floats (inf, NaN included) order like their bit patterns
    unsigned b = __float_as_uint(nrm);
    b = __reduce_max_sync(__activemask(), b);
    if ((threadIdx.x & 31) == 0 && b > *reinterpret_cast<volatile unsigned*>(w + W_KMAX)) atomicMax(reinterpret_cast<unsigned*>(w + W_KMAX), b);
}

} // namespace erp
