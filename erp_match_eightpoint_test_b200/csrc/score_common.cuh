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
__device__ __forceinline__ void write_e_row(const float e[9], float big, float* __restrict__ row, float gain = 1.f)
{
    float v[32];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const float eg = __fmul_rn(e[i], gain);
        const float hi = tf32_rna(eg), lo = tf32_rna(__fsub_rn(eg, hi));
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
    // set once per call by the correspondence prep: max |l||r|, m, tiles, pass-A tiles, max (|l|^2 + |r|^2), max |r|^2 (float bits)
    W_KMAX = 0, W_M = 1, W_NCT = 2, W_N0 = 3, W_RMAX = 4, W_RRMAX = 5,
    W_AMAX = 6 /* 2 words */, W_DONE = 8, W_LENF = 10, W_LSTAR = 11, W_REMAIN = 12,
    W_DYN_A = 16, W_DYN_B = 20, W_DYN_C = 24,                // { rows of the A matrix, first tile, end tile, m }
    W_WORDS = 28, W_CHUNK0 = 6                               // words [W_CHUNK0, W_WORDS) are reset per hypothesis chunk
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
    const float rr = c < m ? r.x * r.x + r.y * r.y + r.z * r.z : 0.f, ll = c < m ? l.x * l.x + l.y * l.y + l.z * l.z : 0.f;
    unsigned b = __float_as_uint(nrm), b1 = __float_as_uint(ll + rr), b2 = __float_as_uint(rr);
    b = __reduce_max_sync(__activemask(), b);
    b1 = __reduce_max_sync(__activemask(), b1);
    b2 = __reduce_max_sync(__activemask(), b2);
    if ((threadIdx.x & 31) == 0) {
        if (b > *reinterpret_cast<volatile unsigned*>(w + W_KMAX)) atomicMax(reinterpret_cast<unsigned*>(w + W_KMAX), b);
        if (b1 > *reinterpret_cast<volatile unsigned*>(w + W_RMAX)) atomicMax(reinterpret_cast<unsigned*>(w + W_RMAX), b1);
        if (b2 > *reinterpret_cast<volatile unsigned*>(w + W_RRMAX)) atomicMax(reinterpret_cast<unsigned*>(w + W_RRMAX), b2);
    }
}

// Sampson and angular residual tests depend on the pair (h, c) through |E r|^2 and |E^T l|^2, which the residual GEMM
// does not produce.  For the BOUND they are replaced by their maximum over unit-ish bearings:
//     Sampson  res^2 < tau^2 (|E r|^2 + |E^T l|^2) <= tau^2 s1^2 (|r|^2 + |l|^2)    =>  |res| < T_h = tau  s1 sqrt(max_c (|l|^2 + |r|^2))
//     angular  res^2 < sin^2(tau) |E r|^2          <= sin^2(tau) s1^2 |r|^2         =>  |res| < T_h = sin(tau) s1 sqrt(max_c |r|^2)
// with s1 the largest singular value of the scaled E.  T_h is a per-HYPOTHESIS threshold: its operand row is scaled by
// g_h = tau / max(T_h, tau) <= 1 and the kernel keeps its single threshold tau (+ band).  The count is then an upper
// bound of the exact count (score_kernel / the oracle) -- looser than for the algebraic test, so more contenders are
// re-scored exactly, but the winner is the same.  (1 + 1e-5) absorbs the fp32 roundings of the exact test.
struct RowScale { int metric = ERP_METRIC_ALGEBRAIC; float tau = 0.f, sin_tau = 0.f; };
__device__ __forceinline__ float row_gain(const float e[9], const RowScale& rs, const int32_t* __restrict__ w)
{
    if (rs.metric == ERP_METRIC_ALGEBRAIC) return 1.f;
    double m[6] = {0, 0, 0, 0, 0, 0};                        // E^T E, upper triangle
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double a = e[3 * i], b = e[3 * i + 1], c = e[3 * i + 2];
        m[0] += a * a; m[1] += a * b; m[2] += a * c; m[3] += b * b; m[4] += b * c; m[5] += c * c;
    }
    const double t = m[0] + m[3] + m[5];
    const double c1 = (m[0] * m[3] - m[1] * m[1]) + (m[0] * m[5] - m[2] * m[2]) + (m[3] * m[5] - m[4] * m[4]);
    const double disc = t * t - 4.0 * c1;
    // largest eigenvalue of a rank-2 E^T E; the Frobenius bound t covers everything else (NaN included)
    double s1sq = 0.5 * (t + sqrt(disc > 0.0 ? disc : 0.0)) * (1.0 + 1e-6) + 1e-9 * t;
    if (!(s1sq < t)) s1sq = t;
    const float R = __uint_as_float((unsigned)w[rs.metric == ERP_METRIC_SAMPSON ? W_RMAX : W_RRMAX]);
    const double T = (double)(rs.metric == ERP_METRIC_SAMPSON ? rs.tau : rs.sin_tau) * sqrt(s1sq * (double)R) * (1.0 + 1e-5);
    if (!(T < (double)INFINITY)) return 0.f;                 // non-finite input: residual 0, always counted
    return T > (double)rs.tau ? (float)((double)rs.tau / T) : 1.f;
}

} // namespace erp
