// tc_common.cuh -- inline-PTX wrappers (mbarrier, TMA, tcgen05 / TMEM) and the tensor-map encoder
// shared by the two tcgen05 kernels (knn_tc.cu: distances, score_tc.cu: residuals).
#pragma once

#include "common.cuh"

#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through the runtime)

namespace erp {

constexpr int TC_SMEM_LIMIT = 232448;   // 227 KB of dynamic shared memory per CTA

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, not as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, 128 x 256 x 8, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with A read from tensor memory (lane = row, one tf32 per column): only B is fetched from shared memory
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared memory -> tensor memory, 128 rows x 32 bytes (8 columns) of the matrix a K-major descriptor names; asynchronous,
// ordered with the MMAs the same thread issues afterwards
__device__ __forceinline__ void tc_cp_128x256b(uint32_t taddr, uint64_t s_desc)
{
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(s_desc) : "memory");
}
// 32 lanes x 32 consecutive columns: thread i of the warp receives lane (base + i)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, 128-byte swizzle, rows of 128 bytes, 8-row groups of
// 1024 bytes (SBO), descriptor version 1 (sm_100).  addr may point inside the first swizzle row
// (k advance of 32 bytes per UMMA K step).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                          // version
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// K-major rows of 32 bytes (ONE tf32 k step), 32-byte swizzle: 8-row groups are 256 bytes apart
__device__ __forceinline__ uint64_t smem_desc_sw32(uint32_t addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;                          // SWIZZLE_32B
    return d;
}
// instruction descriptor: D fp32, A/B tf32, both K-major, N = 256, M = 128


// wait for this thread's outstanding tcgen05.ld; the registers are operands so that no use of
// them can be scheduled above the wait
__device__ __forceinline__ void tc_wait_ld32(uint32_t (&v)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tc_wait_ld8(uint32_t (&v)[8])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])
                 :
                 : "memory");
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn lookup_encode_fn()
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        return reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
    return nullptr;
}
inline EncodeTiledFn encode_fn()
{
    static const EncodeTiledFn fn = lookup_encode_fn();     // C++11 magic static: initialised once, thread safe
    return fn;
}

// rows x row_floats fp32, row major; box = 32 floats (one 128-byte swizzle row) x box_rows, zero fill out of bounds;
// narrow = true: box = 8 floats (one 32-byte swizzle row) x box_rows for the single-k-step operands
inline int make_map(CUtensorMap* map, const float* base, int rows, int row_floats, int box_rows, bool narrow = false)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return ERP_E_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)row_floats, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_floats * sizeof(float)};
    cuuint32_t box[2] = {narrow ? 8u : 32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, narrow ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return ERP_E_CUDA; }
    return ERP_OK;
}


// ---- running top-4 of one query row, shared by the two distance kernels -----------------------
// A column is a candidate iff its approximate score is below
//     min( 4th best so far,  2nd best so far + slack ),      slack = 2 * kappa * (|q|^2 + max|t|^2).
// Columns above "2nd best + 2 eps" can never be one of the exact two nearest neighbours (their exact
// distance exceeds the exact distance of the approximate runner-up), so they need not be kept: the
// rare insert path then runs ~2 ln(n) times per row instead of 4 ln(n) or 8 ln(n), and the list's
// certificate bound is that same minimum (both terms only decrease while the scan proceeds).
__device__ __forceinline__ float slack_of(float qn, float tn_max, float kappa)
{
    return __fmul_rn(__fmul_rn(2.0002f * kappa, __fadd_rn(qn, tn_max)), 1.0f);
}
// sorted insert for any K (ascending; strict <: equal scores keep the earlier index); all stages are independent
template <int K>
__device__ __forceinline__ void topk_insert(float s, int idx, float (&bs)[K], int (&bi)[K])
{
#pragma unroll
    for (int j = K - 1; j > 0; j--) {
        const bool up = s < bs[j - 1], here = s < bs[j];
        bs[j] = up ? bs[j - 1] : (here ? s : bs[j]);
        bi[j] = up ? bi[j - 1] : (here ? idx : bi[j]);
    }
    const bool first = s < bs[0];
    bs[0] = first ? s : bs[0];
    bi[0] = first ? idx : bi[0];
}
__device__ __forceinline__ void top4_insert(float s, int idx, float (&bs)[4], int (&bi)[4])
{
    // strict <: equal scores keep the earlier (lower) train index
    if (s < bs[2]) {
        bs[3] = bs[2]; bi[3] = bi[2];
        if (s < bs[1]) {
            bs[2] = bs[1]; bi[2] = bi[1];
            if (s < bs[0]) { bs[1] = bs[0]; bi[1] = bi[0]; bs[0] = s; bi[0] = idx; }
            else { bs[1] = s; bi[1] = idx; }
        } else { bs[2] = s; bi[2] = idx; }
    } else { bs[3] = s; bi[3] = idx; }
}

// 32 accumulator columns of one query row.  Common case: s = acc + |t|^2 (32 independent adds), a min
// tree and ONE warp-uniform threshold test.  Rare case: only the groups of 8 columns that some lane
// needs are re-read from TMEM in a rolled loop, and only the columns below the threshold are visited
// -- the insert cascade exists once per call site, which keeps the loop inside the instruction cache
// (fully unrolled it was 60 KB and 3x slower).  Re-computed scores are bit-identical to the first pass.
template <int K>
__device__ __forceinline__ void scan_chunk(const uint32_t (&v)[32], uint32_t taddr, const float4* __restrict__ tn4, int col0,
                                           float slack, float (&bs)[K], int (&bi)[K])
{
    float gm[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        float4 na = tn4[g * 2], nb = tn4[g * 2 + 1];
        float s0 = __fadd_rn(__uint_as_float(v[g * 8 + 0]), na.x), s1 = __fadd_rn(__uint_as_float(v[g * 8 + 1]), na.y);
        float s2 = __fadd_rn(__uint_as_float(v[g * 8 + 2]), na.z), s3 = __fadd_rn(__uint_as_float(v[g * 8 + 3]), na.w);
        float s4 = __fadd_rn(__uint_as_float(v[g * 8 + 4]), nb.x), s5 = __fadd_rn(__uint_as_float(v[g * 8 + 5]), nb.y);
        float s6 = __fadd_rn(__uint_as_float(v[g * 8 + 6]), nb.z), s7 = __fadd_rn(__uint_as_float(v[g * 8 + 7]), nb.w);
        gm[g] = fminf(fminf(fminf(s0, s1), fminf(s2, s3)), fminf(fminf(s4, s5), fminf(s6, s7)));
    }
    const float thr = fminf(bs[K - 1], __fadd_rn(bs[1], slack));
    const unsigned mine = (gm[0] < thr ? 1u : 0u) | (gm[1] < thr ? 2u : 0u) | (gm[2] < thr ? 4u : 0u) | (gm[3] < thr ? 8u : 0u);
    unsigned need = __reduce_or_sync(0xffffffffu, mine);
    while (need) {                                  // warp-uniform
        const int g = __ffs(need) - 1;
        need &= need - 1;
        uint32_t w[8];
        tc_ld8(taddr + g * 8, w);
        tc_wait_ld8(w);
        const float4 na = tn4[g * 2], nb = tn4[g * 2 + 1];
        if ((mine >> g) & 1u) {
            float e[8];
            e[0] = __fadd_rn(__uint_as_float(w[0]), na.x); e[1] = __fadd_rn(__uint_as_float(w[1]), na.y);
            e[2] = __fadd_rn(__uint_as_float(w[2]), na.z); e[3] = __fadd_rn(__uint_as_float(w[3]), na.w);
            e[4] = __fadd_rn(__uint_as_float(w[4]), nb.x); e[5] = __fadd_rn(__uint_as_float(w[5]), nb.y);
            e[6] = __fadd_rn(__uint_as_float(w[6]), nb.z); e[7] = __fadd_rn(__uint_as_float(w[7]), nb.w);
            unsigned m = (e[0] < thr ? 1u : 0u) | (e[1] < thr ? 2u : 0u) | (e[2] < thr ? 4u : 0u) | (e[3] < thr ? 8u : 0u) |
                         (e[4] < thr ? 16u : 0u) | (e[5] < thr ? 32u : 0u) | (e[6] < thr ? 64u : 0u) | (e[7] < thr ? 128u : 0u);
            while (m) {                             // usually a single column
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const float x = j == 0 ? e[0] : j == 1 ? e[1] : j == 2 ? e[2] : j == 3 ? e[3] : j == 4 ? e[4] : j == 5 ? e[5] : j == 6 ? e[6] : e[7];
                if (x < fminf(bs[K - 1], __fadd_rn(bs[1], slack))) {
                    if (K == 4) top4_insert(x, col0 + g * 8 + j, reinterpret_cast<float (&)[4]>(bs), reinterpret_cast<int (&)[4]>(bi));
                    else topk_insert<K>(x, col0 + g * 8 + j, bs, bi);
                }
            }
        }
        __syncwarp();
    }
}

// Candidate list of the 1xTF32 engine, 8 entries per (query row, list): [0], [1] = the two best scores so far in
// order (all the threshold needs), [2..7] = a FIFO of the other columns that were within the slack of the runner-up when
// they arrived (newest first).  An entry pushed out of the FIFO lowers floor_thr to its score (and with it the threshold),
// every other dropped column scored >= the threshold in force, so min(floor_thr, final threshold) bounds all
// non-candidates of the list.
// Scores carry their column (mod 8) in the three low mantissa bits -- the minimum of the packed keys is the winning
// score AND its position, which replaces a dynamic register index; the 2^-21 relative perturbation is part of the margin
// in F_KAPPA.  refine_kernel takes the list in any order.
//
// scan_chunk_lean: 32 accumulator columns of one query row, all in registers.  Common case: a min tree and ONE compare.
// Rare case (per lane, no TMEM access, no warp-wide vote): the groups of 8 columns whose minimum is below the threshold are
// selected out of the registers behind an opaque group id (so the compiler cannot specialise the code per group), packed,
// and inserted smallest first.  With SHARE the threshold is also exchanged through shared memory with the thread that scans
// the other column range of the same row; any published value is a valid bound, so a lost race only costs pruning.
// With FOLD the accumulator already holds |t|^2 - 2 q.t (the norm entered the MMA as an extra k step).
#if defined(ERP_TC_COUNTERS) && ERP_TC_COUNTERS == 2
static __device__ unsigned long long erp_evt[8];
#endif
template <bool SHARE, bool FOLD>
__device__ __forceinline__ void scan_chunk_lean(const uint32_t (&v)[32], const float4* __restrict__ tn4, int col0, float slack,
                                                float (&bs)[8], int (&bi)[8], float* row_thr, float& floor_thr)
{
    const float other = SHARE ? *reinterpret_cast<volatile float*>(row_thr) : INFINITY;
    float gm[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        float s[8];
        if (FOLD) {
#pragma unroll
            for (int j = 0; j < 8; j++) s[j] = __uint_as_float(v[g * 8 + j]);
        } else {
            const float4 na = tn4[g * 2], nb = tn4[g * 2 + 1];
            const float nrm[8] = {na.x, na.y, na.z, na.w, nb.x, nb.y, nb.z, nb.w};
#pragma unroll
            for (int j = 0; j < 8; j++) s[j] = __fadd_rn(__uint_as_float(v[g * 8 + j]), nrm[j]);
        }
        gm[g] = fminf(fminf(fminf(s[0], s[1]), fminf(s[2], s[3])), fminf(fminf(s[4], s[5]), fminf(s[6], s[7])));
    }
    // floor_thr in the threshold is the capacity rule: once an entry has been pushed out, columns at or above its score
    // cannot improve the list's bound any more (all-equal scores would otherwise insert at every column)
    float thr = fminf(fminf(__fadd_rn(bs[1], slack), other), floor_thr);
    if (SHARE) floor_thr = thr;
    if (fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3])) < thr) {                                     // rare, per lane
        unsigned mine = (gm[0] < thr ? 1u : 0u) | (gm[1] < thr ? 2u : 0u) | (gm[2] < thr ? 4u : 0u) | (gm[3] < thr ? 8u : 0u);
#if defined(ERP_TC_COUNTERS) && ERP_TC_COUNTERS == 2
        const long long e0 = clock64();
        atomicAdd(&erp_evt[4], 1ull);
#endif
        do {
            int g = __ffs(mine) - 1;
            mine &= mine - 1;
            asm volatile("" : "+r"(g));
            const bool pl = (g & 1) != 0, ph = (g & 2) != 0;
            float nrm[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (!FOLD) {
                const float4 na = tn4[g * 2], nb = tn4[g * 2 + 1];
                nrm[0] = na.x; nrm[1] = na.y; nrm[2] = na.z; nrm[3] = na.w; nrm[4] = nb.x; nrm[5] = nb.y; nrm[6] = nb.z; nrm[7] = nb.w;
            }
#if defined(ERP_TC_COUNTERS) && ERP_TC_COUNTERS == 2
            atomicAdd(&erp_evt[5], 1ull);
#endif
            float k[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t lo = pl ? v[8 + j] : v[j], hi = pl ? v[24 + j] : v[16 + j];
                const float raw = __uint_as_float(ph ? hi : lo);
                const float e = FOLD ? raw : __fadd_rn(raw, nrm[j]);
                k[j] = __uint_as_float((__float_as_uint(e) & ~7u) | (uint32_t)j);
            }
            const int cbase = col0 + g * 8;
            for (;;) {
                const float km = fminf(fminf(fminf(k[0], k[1]), fminf(k[2], k[3])), fminf(fminf(k[4], k[5]), fminf(k[6], k[7])));
                if (!(km < thr)) break;
#if defined(ERP_TC_COUNTERS) && ERP_TC_COUNTERS == 2
                atomicAdd(&erp_evt[6], 1ull);
#endif
                const int col = cbase + (int)(__float_as_uint(km) & 7u);
                // strict <: equal scores keep the earlier (lower) train index
                const bool lt0 = km < bs[0], lt1 = km < bs[1];
                const float pv = lt1 ? bs[1] : km;            // leaves the top two (or never enters)
                const int pi = lt1 ? bi[1] : col;
                bs[1] = lt0 ? bs[0] : (lt1 ? km : bs[1]);
                bi[1] = lt0 ? bi[0] : (lt1 ? col : bi[1]);
                bs[0] = lt0 ? km : bs[0];
                bi[0] = lt0 ? col : bi[0];
                thr = fminf(fminf(__fadd_rn(bs[1], slack), other), floor_thr);
                if (pv < thr) {
                    floor_thr = fminf(floor_thr, bs[7]);     // +inf while the FIFO has room
                    thr = fminf(thr, floor_thr);
#pragma unroll
                    for (int i = 7; i > 2; i--) { bs[i] = bs[i - 1]; bi[i] = bi[i - 1]; }
                    bs[2] = pv; bi[2] = pi;
                }
#pragma unroll
                for (int j = 0; j < 8; j++) k[j] = k[j] == km ? INFINITY : k[j];
            }
        } while (mine);
        if (SHARE) {
            if (thr < *reinterpret_cast<volatile float*>(row_thr)) *reinterpret_cast<volatile float*>(row_thr) = thr;
        }
#if defined(ERP_TC_COUNTERS) && ERP_TC_COUNTERS == 2
        const long long e1 = clock64();
        {
            const unsigned am = __activemask();
            const bool warm = (col0 & 0xffff) < 4096;
            if ((threadIdx.x & 31) == __ffs(am) - 1) {
                atomicAdd(&erp_evt[warm ? 2 : 0], (unsigned long long)(e1 - e0));
                atomicAdd(&erp_evt[warm ? 3 : 1], 1ull);
            }
        }
#endif
    }
}

// A CTA's unit range cut into segments: consecutive train tiles of one query tile.
struct SegIter {
    long u, end;
    int n_ttiles, L;
    __device__ SegIter(int n_qtiles, int n_ttiles_, int units_per_cta, int cta)
    {
        L = units_per_cta; n_ttiles = n_ttiles_;
        u = (long)cta * L;
        long total = (long)n_qtiles * n_ttiles;
        end = u + L < total ? u + L : total;
    }
    // next segment: query tile, train tiles [t0, t1), list slot of the segment within its query tile
    __device__ bool next(int& qtile, int& t0, int& t1, int& seg)
    {
        if (u >= end) return false;
        qtile = (int)(u / n_ttiles);
        long row0 = (long)qtile * n_ttiles;
        t0 = (int)(u - row0);
        long rem = end - u;
        t1 = rem < n_ttiles - t0 ? t0 + (int)rem : n_ttiles;
        seg = (int)(u / L - row0 / L);      // CTA boundaries (multiples of L) in (row0, u]
        u += t1 - t0;
        return true;
    }
};


} // namespace erp
