// score_tc.cu -- best-hypothesis search on the tensor cores.  The H x M residual matrix
//     res[h][c] = sum_k Eh[h][k] * (l_c (x) r_c)[k],  k = 0..8          (src/epipolar_tool.cpp:100-107)
// is a GEMM with a 9-long contraction.  One tcgen05 pass with K = 32 carries the three 3xTF32
// products at once:
//     A row (hypothesis)     = [ Eh_hi(9) | Eh_hi(9) | Eh_lo(9) | 0 x 5 ]
//     B row (correspondence) = [ K_hi(9)  | K_lo(9)  | K_hi(9)  | 0 x 5 ]      (128 bytes = one swizzle row)
// so a 128 x 256 tile of residuals costs four UMMA instructions and never leaves the SM.  The
// epilogue (one hypothesis row per thread, tcgen05.ld 32x32b) counts, per hypothesis,
//     c_hi = #{ |res| < tau + delta }  >=  the exact inlier count of the SIMT kernel / the oracle,
// where delta bounds |res_3xTF32 - res_fp32chain| (it scales with sqrt(2) * max|l||r|): two
// instructions per residual (FSET + FADD).  Then, exactly (score_kernel, score.cu):
//     L* = count of the hypothesis with the largest c_hi           (a lower bound of the best count)
//     winner = best exact count over { h : c_hi[h] >= L* }         (nothing outside can reach L*)
// so the packed best is bit-identical to scoring everything with the fp32 fma chain, at a fraction
// of the FP32-pipe work.
//
// Progressive pruning (exact): the bound is accumulated in three passes over the correspondences.
// Pass A covers the first 2048; its best-looking hypothesis is scored exactly -> L*.  Pass B extends
// every bound until a hypothesis that is still empty could no longer reach L* (about m - L*
// correspondences); hypotheses with  bound so far + correspondences not yet seen < L*  are dropped,
// and pass C finishes only the survivors (typically a few % of H).  All lengths stay on the device:
// the kernel reads its hypothesis count and correspondence-tile range from memory.
//
// Work is cut stream-K style over (pair of hypothesis tiles, correspondence tile) units -- each 32 KB
// correspondence tile that TMA brings in feeds two 128-row accumulators, which halves the L2 -> shared
// memory traffic per MMA (with one it was the limiter: tensor pipe 42 % active).  A pair cut between two
// CTAs adds its partial counts atomically.
#include "score_common.cuh"
#include "tc_common.cuh"

namespace erp {

constexpr int SM_ROWS = 128;                 // hypotheses per tile (UMMA M)
constexpr int SN_ROWS = 256;                 // correspondences per tile (UMMA N)
constexpr int SA_BYTES = SM_ROWS * 128;      // 16 KB
constexpr int SB_BYTES = SN_ROWS * 128;      // 32 KB
constexpr int S_SUB = 2;                     // hypothesis tiles that share every correspondence tile (halves the L2 -> smem feed)
constexpr int S_NSLOT = 5;                   // B ring
constexpr int S_EPI_GROUPS = 2;
constexpr int S_EPI_THREADS = S_EPI_GROUPS * 128;
constexpr int S_THREADS = 128 + S_EPI_THREADS;
constexpr int S_EPI_COLS = SN_ROWS / S_EPI_GROUPS;
constexpr int S_SMEM = 2 * S_SUB * SA_BYTES + S_NSLOT * SB_BYTES + 256;
static_assert(S_SMEM <= TC_SMEM_LIMIT, "shared memory budget");
constexpr uint32_t S_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(SN_ROWS >> 3) << 17) | ((uint32_t)(SM_ROWS >> 4) << 24);
// |res_tc - res_chain| <= S_KAPPA * sqrt(2) * max_c |l_c||r_c|: 3xTF32 (2^-20) + the chain's own
// rounding (9 * 2^-24) with a factor ~8 to spare
constexpr float S_KAPPA = 1.0f / 65536.0f;

// ---- operand preparation ---------------------------------------------------------------
// row i of Es = split scaled E of hypothesis list[i] (i < *list_len) or of hypothesis i
__global__ void prep_e_kernel(const double* __restrict__ E, int H, float* __restrict__ Es /* H x 32 */,
                              const int32_t* __restrict__ list, const int32_t* __restrict__ list_len, float big)
{
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (list) H = min(H, *list_len);
    if (h >= H) return;
    float e[9], hi[9], lo[9];
    scale_E(E + (size_t)(list ? list[h] : h) * 9, e);
#pragma unroll
    for (int i = 0; i < 9; i++) { hi[i] = tf32_rna(e[i]); lo[i] = tf32_rna(__fsub_rn(e[i], hi[i])); }
    float row[32];
#pragma unroll
    for (int i = 0; i < 9; i++) { row[i] = hi[i] * big; row[9 + i] = hi[i] * big; row[18 + i] = lo[i] * big; }   // exact: power of two
#pragma unroll
    for (int i = 27; i < 32; i++) row[i] = 0.f;
    float4* o = reinterpret_cast<float4*>(Es + (size_t)h * 32);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = make_float4(row[4 * i], row[4 * i + 1], row[4 * i + 2], row[4 * i + 3]);
}

__global__ void prep_k_kernel(const float4* __restrict__ l4, const float4* __restrict__ r4, int m,
                              float* __restrict__ Ks /* m x 32 */, unsigned* __restrict__ kmax_bits)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    float nrm = 0.f;
    if (c < m) {
        float4 l = l4[c], r = r4[c];
        float k[9], hi[9], lo[9];
        kron9(l, r, k);
#pragma unroll
        for (int i = 0; i < 9; i++) { hi[i] = tf32_rna(k[i]); lo[i] = tf32_rna(__fsub_rn(k[i], hi[i])); }
        float row[32];
#pragma unroll
        for (int i = 0; i < 9; i++) { row[i] = hi[i]; row[9 + i] = lo[i]; row[18 + i] = hi[i]; }
#pragma unroll
        for (int i = 27; i < 32; i++) row[i] = 0.f;
        float4* o = reinterpret_cast<float4*>(Ks + (size_t)c * 32);
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = make_float4(row[4 * i], row[4 * i + 1], row[4 * i + 2], row[4 * i + 3]);
        nrm = sqrtf((l.x * l.x + l.y * l.y + l.z * l.z) * (r.x * r.x + r.y * r.y + r.z * r.z));
    }
    // non-negative floats (inf, NaN included) order like their bit patterns
    unsigned b = __float_as_uint(nrm);
    b = __reduce_max_sync(0xffffffffu, b);
    if ((threadIdx.x & 31) == 0 && b > *reinterpret_cast<volatile unsigned*>(kmax_bits)) atomicMax(kmax_bits, b);
}

// ---- the scoring kernel ------------------------------------------------------------------
struct ScoreTcParams {
    int m;
    float tau;
    float big;                    // 2^k applied to every A row: tau * big in [2^30, 2^31)
    const unsigned* kmax_bits;
    const int32_t* dyn;           // device: { hypotheses (rows of the A matrix), first correspondence tile, end tile }
    int32_t* upper;               // one counter per A row, zeroed by the caller: partial sums are added
};

// the launch's unit grid, read from device memory by every role
struct ScoreShape {
    int H, ct_begin, n_htiles, n_ctiles, units_per_cta;
    __device__ explicit ScoreShape(const ScoreTcParams& p)
    {
        H = p.dyn[0]; ct_begin = p.dyn[1];
        n_ctiles = max(p.dyn[2] - ct_begin, 0);
        n_htiles = (H + SM_ROWS * S_SUB - 1) / (SM_ROWS * S_SUB);            // tile PAIRS
        long total = (long)n_htiles * n_ctiles;
        units_per_cta = (int)((total + gridDim.x - 1) / gridDim.x);
        if (units_per_cta < 1) units_per_cta = 1;
    }
};

// 32 residuals of one hypothesis row -> 1.0f per residual with |res| < thr, summed pairwise into two packed fp32x2
// accumulators (FADD2).  The indicator comes from two different pipes, alternating by pair, so that neither limits:
//   * FSET.BF.LT |res|, thr                    (ALU pipe, half rate)
//   * FADD.SAT thr, -|res|                     (FMA pipe): residuals and threshold arrive scaled by 2^k with thr * 2^k >= 2^30,
//     so thr - |res| is 0 or at least one ulp >= 128 and the saturation yields exactly 0 or 1 (NaN -> 0, like the compare)
// 1.5 instructions per residual over two pipes instead of 2 with the ALU pipe saturated.  Counts stay exact in fp32 up to 2^24.
__device__ __forceinline__ float ind_set(uint32_t bits, float thr)
{
    float r;
    asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(fabsf(__uint_as_float(bits))), "f"(thr));
    return r;
}
__device__ __forceinline__ float ind_sat(uint32_t bits, float thr)
{
    float r;
    asm("add.rn.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(thr), "f"(-fabsf(__uint_as_float(bits))));
    return r;
}
__device__ __forceinline__ void add_pair(unsigned long long& acc, float a, float b)
{
    unsigned long long p;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a), "f"(b));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(p));
}
__device__ __forceinline__ void count_chunk(const uint32_t (&v)[32], float hi, unsigned long long (&acc)[2])
{
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        add_pair(acc[0], ind_sat(v[j], hi), ind_sat(v[j + 1], hi));
        add_pair(acc[1], ind_set(v[j + 2], hi), ind_set(v[j + 3], hi));
    }
}
__device__ __forceinline__ int pair_total(const unsigned long long (&acc)[2])
{
    float a, b, c, d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[0]));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c), "=f"(d) : "l"(acc[1]));
    return (int)((a + b) + (c + d));
}

__global__ void __launch_bounds__(S_THREADS, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap map_e, const __grid_constant__ CUtensorMap map_k, const ScoreTcParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* a_smem = smem;                               // [2 buffers][S_SUB] hypothesis tiles
    uint8_t* b_smem = smem + 2 * S_SUB * SA_BYTES;        // [S_NSLOT] correspondence tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + S_NSLOT * SB_BYTES);
    uint64_t* full = bars;                  // [S_NSLOT]
    uint64_t* empty = bars + S_NSLOT;       // [S_NSLOT]
    uint64_t* afull = bars + 2 * S_NSLOT;   // [2]
    uint64_t* aempty = afull + 2;           // [2]
    uint64_t* tfull = afull + 4;            // [2]
    uint64_t* tempty = afull + 6;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull + 8);

    // roles: warps 0..7 epilogue, then TMEM allocator, idle, TMA producer, MMA issuer.  The scheduler
    // favours the highest warp id among eligible warps: the two single-thread feeders sit on top so
    // that the issue-bound epilogue never delays a TMA or an MMA.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int W_ALLOC = S_EPI_THREADS / 32, W_TMA = W_ALLOC + 2, W_MMA = W_ALLOC + 3;

    if (warp == W_TMA && lane == 0) { tma_prefetch_desc(&map_e); tma_prefetch_desc(&map_k); }
    if (warp == W_MMA && lane == 0) {
        for (int i = 0; i < S_NSLOT; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; i++) {
            mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1);
            mbar_init(&tfull[i], 1); mbar_init(&tempty[i], S_EPI_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_ALLOC) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // the four feeder warps need few registers; the eight epilogue warps take them over (128 accumulator columns in flight)
    if (warp >= W_ALLOC) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
    if (warp == W_TMA) {
        // ================================================================ TMA producer
        if (lane == 0) {
            uint32_t slot = 0, ph = 0, seg_n = 0;
            const ScoreShape sh(p);
            SegIter it(sh.n_htiles, sh.n_ctiles, sh.units_per_cta, blockIdx.x);
            int ht, c0, c1, seg;
            for (; it.next(ht, c0, c1, seg); seg_n++) {
                c0 += sh.ct_begin; c1 += sh.ct_begin;
                const uint32_t ab = seg_n & 1;
                mbar_wait(&aempty[ab], ((seg_n >> 1) & 1) ^ 1);
                mbar_expect_tx(&afull[ab], S_SUB * SA_BYTES);
#pragma unroll
                for (int sub = 0; sub < S_SUB; sub++)
                    tma_load_2d(&map_e, &afull[ab], a_smem + (ab * S_SUB + sub) * SA_BYTES, 0, (ht * S_SUB + sub) * SM_ROWS);
                for (int ct = c0; ct < c1; ct++) {
                    mbar_wait(&empty[slot], ph ^ 1);
                    mbar_expect_tx(&full[slot], SB_BYTES);
                    tma_load_2d(&map_k, &full[slot], b_smem + slot * SB_BYTES, 0, ct * SN_ROWS);
                    if (++slot == S_NSLOT) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == W_MMA) {
        // ================================================================ MMA issuer
        if (lane == 0) {
            uint32_t slot = 0, ph = 0, seg_n = 0, tile_n = 0;
            const uint32_t a_base = smem_u32(a_smem), b_base = smem_u32(b_smem);
            const ScoreShape sh(p);
            SegIter it(sh.n_htiles, sh.n_ctiles, sh.units_per_cta, blockIdx.x);
            int ht, c0, c1, seg;
            for (; it.next(ht, c0, c1, seg); seg_n++) {
                const uint32_t ab = seg_n & 1;
                mbar_wait(&afull[ab], (seg_n >> 1) & 1);
                for (int ct = c0; ct < c1; ct++) {
                    mbar_wait(&full[slot], ph);
                    const uint32_t b = b_base + slot * SB_BYTES;
#pragma unroll
                    for (int sub = 0; sub < S_SUB; sub++, tile_n++) {
                        const uint32_t acc = tile_n & 1;
                        const uint32_t a = a_base + (ab * S_SUB + sub) * SA_BYTES;
                        mbar_wait(&tempty[acc], ((tile_n >> 1) & 1) ^ 1);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + acc * SN_ROWS;
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            tc_mma_tf32(d_tmem, smem_desc_sw128(a + k * 32), smem_desc_sw128(b + k * 32), S_IDESC, k != 0);
                        tc_commit(&tfull[acc]);
                    }
                    tc_commit(&empty[slot]);
                    if (++slot == S_NSLOT) { slot = 0; ph ^= 1; }
                }
                tc_commit(&aempty[ab]);
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;" ::: "memory");
        // ================================================================ epilogue
        const int ew = warp & 3, cg = warp >> 2;
        const int row = ew * 32 + lane;
        // band around tau inside which the tensor-core value cannot decide the test
        const float kmax = __uint_as_float(*p.kmax_bits);
        const float delta = S_KAPPA * 1.41421356f * kmax;
        float hi = (p.tau + delta) * p.big;                             // the A operand carries the factor p.big (a power of two)
        if (!(delta < INFINITY)) hi = INFINITY;                         // non-finite input: every finite residual may be an inlier
        // the threshold is re-read (volatile) after every hand-back of an accumulator: the counting depends on that load, so
        // the assembler cannot sink the TMEM loads into the counting code and delay the hand-back
        volatile float* hi_sh = reinterpret_cast<volatile float*>(reinterpret_cast<uint8_t*>(bars) + 192);
        *hi_sh = hi;
        uint32_t tile_n = 0;
        const ScoreShape sh(p);
        SegIter it(sh.n_htiles, sh.n_ctiles, sh.units_per_cta, blockIdx.x);
        int ht, c0, c1, seg;
        while (it.next(ht, c0, c1, seg)) {
            c0 += sh.ct_begin; c1 += sh.ct_begin;
            unsigned long long acc4[S_SUB][2];
#pragma unroll
            for (int sub = 0; sub < S_SUB; sub++) acc4[sub][0] = acc4[sub][1] = 0ull;
            int pad_total = 0;
            // Software pipeline over the tiles of the segment: while a tile is counted out of registers, the next tile's
            // accumulator is read into the buffers that have just been consumed (first two chunks after the first half of
            // the counting, the other two after the second half) and handed back -- the TMEM read latency and the wait
            // for the MMA hide behind the counting instead of adding to it.
            static_assert(S_EPI_COLS == 128 && S_SUB == 2, "four chunks of 32 columns per warp and tile, two tiles per round");
            uint32_t va[32], vb[32], vc[32], vd[32];
            const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16) + cg * S_EPI_COLS;
            if (c0 < c1) {
                const uint32_t acc = tile_n & 1;
                mbar_wait(&tfull[acc], (tile_n >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = lane_base + acc * SN_ROWS;
                tc_ld32(taddr, va); tc_ld32(taddr + 32, vb); tc_ld32(taddr + 64, vc); tc_ld32(taddr + 96, vd);
                tc_wait_ld32(va); tc_wait_ld32(vb); tc_wait_ld32(vc); tc_wait_ld32(vd);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                tile_n++;
            }
            for (int ct = c0; ct < c1; ct++) {
                const int colbase = ct * SN_ROWS + cg * S_EPI_COLS;
#pragma unroll
                for (int sub = 0; sub < S_SUB; sub++) {
                    // registers hold tile (ct, sub); tile_n already names the next one
                    const bool has_next = sub + 1 < S_SUB || ct + 1 < c1;
                    const uint32_t acc = tile_n & 1;
                    const uint32_t taddr = lane_base + acc * SN_ROWS;
                    const float hv = *hi_sh;
                    count_chunk(va, hv, acc4[sub]);
                    count_chunk(vb, hv, acc4[sub]);
                    if (has_next) {
                        mbar_wait(&tfull[acc], (tile_n >> 1) & 1);
                        tc_fence_after();
                        tc_ld32(taddr, va); tc_ld32(taddr + 32, vb);
                    }
                    count_chunk(vc, hv, acc4[sub]);
                    count_chunk(vd, hv, acc4[sub]);
                    if (has_next) {
                        tc_ld32(taddr + 64, vc); tc_ld32(taddr + 96, vd);
                        tc_wait_ld32(va); tc_wait_ld32(vb); tc_wait_ld32(vc); tc_wait_ld32(vd);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[acc]);
                        tile_n++;
                    }
                }
                // zero-filled columns past m have res = 0 exactly and were counted
                pad_total += S_EPI_COLS - min(max(p.m - colbase, 0), S_EPI_COLS);
            }
#pragma unroll
            for (int sub = 0; sub < S_SUB; sub++) {
                const int h = (ht * S_SUB + sub) * SM_ROWS + row;
                if (h < sh.H) {
                    int n = pair_total(acc4[sub]);
                    if (0.f < hi) n -= pad_total;
                    if (n) atomicAdd(p.upper + h, n);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_ALLOC) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// packed (bound << 32 | ~h) maximum: the hypothesis with the largest bound, lowest index on ties
__global__ void upper_argmax_kernel(const int32_t* __restrict__ upper, int H, unsigned long long* __restrict__ best)
{
    unsigned long long b = 0;
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < H; h += gridDim.x * blockDim.x) {
        unsigned long long v = ((unsigned long long)(uint32_t)upper[h] << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)h);
        b = v > b ? v : b;
    }
    for (int o = 16; o > 0; o >>= 1) { unsigned long long y = __shfl_down_sync(0xffffffffu, b, o); b = y > b ? y : b; }
    if ((threadIdx.x & 31) == 0) atomicMax(best, b);
}

// device words shared by the passes (int32 view of the misc scratch)
enum { W_KMAX = 0, W_LEN1 = 2, W_LENF = 3, W_AMAX = 4 /* 2 words */, W_DYN_A = 8, W_DYN_B = 12, W_DYN_C = 16, W_LSTAR = 20,
       W_REMAIN = 21, W_WORDS = 24 };

__global__ void set_dyn_kernel(int32_t* __restrict__ dyn, int H, int ct_begin, int ct_end)
{
    dyn[0] = H; dyn[1] = ct_begin; dyn[2] = ct_end;
}

__global__ void first_candidate_kernel(const unsigned long long* __restrict__ best, int32_t* __restrict__ list, int32_t* __restrict__ len)
{
    list[0] = (int32_t)(0xFFFFFFFFu - (uint32_t)(*best & 0xFFFFFFFFull));
    *len = 1;
}

// after pass A: L* is known (exact count of the best-looking hypothesis).  Pass B extends the bounds
// to tile ct1: far enough that a hypothesis with nothing so far can no longer reach L*.
__global__ void plan_passes_kernel(int32_t* __restrict__ w, const int32_t* __restrict__ exact_first, int H, int m, int n0, int n_ct)
{
    const int lstar = *exact_first;
    w[W_LSTAR] = lstar;
    // correspondences a discarded hypothesis may still be missing: m - seen < L*  <=>  seen > m - L*
    long need = (long)m - lstar;
    need += need / 8 + 2 * SN_ROWS;                        // slack: bad hypotheses still collect a few inliers
    int ct1 = (int)((need + SN_ROWS - 1) / SN_ROWS);
    if (ct1 < n0) ct1 = n0;
    if (ct1 > n_ct || lstar * 4 < m) ct1 = n_ct;           // weak best model: pruning cannot pay, finish in pass B
    w[W_DYN_B] = H; w[W_DYN_B + 1] = n0; w[W_DYN_B + 2] = ct1;
    w[W_DYN_C] = 0; w[W_DYN_C + 1] = ct1; w[W_DYN_C + 2] = n_ct;           // [0] = survivors, counted by the select kernel
    long seen = (long)ct1 * SN_ROWS;
    w[W_REMAIN] = seen >= m ? 0 : (int)(m - seen);
}

// survivors of pass B: bound so far + correspondences not yet seen >= L*.  One atomic per block (the per-warp version
// spent its 24 us on ~27k atomics to one word); launched with 256 threads.
__global__ void __launch_bounds__(256)
survivor_select_kernel(const int32_t* __restrict__ upper, int H, int32_t* __restrict__ w, int32_t* __restrict__ list)
{
    __shared__ int wcount[8], wbase[8];
    const int h = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool keep = h < H && upper[h] + w[W_REMAIN] >= w[W_LSTAR];
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wcount[wid] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
        for (int i = 0; i < 8; i++) { wbase[i] = total; total += wcount[i]; }
        const int base = total ? atomicAdd(&w[W_DYN_C], total) : 0;
        for (int i = 0; i < 8; i++) wbase[i] += base;
    }
    __syncthreads();
    if (keep) list[wbase[wid] + __popc(bal & ((1u << lane) - 1))] = h;
}

// contenders: survivors whose completed bound still reaches L*
__global__ void final_select_kernel(const int32_t* __restrict__ upper, const int32_t* __restrict__ upper2,
                                    const int32_t* __restrict__ survivors, int32_t* __restrict__ w, int32_t* __restrict__ list)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    int h = s < w[W_DYN_C] ? survivors[s] : -1;
    bool keep = h >= 0 && upper[h] + upper2[s] >= w[W_LSTAR];
    unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (!bal) return;
    int lane = threadIdx.x & 31, base = 0;
    if (lane == 0) base = atomicAdd(&w[W_LENF], __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) list[base + __popc(bal & ((1u << lane) - 1))] = h;
}

bool score_tc_preferred(int H, int m) { return (double)H * (double)m >= 3.0e7 && m >= 1 && m < (1 << 24) && H >= 1; }

static int launch_score_tc(erp_ctx* ctx, const CUtensorMap& me, const CUtensorMap& mk, const ScoreTcParams& p)
{
    static bool configured = false;
    if (!configured) {
        ERP_CUDA(cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM));
        configured = true;
    }
    cudaEvent_t e0, e1;
    ERP_TRY(score_event(ctx, &e0));
    score_tc_kernel<<<ctx->sm_count, S_THREADS, S_SMEM, ctx->stream>>>(me, mk, p);
    ERP_LAUNCH(ctx, "score_tc_kernel");
    ERP_TRY(score_event(ctx, &e1));
    return ERP_OK;
}

// merges into *d_best the packed (count << 32 | ~id) of the best of the H hypotheses, exactly as
// erp_score_dev + best_kernel would (algebraic residual)
int score_tc_best(erp_ctx* ctx, const double* d_E, int H, const float* d_l4, const float* d_r4, int m, float tau,
                  uint64_t hyp0, int32_t* d_counts_scratch, uint64_t* d_best)
{
    if (m >= (1 << 24)) { set_error("score_tc_best: more than 2^24 correspondences"); return ERP_E_LIMIT; }
    // power-of-two factor on the hypothesis operand: the scaled threshold lies in [2^30, 2^31) (see count_chunk)
    int ex = 0;
    frexpf(tau > 1e-30f ? tau : 1e-30f, &ex);            // tau = f * 2^ex, f in [0.5, 1)
    const float big = ldexpf(1.0f, 31 - ex);
    int st = ERP_OK;
    float* Es = ctx->scratch<float>(S_SC_E, (size_t)H * 32, &st);
    float* Es2 = ctx->scratch<float>(S_SC_E2, (size_t)H * 32, &st);
    float* Ks = ctx->scratch<float>(S_SC_K, (size_t)m * 32, &st);
    int32_t* w = ctx->scratch<int32_t>(S_SC_MISC, W_WORDS, &st);
    int32_t* upper = ctx->scratch<int32_t>(S_SC_BOUNDS, (size_t)H * 2, &st);   // [H] all hypotheses, [H] pass C rows
    int32_t* list = ctx->scratch<int32_t>(S_SC_LIST, (size_t)H * 2 + 1, &st);  // [H] survivors, [H] contenders, [1] first
    ERP_TRY(st);
    int32_t *upper2 = upper + H, *survivors = list, *contenders = list + H, *first = list + 2 * (size_t)H;
    const int n_ct = cdiv(m, SN_ROWS), n0 = n_ct < 8 ? n_ct : 8;
    ERP_CUDA(cudaMemsetAsync(w, 0, W_WORDS * sizeof(int32_t), ctx->stream));
    set_dyn_kernel<<<1, 1, 0, ctx->stream>>>(w + W_DYN_A, H, 0, n0);
    ERP_LAUNCH(ctx, "set_dyn_kernel");
    ERP_CUDA(cudaMemsetAsync(upper, 0, sizeof(int32_t) * (size_t)H * 2, ctx->stream));
    prep_e_kernel<<<cdiv(H, 256), 256, 0, ctx->stream>>>(d_E, H, Es, nullptr, nullptr, big);
    ERP_LAUNCH(ctx, "prep_e_kernel");
    prep_k_kernel<<<cdiv(m, 256), 256, 0, ctx->stream>>>((const float4*)d_l4, (const float4*)d_r4, m, Ks, (unsigned*)(w + W_KMAX));
    ERP_LAUNCH(ctx, "prep_k_kernel");
    CUtensorMap me, me2, mk;
    ERP_TRY(make_map(&me, Es, H, 32, SM_ROWS));
    ERP_TRY(make_map(&me2, Es2, H, 32, SM_ROWS));
    ERP_TRY(make_map(&mk, Ks, m, 32, SN_ROWS));
    ScoreTcParams p;
    p.m = m; p.tau = tau; p.big = big; p.kmax_bits = (const unsigned*)(w + W_KMAX);

    // pass A: every hypothesis, the first n0 correspondence tiles
    p.dyn = w + W_DYN_A; p.upper = upper;
    ERP_TRY(launch_score_tc(ctx, me, mk, p));
    unsigned long long* amax = reinterpret_cast<unsigned long long*>(w + W_AMAX);
    upper_argmax_kernel<<<min(cdiv(H, 256), ctx->sm_count * 4), 256, 0, ctx->stream>>>(upper, H, amax);
    ERP_LAUNCH(ctx, "upper_argmax_kernel");
    first_candidate_kernel<<<1, 1, 0, ctx->stream>>>(amax, first, w + W_LEN1);
    ERP_LAUNCH(ctx, "first_candidate_kernel");
    // exact count of the most promising hypothesis over ALL correspondences: counts_scratch[0] = L*
    ERP_TRY(score_list_best(ctx, d_E, 1, first, w + W_LEN1, d_l4, d_r4, m, tau, hyp0, d_counts_scratch, d_best));
    plan_passes_kernel<<<1, 1, 0, ctx->stream>>>(w, d_counts_scratch, H, m, n0, n_ct);
    ERP_LAUNCH(ctx, "plan_passes_kernel");

    // pass B: every hypothesis, tiles [n0, ct1)
    p.dyn = w + W_DYN_B;
    ERP_TRY(launch_score_tc(ctx, me, mk, p));
    survivor_select_kernel<<<cdiv(H, 256), 256, 0, ctx->stream>>>(upper, H, w, survivors);
    ERP_LAUNCH(ctx, "survivor_select_kernel");

    // pass C: the survivors, tiles [ct1, n_ct)
    prep_e_kernel<<<cdiv(H, 256), 256, 0, ctx->stream>>>(d_E, H, Es2, survivors, w + W_DYN_C, big);
    ERP_LAUNCH(ctx, "prep_e_kernel(survivors)");
    p.dyn = w + W_DYN_C; p.upper = upper2;
    ERP_TRY(launch_score_tc(ctx, me2, mk, p));
    final_select_kernel<<<cdiv(H, 256), 256, 0, ctx->stream>>>(upper, upper2, survivors, w, contenders);
    ERP_LAUNCH(ctx, "final_select_kernel");
    ctx->sc_misc_dev = w;
    return score_list_best(ctx, d_E, H, contenders, w + W_LENF, d_l4, d_r4, m, tau, hyp0, d_counts_scratch, d_best);
}

} // namespace erp
