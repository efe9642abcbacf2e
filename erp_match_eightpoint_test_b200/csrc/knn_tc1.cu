// knn_tc1.cu -- the fast tcgen05 distance engine: ONE TF32 product per tile, a wider candidate list
// and the same exact certificate.
//
// knn_tc.cu reproduces fp32 ranking by arithmetic (3xTF32).  The certificate of refine_kernel makes
// that unnecessary: any approximate score with a known error bound yields the exact 2-NN as long as
//     d2_exact(second) < min_lists(s_kth) + |q|^2 - eps        (otherwise the query is re-scanned)
// holds, and a single TF32 product has a worst-case error of 2^-10 (|q|^2 + |t|^2) (inputs rounded
// to 11 significant bits, products exact, fp32 accumulation).  Only columns within twice that error
// of the running second best are kept (up to 8 per list: near-ties overflow 4), which for SURF-like
// data leaves almost nothing to re-scan, so the kernel issues a third of the tensor work:
//     8 UMMA instructions per 128 x 256 tile instead of 24.
// At that rate the L2 -> shared-memory feed becomes the limiter, so every 64 KB train tile that TMA
// brings in is multiplied against TWO resident query tiles (256 queries), alternating between the
// two TMEM accumulators.  Results are identical to the exact engine by construction.
#include "tc_common.cuh"

namespace erp {

constexpr int F_BM = 128;                 // queries per sub-tile (UMMA M)
constexpr int F_SUB_MAX = 2;              // resident query sub-tiles sharing every train tile (1 when D > 96: smem)
constexpr int F_BN = 256;                 // train rows per tile (UMMA N)
constexpr int F_KC = 32;                  // floats per k chunk (one 128-byte swizzle row)
constexpr int F_QCH = F_BM * F_KC * 4;    // 16 KB
constexpr int F_TCH = F_BN * F_KC * 4;    // 32 KB
constexpr int F_EPI_GROUPS = 2;
constexpr int F_EPI_THREADS = F_EPI_GROUPS * 128;
constexpr int F_THREADS = 128 + F_EPI_THREADS;
constexpr int F_EPI_COLS = F_BN / F_EPI_GROUPS;
constexpr int F_TOPK = 8;
constexpr int F_MAX_SEG = 128 / (F_TOPK * F_EPI_GROUPS);    // up to 8 segments per query pair (refine_kernel: 128 candidate slots)
constexpr int F_AUG_T = F_BN * 32;        // 8 KB: one extra k step (8 floats) per train row, 32-byte swizzle rows
constexpr int F_AUG_Q = F_BM * 32;        // 4 KB: the constant query side of that k step
constexpr int F_BARS = 256;               // mbarriers + TMEM slot
// Instrumented builds (scripts/build_variant.sh N): -DERP_TC_COUNTERS=1 adds clock64 phase timers to the MMA issuer and the
// epilogue warps, =2 counts and times the insert path (tc_common.cuh); both print after every launch.  Off by default.
#ifndef ERP_TC_COUNTERS
#define ERP_TC_COUNTERS 0
#endif
#if ERP_TC_COUNTERS == 1
static __device__ unsigned long long erp_clk[8];
#define TCC_DECL unsigned long long lclk[6] = {0, 0, 0, 0, 0, 0}
#define TCC_CLOCK(name) const long long name = clock64()
#define TCC_ADD(i, v) (lclk[i] += (unsigned long long)(v))
#define TCC_FLUSH() do { if ((threadIdx.x & 31) == 0) for (int i_ = 0; i_ < 6; i_++) if (lclk[i_]) atomicAdd(&erp_clk[i_], lclk[i_]); } while (0)
#else
#define TCC_DECL
#define TCC_CLOCK(name)
#define TCC_ADD(i, v)
#define TCC_FLUSH()
#endif
// |s_tc - s_exact| <= 2^-10 (|q|^2 + max|t|^2) in the worst case: both operands are rounded to 11
// significant bits (relative error 2^-11 each), products are exact, and 2|q||t| <= |q|^2 + |t|^2.
// On top of 2^-10: the second-order rounding term (2^-22 relative = 2^-12 of the budget), the column id packed into the
// three low mantissa bits of a score (<= 7 ulp = 2^-21 relative), and the fp32 accumulation inside the tensor pipe,
// whose rounding mode is not documented: with truncating adds over D + 3 <= 131 terms the worst case is 131 * 2^-24
// of the scale = 0.8 % of 2^-10.  Together < 2 %; 5 % is taken.  refine_kernel reports the deviation it observes
// (0.28 of the bound on the synthetic sets; tests/test_gpu_parity.py holds a worst-case-rounding adversarial set).
constexpr double F_KAPPA = 1.05 / 1024.0;
constexpr uint32_t F_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(F_BN >> 3) << 17) | ((uint32_t)(F_BM >> 4) << 24);

__host__ __device__ constexpr int f_sub(int kch) { return kch <= 3 ? F_SUB_MAX : 1; }
// D <= 64: |t|^2 enters the accumulator through the tensor core (an extra k step whose train side is the
// norm split in three tf32 pieces and whose query side is 1,1,1), so the epilogue neither loads nor adds it.
// Wider descriptors keep the norm add: the 20 KB of operand staging would cost them a ring slot they need.
__host__ __device__ constexpr bool f_fold(int kch) { return kch <= 2; }
__host__ __device__ constexpr int f_tail(int kch, bool share)
{
    return (f_fold(kch) ? 2 * F_AUG_T + F_AUG_Q : 2 * F_BN * 4) + F_BARS + (share ? f_sub(kch) * F_BM * 4 : 0);
}
__host__ __device__ constexpr int f_slots_for(int kch, bool share) { return (TC_SMEM_LIMIT - f_sub(kch) * kch * F_QCH - f_tail(kch, share)) / F_TCH; }
// the per-row threshold exchange between the two column groups needs 512 B per sub-tile: on unless it costs the look-ahead slot
__host__ __device__ constexpr bool f_share(int kch) { return f_slots_for(kch, true) >= kch + 1; }
__host__ __device__ constexpr int f_slots(int kch) { return f_slots_for(kch, f_share(kch)); }
__host__ __device__ constexpr int f_smem(int kch) { return f_sub(kch) * kch * F_QCH + f_slots(kch) * F_TCH + f_tail(kch, f_share(kch)); }

// rows -> tf32-rounded rows (dpad floats), optional norms (+inf padding) and their maximum
// LPR lanes own one row (4 floats per lane and pass); a warp covers 32 / LPR rows
// one operand of a prep launch: rows x, scaled, rounded into out; norms (+inf padding up to n_pad), their maximum, and the
// 8-float norm operand rows when aug is given
struct RoundJob {
    const float* x; int n; float scale; float* out; float* norm; int n_pad; unsigned* max_bits; float* aug; int blocks;
};

// both operands of a search in ONE launch: blocks [0, a.blocks) round the queries, the rest the train rows
template <int LPR>
__global__ void __launch_bounds__(256)
round_kernel(const RoundJob a, const RoundJob b, int dim, int dpad)
{
    constexpr int RPW = 32 / LPR;
    const bool second = (int)blockIdx.x >= a.blocks;
    const RoundJob& j = second ? b : a;
    const float* __restrict__ x = j.x;
    float* __restrict__ out = j.out;
    float* __restrict__ norm = j.norm;
    float* __restrict__ aug = j.aug;
    unsigned* __restrict__ max_bits = j.max_bits;
    const int n = j.n, n_pad = j.n_pad;
    const float scale = j.scale;
    const int lane = threadIdx.x & 31, sub = lane / LPR, l = lane % LPR;
    const int row = (((int)blockIdx.x - (second ? a.blocks : 0)) * 8 + (threadIdx.x >> 5)) * RPW + sub;
    double acc = 0.0;
    if (row < n) {
        for (int k = l * 4; k < dpad; k += LPR * 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < dim) v = *reinterpret_cast<const float4*>(x + (size_t)row * dim + k);
            acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
            float4 hi = make_float4(tf32_rna(__fmul_rn(v.x, scale)), tf32_rna(__fmul_rn(v.y, scale)),
                                    tf32_rna(__fmul_rn(v.z, scale)), tf32_rna(__fmul_rn(v.w, scale)));
            *reinterpret_cast<float4*>(out + (size_t)row * dpad + k) = hi;
        }
    }
    if (norm) {
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (l == 0 && row < n_pad) {
            float f = row < n ? (float)acc : INFINITY;
            norm[row] = f;
            if (aug) {
                // |t|^2 = hi + mid + lo exactly, every piece a tf32 number; rows past the end score 1e30
                const float g = row < n ? f : 1e30f;
                const float hi = tf32_rna(g), r1 = __fsub_rn(g, hi), mid = tf32_rna(r1), lo = __fsub_rn(r1, mid);
                *reinterpret_cast<float4*>(aug + (size_t)row * 8) = make_float4(hi, mid, lo, 0.f);
                *reinterpret_cast<float4*>(aug + (size_t)row * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // non-negative floats (and +inf, NaN) order like their bit patterns
            // one hot word for the whole array: read it first, the atomic only fires while the maximum still grows
            if (row < n && __float_as_uint(f) > *reinterpret_cast<volatile unsigned*>(max_bits)) atomicMax(max_bits, __float_as_uint(f));
        }
    }
}

static int launch_prep(erp_ctx* ctx, RoundJob a, RoundJob b, int dim, int dpad)
{
    // 8 / 16 / 32 lanes per row for dpad = 32 / 64 / (96, 128)
    const int lpr = dpad <= 32 ? 8 : dpad <= 64 ? 16 : 32;
    const int rows_per_block = 8 * (32 / lpr);
    a.blocks = cdiv(a.n_pad, rows_per_block);
    b.blocks = b.x ? cdiv(b.n_pad, rows_per_block) : 0;
    const int grid = a.blocks + b.blocks;
    if (lpr == 8) round_kernel<8><<<grid, 256, 0, ctx->stream>>>(a, b, dim, dpad);
    else if (lpr == 16) round_kernel<16><<<grid, 256, 0, ctx->stream>>>(a, b, dim, dpad);
    else round_kernel<32><<<grid, 256, 0, ctx->stream>>>(a, b, dim, dpad);
    ERP_LAUNCH(ctx, "round_kernel");
    return ERP_OK;
}

struct Tc1Params {
    int nq, nt;
    int n_qpairs, n_ttiles;   // n_qpairs counts groups of f_sub(kch) query sub-tiles
    int units_per_cta;        // CTA b owns units [b*L, (b+1)*L) of the n_qpairs x n_ttiles grid
    int n_seg;                // candidate lists per query = n_seg x F_EPI_GROUPS
    const float* tn;
    const float* qn;          // nq query norms |q|^2
    const unsigned* tn_max_bits;
    int32_t* cand_idx;        // nq x lists x F_TOPK, pre-set to -1
    float* cand_s;
    float* cand_thr;          // nq x lists: lower bound of the approximate score of every non-candidate
    int32_t* row_thr;         // nq ordered-int keys (or null): the lowest threshold any segment of the row has published
};

// monotone float <-> int key (scores may be negative), so that atomicMin orders thresholds
__device__ __forceinline__ int32_t thr_key(float f) { const int32_t k = __float_as_int(f); return k >= 0 ? k : k ^ 0x7fffffff; }
__device__ __forceinline__ float key_thr(int32_t k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

template <int KCH>   // k chunks of 32 floats: dpad = 32 * KCH
__global__ void __launch_bounds__(F_THREADS, 1)
knn2_tc1_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t,
                const __grid_constant__ CUtensorMap map_aug, const Tc1Params p)
{
    constexpr int NS = f_slots(KCH), F_SUB = f_sub(KCH);
    constexpr bool FOLD = f_fold(KCH), SHARE = f_share(KCH);
    static_assert(NS >= KCH + 1, "T ring must hold one tile plus a chunk of look-ahead");
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* q_smem = smem;                                   // [F_SUB][KCH] chunks
    uint8_t* t_smem = smem + F_SUB * KCH * F_QCH;             // NS slots
    uint8_t* aug_t = t_smem + NS * F_TCH;                     // FOLD: [2][F_AUG_T] by train-tile parity, then the query side
    uint8_t* aug_q = aug_t + 2 * F_AUG_T;
    float* tn_smem = reinterpret_cast<float*>(t_smem + NS * F_TCH);   // !FOLD: [2][F_BN] |t|^2, by train-tile parity
    uint64_t* bars = reinterpret_cast<uint64_t*>(t_smem + NS * F_TCH + (FOLD ? 2 * F_AUG_T + F_AUG_Q : 2 * F_BN * 4));
    uint64_t* full = bars;              // [NS]  TMA -> MMA
    uint64_t* empty = bars + NS;        // [NS]  MMA -> TMA
    uint64_t* qfull = bars + 2 * NS;    // query pair landed
    uint64_t* qempty = qfull + 1;       // query pair no longer read
    uint64_t* tfull = qfull + 2;        // [2] accumulator ready
    uint64_t* tempty = qfull + 4;       // [2] accumulator drained
    uint64_t* nfull = qfull + 6;        // [2] |t|^2 of a train tile landed
    uint64_t* nempty = qfull + 8;       // [2] ... and consumed by both sub-tiles (FOLD: by their MMAs)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qfull + 10);
    static_assert((2 * NS + 10) * 8 + 4 <= F_BARS, "barrier block");
    float* thr_sh = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + F_BARS);   // SHARE: [F_SUB][F_BM] row thresholds

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TCC_DECL;
    constexpr int W_ALLOC = F_EPI_THREADS / 32, W_NORM = W_ALLOC + 1, W_TMA = W_ALLOC + 2, W_MMA = W_ALLOC + 3;

    if (warp == W_TMA && lane == 0) { tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_t); if (FOLD) tma_prefetch_desc(&map_aug); }
    if (warp == W_MMA && lane == 0) {
        for (int i = 0; i < NS; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(qfull, 1); mbar_init(qempty, 1);
        for (int i = 0; i < 2; i++) {
            mbar_init(&tfull[i], 1); mbar_init(&tempty[i], F_EPI_THREADS / 32);
            mbar_init(&nfull[i], 1); mbar_init(&nempty[i], FOLD ? 1 : F_EPI_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_ALLOC) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (FOLD) {
        // query side of the norm step: (1,1,1,0) in both 16-byte halves of every row, which the swizzle cannot change
        for (int i = threadIdx.x; i < F_AUG_Q / 4; i += F_THREADS) reinterpret_cast<float*>(aug_q)[i] = (i & 3) != 3 ? 1.0f : 0.0f;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // the four feeder warps need few registers; the eight epilogue warps take them over (4 chunks of 32 accumulator columns in flight)
    if (warp >= W_ALLOC) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
    if (warp == W_TMA) {
        // ================================================================ TMA producer
        if (lane == 0) {
            uint32_t slot = 0, ph = 0, seg_n = 0;
            SegIter it(p.n_qpairs, p.n_ttiles, p.units_per_cta, blockIdx.x);
            int qp, t0, t1, seg;
            for (; it.next(qp, t0, t1, seg); seg_n++) {
                mbar_wait(qempty, (seg_n & 1) ^ 1);
                mbar_expect_tx(qfull, F_SUB * KCH * F_QCH);
#pragma unroll
                for (int sub = 0; sub < F_SUB; sub++)
#pragma unroll
                    for (int c = 0; c < KCH; c++)
                        tma_load_2d(&map_q, qfull, q_smem + (sub * KCH + c) * F_QCH, c * F_KC, (qp * F_SUB + sub) * F_BM);
                for (int tt = t0; tt < t1; tt++) {
#pragma unroll
                    for (int c = 0; c < KCH; c++) {
                        mbar_wait(&empty[slot], ph ^ 1);
                        mbar_expect_tx(&full[slot], F_TCH);
                        tma_load_2d(&map_t, &full[slot], t_smem + slot * F_TCH, c * F_KC, tt * F_BN);
                        if (++slot == NS) { slot = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == W_MMA) {
        // ================================================================ MMA issuer
        if (lane == 0) {
            uint32_t slot = 0, ph = 0, seg_n = 0, tile_n = 0;
            const uint32_t q_base = smem_u32(q_smem), t_base = smem_u32(t_smem);
            const uint64_t aq_desc = smem_desc_sw32(smem_u32(aug_q));
            uint32_t tt_n = 0;
            SegIter it(p.n_qpairs, p.n_ttiles, p.units_per_cta, blockIdx.x);
            int qp, t0, t1, seg;
            for (; it.next(qp, t0, t1, seg); seg_n++) {
                mbar_wait(qfull, seg_n & 1);
                for (int tt = t0; tt < t1; tt++, tt_n++) {
                    // the KCH chunks of this train tile occupy KCH consecutive ring slots
                    uint32_t s0 = slot, p0 = ph;
#pragma unroll
                    for (int sub = 0; sub < F_SUB; sub++, tile_n++) {
                        const uint32_t acc = tile_n & 1;
                        TCC_CLOCK(m0);
                        mbar_wait(&tempty[acc], ((tile_n >> 1) & 1) ^ 1);
                        TCC_ADD(3, clock64() - m0); TCC_ADD(4, 1);
                        const uint32_t d_tmem = tmem_base + acc * F_BN;
                        uint32_t sl = s0, pp = p0;
#pragma unroll
                        for (int c = 0; c < KCH; c++) {
                            TCC_CLOCK(f0);
                            if (sub == 0) mbar_wait(&full[sl], pp);
                            TCC_ADD(5, clock64() - f0);
                            tc_fence_after();
                            const uint32_t a = q_base + (sub * KCH + c) * F_QCH, b = t_base + sl * F_TCH;
#pragma unroll
                            for (int k = 0; k < F_KC / 8; k++)
                                tc_mma_tf32(d_tmem, smem_desc_sw128(a + k * 32), smem_desc_sw128(b + k * 32), F_IDESC, (c | k) != 0);
                            if (sub == F_SUB - 1) tc_commit(&empty[sl]);
                            if (++sl == NS) { sl = 0; pp ^= 1; }
                        }
                        if (FOLD) {
                            const uint32_t nb = tt_n & 1;
                            if (sub == 0) mbar_wait(&nfull[nb], (tt_n >> 1) & 1);
                            tc_fence_after();
                            tc_mma_tf32(d_tmem, aq_desc, smem_desc_sw32(smem_u32(aug_t + nb * F_AUG_T)), F_IDESC, 1);
                            if (sub == F_SUB - 1) tc_commit(&nempty[nb]);
                        }
                        tc_commit(&tfull[acc]);
                        if (sub == F_SUB - 1) { slot = sl; ph = pp; }
                    }
                }
                tc_commit(qempty);
            }
        }
    } else if (warp == W_NORM) {
        // ================================================================ |t|^2 producer (one bulk copy per train tile)
        if (lane == 0) {
            uint32_t tt_n = 0;
            SegIter it(p.n_qpairs, p.n_ttiles, p.units_per_cta, blockIdx.x);
            int qp, t0, t1, seg;
            while (it.next(qp, t0, t1, seg)) {
                for (int tt = t0; tt < t1; tt++, tt_n++) {
                    const uint32_t nb = tt_n & 1;
                    mbar_wait(&nempty[nb], ((tt_n >> 1) & 1) ^ 1);
                    if (FOLD) {
                        mbar_expect_tx(&nfull[nb], F_AUG_T);
                        tma_load_2d(&map_aug, &nfull[nb], aug_t + nb * F_AUG_T, 0, tt * F_BN);
                        continue;
                    }
                    mbar_expect_tx(&nfull[nb], F_BN * 4);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(tn_smem + nb * F_BN)), "l"(p.tn + (size_t)tt * F_BN), "r"(F_BN * 4), "r"(smem_u32(&nfull[nb]))
                                 : "memory");
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;" ::: "memory");
        // ================================================================ epilogue (2 column groups x 4 lane quarters)
        const int ew = warp & 3, cg = warp >> 2;
        const int row = ew * 32 + lane;
        uint32_t tile_n = 0, tt_n = 0;
        SegIter it(p.n_qpairs, p.n_ttiles, p.units_per_cta, blockIdx.x);
        int qp, t0, t1, seg;
        while (it.next(qp, t0, t1, seg)) {
            float bs[F_SUB][F_TOPK];
            int bi[F_SUB][F_TOPK];
#pragma unroll
            for (int sub = 0; sub < F_SUB; sub++)
#pragma unroll
                for (int j = 0; j < F_TOPK; j++) { bs[sub][j] = INFINITY; bi[sub][j] = -1; }
            float slack[F_SUB];
#pragma unroll
            for (int sub = 0; sub < F_SUB; sub++) {
                const int qrow = (qp * F_SUB + sub) * F_BM + row;
                slack[sub] = slack_of(qrow < p.nq ? p.qn[qrow] : 0.f, __uint_as_float(*p.tn_max_bits), (float)F_KAPPA);
            }
            float floor_thr[F_SUB];
#pragma unroll
            for (int sub = 0; sub < F_SUB; sub++) floor_thr[sub] = INFINITY;
            // A query row cut into several segments is scanned by several CTAs at once, each starting cold.  Any threshold
            // one of them has reached (second best so far + slack) bounds the row's exact two nearest for ALL of them, so
            // the segments publish theirs (atomicMin on an ordered key) and fold the others' in every few tiles: the row
            // then warms up once, not once per segment.  Racy on purpose: every published value is a valid bound, and
            // floor_thr -- the lowest threshold ever in force -- is what the list's certificate bound is built from.
            const bool seed = SHARE && p.row_thr != nullptr;
            float published[F_SUB];
#pragma unroll
            for (int sub = 0; sub < F_SUB; sub++) published[sub] = INFINITY;
            if (SHARE) {
                // every epilogue warp has left the previous query rows before their thresholds are replaced
                asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_THREADS) : "memory");
#pragma unroll
                for (int sub = 0; sub < F_SUB; sub++) {
                    const int qrow = (qp * F_SUB + sub) * F_BM + row;
                    float g = INFINITY;
                    if (seed && qrow < p.nq) g = key_thr(*reinterpret_cast<volatile int32_t*>(p.row_thr + qrow));
                    thr_sh[sub * F_BM + row] = g;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_THREADS) : "memory");
            }
            for (int tt = t0; tt < t1; tt++, tt_n++) {
                const uint32_t nb = tt_n & 1;
                // every fourth tile: fetch the row's published bound now, use it after this tile's scan
                const bool sync_thr = seed && ((tt - t0) & 3) == 3;
                int32_t fetched[F_SUB];
                if (sync_thr) {
#pragma unroll
                    for (int sub = 0; sub < F_SUB; sub++) {
                        const int qrow = (qp * F_SUB + sub) * F_BM + row;
                        fetched[sub] = qrow < p.nq ? *reinterpret_cast<volatile int32_t*>(p.row_thr + qrow) : 0x7f7f7f7f;
                    }
                }
                if (!FOLD) mbar_wait(&nfull[nb], (tt_n >> 1) & 1);
                const float4* tn4 = reinterpret_cast<const float4*>(tn_smem + nb * F_BN + cg * F_EPI_COLS);
                const int colbase = tt * F_BN + cg * F_EPI_COLS;
#pragma unroll
                for (int sub = 0; sub < F_SUB; sub++, tile_n++) {
                    const uint32_t acc = tile_n & 1;
                    TCC_CLOCK(c0);
                    mbar_wait(&tfull[acc], (tile_n >> 1) & 1);
                    TCC_CLOCK(c1);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * F_BN + cg * F_EPI_COLS;
                    // all four chunks of this warp's share go to registers at once (the epilogue warps hold 224 registers, see
                    // setmaxnreg above), the accumulator is handed back, and only then are the columns examined: a slow insert
                    // path delays this warp's next read, not the next MMA
                    static_assert(F_EPI_COLS == 128, "four chunks of 32 columns per warp and tile");
                    uint32_t va[32], vb[32], vc[32], vd[32];
                    tc_ld32(taddr, va);
                    tc_ld32(taddr + 32, vb);
                    tc_ld32(taddr + 64, vc);
                    tc_ld32(taddr + 96, vd);
                    tc_wait_ld32(va);
                    tc_wait_ld32(vb);
                    tc_wait_ld32(vc);
                    tc_wait_ld32(vd);
                    if (FOLD) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[acc]);
                    }
                    scan_chunk_lean<SHARE, FOLD>(va, tn4, colbase, slack[sub], bs[sub], bi[sub], thr_sh + sub * F_BM + row, floor_thr[sub]);
                    scan_chunk_lean<SHARE, FOLD>(vb, tn4 + 8, colbase + 32, slack[sub], bs[sub], bi[sub], thr_sh + sub * F_BM + row, floor_thr[sub]);
                    scan_chunk_lean<SHARE, FOLD>(vc, tn4 + 16, colbase + 64, slack[sub], bs[sub], bi[sub], thr_sh + sub * F_BM + row, floor_thr[sub]);
                    scan_chunk_lean<SHARE, FOLD>(vd, tn4 + 24, colbase + 96, slack[sub], bs[sub], bi[sub], thr_sh + sub * F_BM + row, floor_thr[sub]);
                    if (!FOLD) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[acc]);
                    }
                    TCC_ADD(0, c1 - c0); TCC_ADD(1, clock64() - c1); TCC_ADD(2, 1);
                }
                if (!FOLD) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&nempty[nb]);
                }
                if (sync_thr) {
#pragma unroll
                    for (int sub = 0; sub < F_SUB; sub++) {
                        const int qrow = (qp * F_SUB + sub) * F_BM + row;
                        if (qrow >= p.nq) continue;
                        const float mine = floor_thr[sub];
                        if (mine < published[sub]) { atomicMin(p.row_thr + qrow, thr_key(mine)); published[sub] = mine; }
                        const float g = key_thr(fetched[sub]);
                        volatile float* rt = thr_sh + sub * F_BM + row;
                        if (g < *rt) *rt = g;
                    }
                }
            }
#pragma unroll
            for (int sub = 0; sub < F_SUB; sub++) {
                const int qg = (qp * F_SUB + sub) * F_BM + row;
                if (seed && qg < p.nq && floor_thr[sub] < published[sub]) atomicMin(p.row_thr + qg, thr_key(floor_thr[sub]));
                if (qg < p.nq) {
                    const size_t l = ((size_t)qg * p.n_seg + seg) * F_EPI_GROUPS + cg;
#pragma unroll
                    for (int j = 0; j < F_TOPK; j += 4) {
                        *reinterpret_cast<int4*>(p.cand_idx + l * F_TOPK + j) = make_int4(bi[sub][j], bi[sub][j + 1], bi[sub][j + 2], bi[sub][j + 3]);
                        *reinterpret_cast<float4*>(p.cand_s + l * F_TOPK + j) = make_float4(bs[sub][j], bs[sub][j + 1], bs[sub][j + 2], bs[sub][j + 3]);
                    }
                    p.cand_thr[l] = fminf(floor_thr[sub], __fadd_rn(bs[sub][1], slack[sub]));
                }
            }
        }
    }

    TCC_FLUSH();
    tc_fence_before();
    __syncthreads();
    if (warp == W_ALLOC) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

#if ERP_TC_COUNTERS == 2
__global__ void dbg_print_kernel()
{
    printf("warp-events (first lane) steady %llu, %.0f clk each; warm-up %llu, %.0f clk each\n", erp_evt[1], (double)erp_evt[0] / erp_evt[1], erp_evt[3], (double)erp_evt[2] / erp_evt[3]);
    printf("lane-events %llu groups %llu inserts %llu\n", erp_evt[4], erp_evt[5], erp_evt[6]);
    for (int i = 0; i < 8; i++) erp_evt[i] = 0;
}
#endif
#if ERP_TC_COUNTERS == 1
__global__ void dbg_print_kernel()
{
    printf("epi: wait_tfull %.0f scan %.0f per warp-tile (%llu warp-tiles) | mma: wait_tempty %.0f wait_full(sum over chunks) %.0f per tile (%llu tiles)\n",
           (double)erp_clk[0] / erp_clk[2], (double)erp_clk[1] / erp_clk[2], erp_clk[2], (double)erp_clk[3] / erp_clk[4], (double)erp_clk[5] / erp_clk[4], erp_clk[4]);
    for (int i = 0; i < 8; i++) erp_clk[i] = 0;
}
#endif
template <int KCH>
static int launch_tc1(erp_ctx* ctx, const CUtensorMap& mq, const CUtensorMap& mt, const CUtensorMap& ma, const Tc1Params& p, int grid)
{
    constexpr int smem = f_smem(KCH);
    static_assert(smem <= TC_SMEM_LIMIT, "shared memory budget");
    static std::atomic<uint64_t> configured{0};
    ERP_TRY(ensure_dynamic_smem(ctx, knn2_tc1_kernel<KCH>, smem, configured));
    knn2_tc1_kernel<KCH><<<grid, F_THREADS, smem, ctx->stream>>>(mq, mt, ma, p);
    ERP_LAUNCH(ctx, "knn2_tc1_kernel");
#if ERP_TC_COUNTERS
    dbg_print_kernel<<<1, 1, 0, ctx->stream>>>();
    cudaStreamSynchronize(ctx->stream);
#endif
    return ERP_OK;
}

int knn2_tc1(erp_ctx* ctx, const float* d_q, int nq, const float* d_t, int nt, int dim,
             int32_t* d_idx2, float* d_dist2, double* d_d2)
{
    if (!knn2_tc_supported(nq, nt, dim)) { set_error("tcgen05 engine: unsupported shape nq=%d nt=%d dim=%d", nq, nt, dim); return ERP_E_DIM; }
    const int kch = (dim + F_KC - 1) / F_KC, dpad = kch * F_KC;
    const int n_qpairs = cdiv(nq, F_BM * f_sub(kch)), n_ttiles = cdiv(nt, F_BN);
    // stream-K plan: equal contiguous unit ranges, a query pair cut in at most F_MAX_SEG segments
    long total = (long)n_qpairs * n_ttiles;
    long L = (total + ctx->sm_count - 1) / ctx->sm_count;
    long lmin = (n_ttiles + (F_MAX_SEG - 2) - 1) / (F_MAX_SEG - 2);
    if (L < lmin) L = lmin;
    if (L < 1) L = 1;
    const int grid = (int)((total + L - 1) / L);
    long segs = (n_ttiles + L - 1) / L + 1;
    const int n_seg = (int)(segs > F_MAX_SEG ? F_MAX_SEG : segs);
    const int n_lists = n_seg * F_EPI_GROUPS;

    int st = ERP_OK;
    float* qs = ctx->scratch<float>(S_TC_Q, (size_t)nq * dpad, &st);
    float* ts = ctx->scratch<float>(S_TC_T, (size_t)nt * dpad, &st);
    float* tn = ctx->scratch<float>(S_TC_TN, (size_t)n_ttiles * F_BN * 9, &st);      // norms, then the 8-float norm operand rows
    int32_t* cand = ctx->scratch<int32_t>(S_TC_CAND, (size_t)nq * n_lists * (F_TOPK * 2 + 1), &st);
    float* qn = ctx->scratch<float>(S_TC_QN, (size_t)nq + 8, &st);
    int32_t* list = ctx->scratch<int32_t>(S_TC_LIST, (size_t)nq + 8, &st);
    int32_t* misc = ctx->scratch<int32_t>(S_TC_MISC, 8, &st);
    // rows cut into segments exchange their thresholds (see the kernel): 12.5k x 100k (one rank of an 8-way split, 5 segments
    // per row) 0.308 -> 0.260 ms, 25k x 100k 0.519 -> 0.474, 100k x 100k (2 segments) 1.759 -> 1.731
    static const int seed_from = [] { const char* e = getenv("ERP_B200_SEED_SEGS"); return e ? atoi(e) : 2; }();     // experiments: 99 = never
    int32_t* row_thr = n_seg >= seed_from ? ctx->scratch<int32_t>(S_TC_ROWTHR, (size_t)nq + 8, &st) : nullptr;
    ERP_TRY(st);
    if (row_thr) ERP_CUDA(cudaMemsetAsync(row_thr, 0x7f, sizeof(int32_t) * (size_t)nq, ctx->stream));
    float* cand_s = reinterpret_cast<float*>(cand + (size_t)nq * n_lists * F_TOPK);
    float* cand_thr = cand_s + (size_t)nq * n_lists * F_TOPK;
    ERP_CUDA(cudaMemsetAsync(cand_thr, 0x7f, (size_t)nq * n_lists * sizeof(float), ctx->stream));              // ~3.4e38: "unbounded"
    ERP_TRY(tc_misc_begin(ctx, misc));
    ERP_CUDA(cudaMemsetAsync(cand, 0xFF, (size_t)nq * n_lists * F_TOPK * sizeof(int32_t), ctx->stream));

    float* aug = tn + (size_t)n_ttiles * F_BN;
    RoundJob jq = {d_q, nq, -2.0f, qs, qn, nq, reinterpret_cast<unsigned*>(misc + 3), nullptr, 0};
    RoundJob jt = {ctx->tc_chunk == 0 ? d_t : nullptr, nt, 1.0f, ts, tn, n_ttiles * F_BN, reinterpret_cast<unsigned*>(misc + 1),
                   f_fold(kch) ? aug : nullptr, 0};
    ERP_TRY(launch_prep(ctx, jq, jt, dim, dpad));

    CUtensorMap mq, mt;
    ERP_TRY(make_map(&mq, qs, nq, dpad, F_BM));
    ERP_TRY(make_map(&mt, ts, nt, dpad, F_BN));
    CUtensorMap ma = mt;
    if (f_fold(kch)) ERP_TRY(make_map(&ma, aug, n_ttiles * F_BN, 8, F_BN, true));
    Tc1Params p;
    p.nq = nq; p.nt = nt; p.n_qpairs = n_qpairs; p.n_ttiles = n_ttiles; p.units_per_cta = (int)L; p.n_seg = n_seg;
    p.tn = tn; p.cand_idx = cand; p.cand_s = cand_s; p.cand_thr = cand_thr; p.qn = qn;
    p.tn_max_bits = reinterpret_cast<const unsigned*>(misc + 1);
    p.row_thr = row_thr;

    if (ctx->tc_chunk == 0) ERP_CUDA(record_timing(ctx, ctx->ev_k0));
    switch (kch) {
    case 1: ERP_TRY(launch_tc1<1>(ctx, mq, mt, ma, p, grid)); break;
    case 2: ERP_TRY(launch_tc1<2>(ctx, mq, mt, ma, p, grid)); break;
    case 3: ERP_TRY(launch_tc1<3>(ctx, mq, mt, ma, p, grid)); break;
    default: ERP_TRY(launch_tc1<4>(ctx, mq, mt, ma, p, grid)); break;
    }
    ERP_CUDA(record_timing(ctx, ctx->ev_k1));

    ERP_TRY(refine_launch(ctx, d_q, nq, d_t, nt, dim, n_lists, F_TOPK, F_KAPPA, cand, cand_s, cand_thr, reinterpret_cast<unsigned*>(misc),
                          d_idx2, d_dist2, d_d2, list));
    ERP_TRY(knn2_exact_rescan(ctx, d_q, nq, d_t, nt, dim, list, misc, nq, d_idx2, d_dist2, d_d2));
    ERP_TRY(tc_misc_end(ctx, misc));

    ctx->knn_stats[0] = ERP_ENGINE_TCGEN05_1X;
    ctx->knn_stats[1] = -1;
    ctx->knn_stats[2] = n_seg;
    ctx->knn_stats[3] = (int)L;
    ctx->knn_stats[4] = grid;
    ctx->tc_misc_dev = misc;
    return ERP_OK;
}

} // namespace erp
