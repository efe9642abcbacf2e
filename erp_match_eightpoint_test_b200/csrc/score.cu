// score.cu -- hypothesis scoring, inlier mask, best-model selection and the RANSAC glue.
//
// The residual is the one /root/reference/src/epipolar_tool.cpp:100-107 evaluates
// (result = l . (E^T p), |result| < 0.002), written for eight_point's convention l^T E r.
// The arithmetic is fp32 with an explicit fma chain and is spelled identically in the oracle
// (oracle/erp_oracle.c: is_inlier), so inlier COUNTS are bit-exact, not approximately equal:
//     Eh   = (float)(E * sqrt(2)/|E|_F)
//     k_ab = l_a * r_b
//     res  = fma(Eh8,k8, ... fma(Eh1,k1, Eh0*k0))
#include "score_common.cuh"

namespace erp {

constexpr int TH = 128;         // hypotheses per block: their scaled E stay in shared memory
constexpr int SC_THREADS = 256;

// grid (ceil(H/TH), msplit).  Each thread keeps C correspondences (their 9 products k_ab) in
// registers and walks the block's hypothesis tile, whose 9 coefficients arrive as shared-memory
// broadcasts: 9 FMA-pipe instructions per (hypothesis, correspondence) and nothing else in the
// inner loop.  The correspondence slice is read once per hypothesis tile (32 B each, coalesced
// float4 pairs); counts are reduced in-warp (REDUX) and merged in shared memory.
template <int METRIC, int C>
__global__ void __launch_bounds__(SC_THREADS)
score_kernel(const double* __restrict__ E, int H, const float4* __restrict__ l4,
             const float4* __restrict__ r4, int m_cap, const int32_t* __restrict__ m_dev, float tau, float tau2, float sin2,
             int32_t* __restrict__ counts,
             const int32_t* __restrict__ hlist, const int32_t* __restrict__ hlist_len,
             unsigned long long* __restrict__ best, unsigned long long hyp0, int32_t* __restrict__ tile_done)
{
    // with hlist: row i of the launch is hypothesis hlist[i], i < *hlist_len (blocks past the end
    // exit: the host does not know the length); counts are indexed by list position.
    // with best: the block that completes a hypothesis tile (the last of its gridDim.y correspondence slices, found
    // with one ticket per tile) merges the tile's packed (count << 32 | ~id) maximum into *best.
    __shared__ __align__(16) float Es[TH][12];
    __shared__ int cs[TH];
    __shared__ int last_slice;
    const int m = dev_len(m_dev, m_cap);
    if (hlist) H = min(H, *hlist_len);
    // hypothesis tiles are walked with a grid stride: list launches are sized for a plausible length,
    // not for the worst case (thousands of blocks that would only read the length and exit)
    for (int h0 = blockIdx.x * TH; h0 < H; h0 += gridDim.x * TH) {
    for (int t = threadIdx.x; t < TH; t += SC_THREADS) {
        int h = h0 + t;
        float e[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (h < H) scale_E(E + (size_t)(hlist ? hlist[h] : h) * 9, e);
#pragma unroll
        for (int i = 0; i < 9; i++) Es[t][i] = e[i];
        cs[t] = 0;
    }
    __syncthreads();
    const int nh = min(TH, H - h0);
    const int per = (m + gridDim.y - 1) / gridDim.y;
    const int c0 = blockIdx.y * per, c1 = min(m, c0 + per);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int base = c0; base < c1; base += SC_THREADS * C) {
        float k[C][9];
        float4 lv[C], rv[C];
        bool live[C];
#pragma unroll
        for (int j = 0; j < C; j++) {
            int c = base + j * SC_THREADS + threadIdx.x;
            live[j] = c < c1;
            lv[j] = live[j] ? l4[c] : zero;
            rv[j] = live[j] ? r4[c] : zero;
            kron9(lv[j], rv[j], k[j]);
        }
#pragma unroll 2
        for (int h = 0; h < nh; h++) {
            float4 e0 = *reinterpret_cast<const float4*>(&Es[h][0]);
            float4 e1 = *reinterpret_cast<const float4*>(&Es[h][4]);
            float e8 = Es[h][8];
            float e[9] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w, e8};
            int c = 0;
#pragma unroll
            for (int j = 0; j < C; j++) c += (int)(inlier<METRIC>(e, k[j], lv[j], rv[j], tau, tau2, sin2) & live[j]);
            c = __reduce_add_sync(0xffffffffu, c);
            if ((threadIdx.x & 31) == 0 && c) atomicAdd(&cs[h], c);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nh; t += SC_THREADS) {
        if (gridDim.y == 1) counts[h0 + t] = cs[t];
        else if (cs[t]) atomicAdd(&counts[h0 + t], cs[t]);
    }
    if (best) {
        if (gridDim.y > 1) {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) last_slice = atomicAdd(&tile_done[h0 / TH], 1) == (int)gridDim.y - 1;
            __syncthreads();
        }
        if (gridDim.y == 1 || last_slice) {
            __threadfence();
            unsigned long long b = 0;
            for (int t = threadIdx.x; t < nh; t += SC_THREADS) {
                const int c = gridDim.y == 1 ? cs[t] : *reinterpret_cast<volatile int32_t*>(&counts[h0 + t]);
                const unsigned long long id = hyp0 + (unsigned long long)(hlist ? hlist[h0 + t] : h0 + t);
                const unsigned long long pk = ((unsigned long long)(uint32_t)c << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)id);
                b = pk > b ? pk : b;
            }
            for (int o = 16; o > 0; o >>= 1) { unsigned long long y = __shfl_down_sync(0xffffffffu, b, o); b = y > b ? y : b; }
            if ((threadIdx.x & 31) == 0 && b) atomicMax(best, b);
        }
    }
    __syncthreads();
    }
}

// Exact counts of a LIST of hypotheses (the contenders of the tensor-core search: a few hundred of a million) and their
// packed best, algebraic residual.  Roles are transposed with respect to score_kernel: a thread owns TWO hypotheses (their
// coefficients interleaved in packed fp32x2 registers), a block stages a slice of correspondences as Kronecker products
// in shared memory and every thread walks it with broadcast reads -- no cross-thread reduction per hypothesis, full
// instruction-level parallelism across correspondences, and many small blocks (tile x slice) so that a short list still
// fills the machine.  mul.rn.f32x2 / fma.rn.f32x2 round each half like the scalar instructions: the fma chain is the
// one used everywhere else and counts are bit-exact.
constexpr int SL_THREADS = 64;              // threads per block
constexpr int SL_HYPS = 2 * SL_THREADS;     // hypotheses per block
constexpr int SL_CORR = 256;                // correspondences per block
__device__ __forceinline__ unsigned long long pack2(float a, float b)
{
    unsigned long long p;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a), "f"(b));
    return p;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__global__ void __launch_bounds__(SL_THREADS)
score_list_kernel(const double* __restrict__ E, int H_cap, const int32_t* __restrict__ hlist, const int32_t* __restrict__ hlist_len,
                  const float4* __restrict__ l4, const float4* __restrict__ r4, int m_cap, const int32_t* __restrict__ m_dev, float tau,
                  int32_t* __restrict__ counts, int32_t* __restrict__ tile_done, unsigned long long* __restrict__ best,
                  unsigned long long hyp0)
{
    __shared__ __align__(16) float ks[SL_CORR][12];
    __shared__ int last_slice;
    const int H = min(H_cap, *hlist_len);
    if ((int)blockIdx.x * SL_HYPS >= H) return;
    const int m = dev_len(m_dev, m_cap);
    const int c0 = blockIdx.y * SL_CORR, n = max(0, min(SL_CORR, m - c0));
    for (int i = threadIdx.x; i < n; i += SL_THREADS) {
        float k[9];
        kron9(l4[c0 + i], r4[c0 + i], k);
        *reinterpret_cast<float4*>(&ks[i][0]) = make_float4(k[0], k[1], k[2], k[3]);
        *reinterpret_cast<float4*>(&ks[i][4]) = make_float4(k[4], k[5], k[6], k[7]);
        ks[i][8] = k[8];
    }
    __syncthreads();
    for (int h0 = blockIdx.x * SL_HYPS; h0 < H; h0 += gridDim.x * SL_HYPS) {
        const int ha = h0 + 2 * threadIdx.x, hb = ha + 1;
        float ea[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, eb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (ha < H) scale_E(E + (size_t)hlist[ha] * 9, ea);
        if (hb < H) scale_E(E + (size_t)hlist[hb] * 9, eb);
        unsigned long long e2[9];
#pragma unroll
        for (int i = 0; i < 9; i++) e2[i] = pack2(ea[i], eb[i]);
        int cnt_a = 0, cnt_b = 0;
#pragma unroll 4
        for (int i = 0; i < n; i++) {
            const float4 k0 = *reinterpret_cast<const float4*>(&ks[i][0]), k1 = *reinterpret_cast<const float4*>(&ks[i][4]);
            const float k8 = ks[i][8];
            unsigned long long res = mul2(e2[0], pack2(k0.x, k0.x));
            res = fma2(e2[1], pack2(k0.y, k0.y), res); res = fma2(e2[2], pack2(k0.z, k0.z), res); res = fma2(e2[3], pack2(k0.w, k0.w), res);
            res = fma2(e2[4], pack2(k1.x, k1.x), res); res = fma2(e2[5], pack2(k1.y, k1.y), res); res = fma2(e2[6], pack2(k1.z, k1.z), res);
            res = fma2(e2[7], pack2(k1.w, k1.w), res); res = fma2(e2[8], pack2(k8, k8), res);
            float ra, rb;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(ra), "=f"(rb) : "l"(res));
            cnt_a += fabsf(ra) < tau ? 1 : 0;
            cnt_b += fabsf(rb) < tau ? 1 : 0;
        }
        if (ha < H && cnt_a) atomicAdd(&counts[ha], cnt_a);
        if (hb < H && cnt_b) atomicAdd(&counts[hb], cnt_b);
        // the block that completes a hypothesis tile (last of its gridDim.y slices) merges the tile's packed best
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last_slice = atomicAdd(&tile_done[h0 / SL_HYPS], 1) == (int)gridDim.y - 1;
        __syncthreads();
        if (last_slice) {
            __threadfence();
            unsigned long long b = 0;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int h = ha + u;
                if (h < H) {
                    const int c = *reinterpret_cast<volatile int32_t*>(&counts[h]);
                    const unsigned long long id = hyp0 + (unsigned long long)hlist[h];
                    const unsigned long long pk = ((unsigned long long)(uint32_t)c << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)id);
                    b = pk > b ? pk : b;
                }
            }
            for (int o = 16; o > 0; o >>= 1) { unsigned long long y = __shfl_down_sync(0xffffffffu, b, o); b = y > b ? y : b; }
            if ((threadIdx.x & 31) == 0 && b) atomicMax(best, b);
        }
        __syncthreads();
    }
}

// packed best over a count array: (count << 32) | (0xFFFFFFFF - global id)
__global__ void best_kernel(const int32_t* __restrict__ counts, int H, uint64_t hyp0,
                            unsigned long long* __restrict__ best,
                            const int32_t* __restrict__ hlist, const int32_t* __restrict__ hlist_len)
{
    unsigned long long b = 0;
    if (hlist) H = min(H, *hlist_len);
    for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < H; h += gridDim.x * blockDim.x) {
        unsigned long long id = hyp0 + (unsigned long long)(hlist ? hlist[h] : h);
        unsigned long long p = ((unsigned long long)(uint32_t)counts[h] << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)id);
        b = p > b ? p : b;
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long y = __shfl_down_sync(0xffffffffu, b, o);
        b = y > b ? y : b;
    }
    if ((threadIdx.x & 31) == 0 && b) atomicMax(best, b);
}

template <int METRIC>
__global__ void mask_kernel(const double* __restrict__ E9, const float4* __restrict__ l4,
                            const float4* __restrict__ r4, int m, float tau, float tau2, float sin2,
                            uint8_t* __restrict__ mask, int32_t* __restrict__ n_in)
{
    __shared__ float Es[9];
    if (threadIdx.x == 0) scale_E(E9, Es);
    __syncthreads();
    float e[9];
#pragma unroll
    for (int i = 0; i < 9; i++) e[i] = Es[i];
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    bool in = false;
    if (c < m) {
        float4 l = l4[c], r = r4[c];
        float k[9];
        kron9(l, r, k);
        in = inlier<METRIC>(e, k, l, r, tau, tau2, sin2);
        if (mask) mask[c] = in ? 1 : 0;
    }
    int s = __reduce_add_sync(0xffffffffu, in ? 1 : 0);
    if ((threadIdx.x & 31) == 0 && s && n_in) atomicAdd(n_in, s);
}

int gram_batch(erp_ctx* ctx, const double* d_l3, const double* d_r3, int m, const int32_t* d_m, const int32_t* d_samples,
               int H, int S, uint64_t seed, uint64_t hyp0, const uint64_t* d_packed, double* d_G);
int solve_batch(erp_ctx* ctx, const double* d_G, int H, double* d_E, float* d_pose, int max_sweeps = 30);
int solve_min8(erp_ctx* ctx, const double* d_l3, const double* d_r3, int m, const int32_t* d_m, const int32_t* d_samples, int H,
               uint64_t seed, uint64_t hyp0, const uint64_t* d_packed, double* d_E, float* d_pose, const Min8Fused* fused);
int finish_launch(erp_ctx* ctx, const double* d_Eb, const uint64_t* d_packed, const double* d_l3, const double* d_r3,
                  const float* d_l4, const float* d_r4, int m_cap, const int32_t* d_m, int metric, float tau, float tau2, float sin2,
                  uint8_t* d_mask, erp_ransac_result* d_result);

static inline void thresholds(float tau, float& tau2, float& sin2)
{
    tau2 = tau * tau;
    double sd = sin((double)tau);
    sin2 = (float)(sd * sd);
}

int score_launch(erp_ctx* ctx, const double* d_E, int H, const float* d_l4, const float* d_r4, int m_cap, const int32_t* d_m,
                 int metric, float tau, int32_t* d_counts)
{
    float tau2, sin2;
    thresholds(tau, tau2, sin2);
    int gx = cdiv(H, TH);
    // split the correspondences when there are too few hypothesis tiles to fill the machine
    int msplit = 1;
    if (gx < ctx->sm_count * 4) msplit = max(1, min(cdiv(m_cap, SC_THREADS * 8), (ctx->sm_count * 4) / gx));
    // (a zero-padded lane contributes res = 0 < tau, hence the live[] predicate in the kernel)
    if (msplit > 1) ERP_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * (size_t)H, ctx->stream));
    dim3 grid(gx, msplit);
    const float4* l4 = (const float4*)d_l4;
    const float4* r4 = (const float4*)d_r4;
    switch (metric) {
    case ERP_METRIC_ALGEBRAIC: score_kernel<ERP_METRIC_ALGEBRAIC, 8><<<grid, SC_THREADS, 0, ctx->stream>>>(d_E, H, l4, r4, m_cap, d_m, tau, tau2, sin2, d_counts, nullptr, nullptr, nullptr, 0, nullptr); break;
    case ERP_METRIC_SAMPSON: score_kernel<ERP_METRIC_SAMPSON, 4><<<grid, SC_THREADS, 0, ctx->stream>>>(d_E, H, l4, r4, m_cap, d_m, tau, tau2, sin2, d_counts, nullptr, nullptr, nullptr, 0, nullptr); break;
    case ERP_METRIC_ANGULAR: score_kernel<ERP_METRIC_ANGULAR, 4><<<grid, SC_THREADS, 0, ctx->stream>>>(d_E, H, l4, r4, m_cap, d_m, tau, tau2, sin2, d_counts, nullptr, nullptr, nullptr, 0, nullptr); break;
    default: set_error("unknown metric %d", metric); return ERP_E_ARG;
    }
    ERP_LAUNCH(ctx, "score_kernel");
    return ERP_OK;
}

// exact counts of the listed hypotheses, their packed best merged into *d_best by the same launch
int score_list_best(erp_ctx* ctx, const double* d_E, int H_max, const int32_t* d_list, const int32_t* d_len,
                    const float* d_l4, const float* d_r4, int m_cap, const int32_t* d_m, int metric, float tau, uint64_t hyp0,
                    int32_t* d_counts /* H_max + H_max / 128 + 1, scratch */, uint64_t* d_best)
{
    static_assert(SL_HYPS == TH, "one ticket per 128 hypotheses in both list kernels");
    const int tiles = cdiv(H_max, SL_HYPS);
    ERP_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * ((size_t)H_max + tiles), ctx->stream));
    if (metric != ERP_METRIC_ALGEBRAIC) {
        // Sampson / angular: the bound is looser, the list is long (thousands): score_kernel in list mode, hypothesis
        // tiles x correspondence slices, the packed best merged by the last slice of a tile
        float tau2, sin2;
        thresholds(tau, tau2, sin2);
        int msplit = max(1, min(cdiv(m_cap, SC_THREADS * 4), 24));
        dim3 grid(min(tiles, 1024), msplit);
        if (metric == ERP_METRIC_SAMPSON)
            score_kernel<ERP_METRIC_SAMPSON, 4><<<grid, SC_THREADS, 0, ctx->stream>>>(d_E, H_max, (const float4*)d_l4, (const float4*)d_r4, m_cap, d_m, tau, tau2, sin2,
                                                                                    d_counts, d_list, d_len, (unsigned long long*)d_best, (unsigned long long)hyp0, d_counts + H_max);
        else
            score_kernel<ERP_METRIC_ANGULAR, 4><<<grid, SC_THREADS, 0, ctx->stream>>>(d_E, H_max, (const float4*)d_l4, (const float4*)d_r4, m_cap, d_m, tau, tau2, sin2,
                                                                                    d_counts, d_list, d_len, (unsigned long long*)d_best, (unsigned long long)hyp0, d_counts + H_max);
        ERP_LAUNCH(ctx, "score_kernel(list)");
        return ERP_OK;
    }
    // the list is usually short (cfg3: ~900 contenders = 7 hypothesis tiles): tile x slice blocks of 128 threads
    // (score_kernel in list mode took 100 us for them, this kernel 20)
    dim3 grid(min(tiles, 32), max(1, cdiv(m_cap, SL_CORR)));
    score_list_kernel<<<grid, SL_THREADS, 0, ctx->stream>>>(d_E, H_max, d_list, d_len, (const float4*)d_l4, (const float4*)d_r4, m_cap, d_m, tau,
                                                        d_counts, d_counts + H_max, (unsigned long long*)d_best, (unsigned long long)hyp0);
    ERP_LAUNCH(ctx, "score_list_kernel");
    return ERP_OK;
}

int mask_launch(erp_ctx* ctx, const double* d_E9, const float* d_l4, const float* d_r4, int m, int metric,
                float tau, uint8_t* d_mask, int32_t* d_n)
{
    float tau2, sin2;
    thresholds(tau, tau2, sin2);
    if (d_n) ERP_CUDA(cudaMemsetAsync(d_n, 0, sizeof(int32_t), ctx->stream));
    if (m == 0) return ERP_OK;
    const float4* l4 = (const float4*)d_l4;
    const float4* r4 = (const float4*)d_r4;
    int grid = cdiv(m, 256);
    switch (metric) {
    case ERP_METRIC_ALGEBRAIC: mask_kernel<ERP_METRIC_ALGEBRAIC><<<grid, 256, 0, ctx->stream>>>(d_E9, l4, r4, m, tau, tau2, sin2, d_mask, d_n); break;
    case ERP_METRIC_SAMPSON: mask_kernel<ERP_METRIC_SAMPSON><<<grid, 256, 0, ctx->stream>>>(d_E9, l4, r4, m, tau, tau2, sin2, d_mask, d_n); break;
    case ERP_METRIC_ANGULAR: mask_kernel<ERP_METRIC_ANGULAR><<<grid, 256, 0, ctx->stream>>>(d_E9, l4, r4, m, tau, tau2, sin2, d_mask, d_n); break;
    default: set_error("unknown metric %d", metric); return ERP_E_ARG;
    }
    ERP_LAUNCH(ctx, "mask_kernel");
    return ERP_OK;
}

bool ransac_uses_tc(erp_ctx* ctx, int H, int m_cap, int metric)
{
    const int n = H < RANSAC_CHUNK ? H : RANSAC_CHUNK;
    (void)metric;       // all three residuals: Sampson / angular through a per-hypothesis bound (score_common.cuh: row_gain)
    return ctx->engine == ERP_ENGINE_TCGEN05 || ctx->engine == ERP_ENGINE_TCGEN05_1X ||
           (ctx->engine == ERP_ENGINE_AUTO && score_tc_preferred(n, m_cap));
}

// Hypotheses [hyp_offset, hyp_offset + H): sample, solve, score, packed best into *d_packed (max-merged: the caller
// zeroes it).  m lives on the device when d_m is given (m_cap sizes the launches); nothing here waits for the host.
// k_ready: the correspondence operand of the tensor-core search is already in place (fused gather).
int ransac_search(erp_ctx* ctx, const double* d_l3, const double* d_r3, const float* d_l4, const float* d_r4, int m_cap,
                  const int32_t* d_m, uint64_t seed, uint64_t hyp_offset, int H, int S, int metric, float tau, bool k_ready,
                  uint64_t* d_packed)
{
    const int CH = RANSAC_CHUNK;   // hypotheses per pass: 75 MB of E, 2 x 134 MB of split rows (S != 8: 377 MB of Gram)
    int st = ERP_OK;
    int chunk = H < CH ? H : CH;
    double* G = S == 8 ? nullptr : ctx->scratch<double>(S_GRAM, (size_t)chunk * 45, &st);
    double* E = ctx->scratch<double>(S_E, (size_t)chunk * 9, &st);
    int32_t* counts = ctx->scratch<int32_t>(S_COUNTS, (size_t)chunk + chunk / 128 + 2, &st);
    ERP_TRY(st);
    ctx->n_ev_score = 0;
    const bool tc = ransac_uses_tc(ctx, H, m_cap, metric);
    ScoreTcBuffers b;
    Min8Fused fused;
    if (tc) {
        ERP_TRY(score_tc_buffers(ctx, chunk, m_cap, &b));
        if (!k_ready) ERP_TRY(score_tc_prepare(ctx, b, d_l4, d_r4, m_cap, d_m));
        fused.Es = b.Es; fused.big = score_tc_big(tau); fused.upper = b.upper; fused.w = b.w;
        fused.metric = metric; fused.tau = tau; fused.sin_tau = (float)sin((double)tau);
    }
    for (int h0 = 0; h0 < H; h0 += CH) {
        int n = H - h0 < CH ? H - h0 : CH;
        // (the per-chunk words must be clear before the solver publishes pass A's shape)
        if (tc && h0 > 0) ERP_CUDA(cudaMemsetAsync(b.w + W_CHUNK0, 0, (W_WORDS - W_CHUNK0) * sizeof(int32_t), ctx->stream));
        if (S == 8) ERP_TRY(solve_min8(ctx, d_l3, d_r3, m_cap, d_m, nullptr, n, seed, hyp_offset + h0, nullptr, E, nullptr, tc ? &fused : nullptr));
        else {
            ERP_TRY(gram_batch(ctx, d_l3, d_r3, m_cap, d_m, nullptr, n, S, seed, hyp_offset + h0, nullptr, G));
            ERP_TRY(solve_batch(ctx, G, n, E, nullptr));
        }
        if (tc) ERP_TRY(score_tc_search(ctx, b, E, n, d_l4, d_r4, m_cap, metric, tau, hyp_offset + h0, S == 8, counts, d_packed));
        else {
            cudaEvent_t e0, e1;
            ERP_TRY(score_event(ctx, &e0));
            ERP_TRY(score_launch(ctx, E, n, d_l4, d_r4, m_cap, d_m, metric, tau, counts));
            best_kernel<<<min(cdiv(n, 256), ctx->sm_count * 4), 256, 0, ctx->stream>>>(counts, n, hyp_offset + h0,
                                                                                    (unsigned long long*)d_packed, nullptr, nullptr);
            ERP_LAUNCH(ctx, "best_kernel");
            ERP_TRY(score_event(ctx, &e1));
        }
    }
    return ERP_OK;
}

// Replays the winning sample (*d_packed names it) with the arithmetic that scored it, marks its inliers, refits on them
// (eight_point.cpp:16-50 as a least-squares solve) and leaves the whole erp_ransac_result on the device.
int ransac_finish(erp_ctx* ctx, const double* d_l3, const double* d_r3, const float* d_l4, const float* d_r4, int m_cap,
                  const int32_t* d_m, uint64_t seed, const uint64_t* d_packed, int S, int metric, float tau,
                  uint8_t* d_mask, erp_ransac_result* d_result)
{
    ERP_ARG(metric >= 0 && metric <= 2, ERP_E_ARG, "unknown metric %d", metric);
    int st = ERP_OK;
    double* misc = ctx->scratch<double>(S_MISC, 128, &st);
    ERP_TRY(st);
    double *G = misc, *Eb = misc + 45;
    if (S == 8) ERP_TRY(solve_min8(ctx, d_l3, d_r3, m_cap, d_m, nullptr, 1, seed, 0, d_packed, Eb, nullptr, nullptr));
    else {
        ERP_TRY(gram_batch(ctx, d_l3, d_r3, m_cap, d_m, nullptr, 1, S, seed, 0, d_packed, G));
        ERP_TRY(solve_batch(ctx, G, 1, Eb, nullptr));
    }
    float tau2, sin2;
    thresholds(tau, tau2, sin2);
    return finish_launch(ctx, Eb, d_packed, d_l3, d_r3, d_l4, d_r4, m_cap, d_m, metric, tau, tau2, sin2, d_mask, d_result);
}

} // namespace erp

using namespace erp;

ERP_API int erp_score_dev(erp_ctx* ctx, const double* d_E, int H, const float* d_l4, const float* d_r4,
                          int m, int metric, float tau, uint64_t hyp_offset,
                          int32_t* d_counts, uint64_t* d_best_packed)
{
    ERP_ARG(ctx && H >= 0 && m >= 0, ERP_E_ARG, "erp_score_dev: bad argument");
    if (H == 0) return ERP_OK;
    ERP_ARG(d_E && (m == 0 || (d_l4 && d_r4)), ERP_E_ARG, "erp_score_dev: null buffer");
    DeviceGuard g(ctx->device);
    int st = ERP_OK;
    if (!d_counts) { d_counts = ctx->scratch<int32_t>(S_COUNTS, H, &st); ERP_TRY(st); }
    if (m == 0) ERP_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * (size_t)H, ctx->stream));
    else ERP_TRY(score_launch(ctx, d_E, H, d_l4, d_r4, m, nullptr, metric, tau, d_counts));
    if (d_best_packed) {
        best_kernel<<<min(cdiv(H, 256), ctx->sm_count * 4), 256, 0, ctx->stream>>>(d_counts, H, hyp_offset,
                                                                                (unsigned long long*)d_best_packed, nullptr, nullptr);
        ERP_LAUNCH(ctx, "best_kernel");
    }
    return ERP_OK;
}

ERP_API int erp_ransac_local_dev(erp_ctx* ctx, const double* d_l3, const double* d_r3,
                                 const float* d_l4, const float* d_r4, int m, uint64_t seed,
                                 uint64_t hyp_offset, int H, int S, int metric, float tau,
                                 uint64_t* d_packed)
{
    ERP_ARG(ctx && d_l3 && d_r3 && d_l4 && d_r4 && d_packed && H >= 0, ERP_E_ARG, "erp_ransac_local_dev: bad argument");
    ERP_ARG(S >= 8 && S <= 32, ERP_E_ARG, "erp_ransac_local_dev: sample size must be in [8,32], got %d", S);
    ERP_ARG(m >= S, ERP_E_TOO_FEW_POINTS, "erp_ransac_local_dev: %d correspondences < sample size %d", m, S);
    ERP_ARG(hyp_offset + (uint64_t)H <= 0xFFFFFFFFull, ERP_E_LIMIT, "hypothesis ids must fit 32 bits");
    DeviceGuard g(ctx->device);
    ERP_CUDA(cudaMemsetAsync(d_packed, 0, sizeof(uint64_t), ctx->stream));
    return ransac_search(ctx, d_l3, d_r3, d_l4, d_r4, m, nullptr, seed, hyp_offset, H, S, metric, tau, false, d_packed);
}

ERP_API int erp_ransac_finish_dev(erp_ctx* ctx, const double* d_l3, const double* d_r3,
                                  const float* d_l4, const float* d_r4, int m, uint64_t seed,
                                  uint64_t packed, int S, int metric, float tau,
                                  uint8_t* d_mask, erp_ransac_result* result)
{
    ERP_ARG(ctx && d_l3 && d_r3 && d_l4 && d_r4 && result, ERP_E_ARG, "erp_ransac_finish_dev: bad argument");
    ERP_ARG(S >= 8 && S <= 32 && m >= S, ERP_E_TOO_FEW_POINTS, "erp_ransac_finish_dev: bad sample size / too few points");
    DeviceGuard g(ctx->device);
    int st = ERP_OK;
    uint64_t* d_pk = ctx->scratch<uint64_t>(S_PACKED, 4, &st) + 2;              // [2]: the caller's word, away from the search's
    erp_ransac_result* d_res = ctx->scratch<erp_ransac_result>(S_RESULT, 1, &st);
    if (!d_mask) d_mask = ctx->scratch<uint8_t>(S_MASK, (size_t)m, &st);
    ERP_TRY(st);
    ERP_CUDA(cudaMemcpyAsync(d_pk, &packed, sizeof packed, cudaMemcpyHostToDevice, ctx->stream));
    ERP_TRY(ransac_finish(ctx, d_l3, d_r3, d_l4, d_r4, m, nullptr, seed, d_pk, S, metric, tau, d_mask, d_res));
    ERP_CUDA(cudaMemcpyAsync(result, d_res, sizeof *result, cudaMemcpyDeviceToHost, ctx->stream));
    ERP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ERP_OK;
}
