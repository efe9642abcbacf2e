// knn_exact.cu -- exact brute-force 2-NN (fp64 direct form) and the ratio / cross-check
// filter with ordered compaction.
//
// Replaces cv::DescriptorMatcher::knnMatch(k=2) + the ratio loop of
// /root/reference/src/feature_matcher.cpp:42-59 in their exact-L2 limit.  The distance
// chain is the one the oracle spells (oracle/erp_oracle.c: block_d2):
//     d2 = fma(diff_k, diff_k, d2),  diff_k = (double)q_k - (double)t_k,  k ascending
// so indices AND distances are bit-identical to the CPU oracle.  This kernel is the
// ERP_ENGINE_EXACT_SIMT engine, the re-scan path of the tcgen05 engine for queries whose
// GEMM-form candidates are ambiguous, and the on-device checker for it.
#include "common.cuh"

namespace erp {

constexpr int QT = 64;      // queries per block
constexpr int TT = 64;      // train rows per smem tile
constexpr int KC = 32;      // k-chunk staged in smem
constexpr int PITCH = 66;   // doubles per smem row (even: 16-byte aligned double2 loads)

__device__ __forceinline__ bool better(double d, int i, double bd, int bi)
{
    return d < bd || (d == bd && i < bi);
}

__device__ __forceinline__ void top2_insert(double d, int i, double& b0, int& i0, double& b1, int& i1)
{
    if (better(d, i, b1, i1)) {
        if (better(d, i, b0, i0)) { b1 = b0; i1 = i0; b0 = d; i0 = i; }
        else { b1 = d; i1 = i; }
    }
}

// grid.x = ceil(n_rows / QT) row groups, grid.y = T slices.
// Rows: r in [row_base, row_base + n_rows), further limited by *n_rows_dev when given (a list whose
// length only the device knows: blocks past the end exit).  Row r is query qlist[r] (or r).
// With one T slice the block writes final results; with several it writes its partial top-2 to
// partial[r * gridDim.y + slice] for knn2_merge_kernel.
struct Top2 { double d0, d1; int i0, i1; };

__global__ void __launch_bounds__(256)
knn2_exact_kernel(const float* __restrict__ q, int nq, const float* __restrict__ t, int nt, int dim,
                  const int32_t* __restrict__ qlist, int row_base, int n_rows, const int32_t* __restrict__ n_rows_dev,
                  int idx_offset, int32_t* __restrict__ idx2, float* __restrict__ dist2, double* __restrict__ d2out,
                  Top2* __restrict__ partial)
{
    __shared__ __align__(16) double Qs[KC][PITCH];
    __shared__ __align__(16) double Ts[KC][PITCH];
    __shared__ int qrow[QT];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    int row_end = row_base + n_rows;
    if (n_rows_dev) row_end = min(row_end, *n_rows_dev);
    const int q0 = row_base + blockIdx.x * QT;
    if (q0 >= row_end) return;
    if (tid < QT) {
        int r = q0 + tid;
        qrow[tid] = r < row_end ? (qlist ? qlist[r] : r) : -1;
    }
    __syncthreads();
    // T slice of this block (multiples of TT so that tiles never straddle slices)
    const int per = ((nt + (int)gridDim.y - 1) / (int)gridDim.y + TT - 1) / TT * TT;
    const int t_begin = min(nt, (int)blockIdx.y * per), t_end = min(nt, t_begin + per);

    double b0[4], b1[4];
    int i0[4], i1[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { b0[i] = b1[i] = INFINITY; i0[i] = i1[i] = 0x7fffffff; }

    for (int t0 = t_begin; t0 < t_end; t0 += TT) {
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = 0.0;

        for (int kc = 0; kc < dim; kc += KC) {
            __syncthreads();
            // stage 64 rows x 32 k of each operand as doubles, transposed to [k][row]
            for (int v = tid; v < QT * (KC / 4); v += 256) {
                int row = v >> 3, c4 = v & 7, k = kc + c4 * 4;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                int qr = qrow[row];
                if (qr >= 0 && k < dim) a = *reinterpret_cast<const float4*>(q + (size_t)qr * dim + k);
                int tr = t0 + row;
                if (tr < t_end && k < dim) b = *reinterpret_cast<const float4*>(t + (size_t)tr * dim + k);
                Qs[c4 * 4 + 0][row] = (double)a.x; Qs[c4 * 4 + 1][row] = (double)a.y;
                Qs[c4 * 4 + 2][row] = (double)a.z; Qs[c4 * 4 + 3][row] = (double)a.w;
                Ts[c4 * 4 + 0][row] = (double)b.x; Ts[c4 * 4 + 1][row] = (double)b.y;
                Ts[c4 * 4 + 2][row] = (double)b.z; Ts[c4 * 4 + 3][row] = (double)b.w;
            }
            __syncthreads();
            const int klim = min(KC, dim - kc);
#pragma unroll 4
            for (int k = 0; k < klim; k++) {
                double2 qa = *reinterpret_cast<const double2*>(&Qs[k][ty * 4]);
                double2 qb = *reinterpret_cast<const double2*>(&Qs[k][ty * 4 + 2]);
                double2 ta = *reinterpret_cast<const double2*>(&Ts[k][tx * 4]);
                double2 tb = *reinterpret_cast<const double2*>(&Ts[k][tx * 4 + 2]);
                double qv[4] = {qa.x, qa.y, qb.x, qb.y};
                double tv[4] = {ta.x, ta.y, tb.x, tb.y};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        double diff = __dsub_rn(qv[i], tv[j]);
                        acc[i][j] = __fma_rn(diff, diff, acc[i][j]);
                    }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int tr = t0 + tx * 4 + j;
            if (tr < t_end) {
#pragma unroll
                for (int i = 0; i < 4; i++) top2_insert(acc[i][j], tr, b0[i], i0[i], b1[i], i1[i]);
            }
        }
    }

    // merge the 16 column-threads of each query row
    __syncthreads();
    double* md = &Qs[0][0];                       // 64 x 16 x 2 doubles = 2048 <= 32*66
    int* mi = reinterpret_cast<int*>(&Ts[0][0]);  // 64 x 16 x 2 ints
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int row = ty * 4 + i;
        md[(row * 16 + tx) * 2 + 0] = b0[i]; md[(row * 16 + tx) * 2 + 1] = b1[i];
        mi[(row * 16 + tx) * 2 + 0] = i0[i]; mi[(row * 16 + tx) * 2 + 1] = i1[i];
    }
    __syncthreads();
    if (tid < QT && qrow[tid] >= 0) {
        double B0 = INFINITY, B1 = INFINITY;
        int I0 = 0x7fffffff, I1 = 0x7fffffff;
        for (int c = 0; c < 32; c++) top2_insert(md[tid * 32 + c], mi[tid * 32 + c], B0, I0, B1, I1);
        if (gridDim.y > 1) {
            Top2 r; r.d0 = B0; r.d1 = B1; r.i0 = I0; r.i1 = I1;
            partial[(size_t)(q0 + tid) * gridDim.y + blockIdx.y] = r;
            return;
        }
        int o = qrow[tid];
        if (idx2) {
            idx2[2 * o] = I0 == 0x7fffffff ? -1 : I0 + idx_offset;
            idx2[2 * o + 1] = I1 == 0x7fffffff ? -1 : I1 + idx_offset;
        }
        if (dist2) { dist2[2 * o] = (float)sqrt(B0); dist2[2 * o + 1] = (float)sqrt(B1); }
        if (d2out) { d2out[2 * o] = B0; d2out[2 * o + 1] = B1; }
    }
}

// merges the T-slice partials of listed row r (one thread per row)
__global__ void knn2_merge_kernel(const Top2* __restrict__ partial, int slices, const int32_t* __restrict__ qlist,
                                  int n_rows, const int32_t* __restrict__ n_rows_dev, int idx_offset,
                                  int32_t* __restrict__ idx2, float* __restrict__ dist2, double* __restrict__ d2out)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    int row_end = n_rows_dev ? min(n_rows, *n_rows_dev) : n_rows;
    if (r >= row_end) return;
    double B0 = INFINITY, B1 = INFINITY;
    int I0 = 0x7fffffff, I1 = 0x7fffffff;
    for (int s = 0; s < slices; s++) {
        Top2 p = partial[(size_t)r * slices + s];
        top2_insert(p.d0, p.i0, B0, I0, B1, I1);
        top2_insert(p.d1, p.i1, B0, I0, B1, I1);
    }
    int o = qlist ? qlist[r] : r;
    if (idx2) {
        idx2[2 * o] = I0 == 0x7fffffff ? -1 : I0 + idx_offset;
        idx2[2 * o + 1] = I1 == 0x7fffffff ? -1 : I1 + idx_offset;
    }
    if (dist2) { dist2[2 * o] = (float)sqrt(B0); dist2[2 * o + 1] = (float)sqrt(B1); }
    if (d2out) { d2out[2 * o] = B0; d2out[2 * o + 1] = B1; }
}

int knn2_exact(erp_ctx* ctx, const float* d_q, int nq, const float* d_t, int nt, int dim,
               const int32_t* d_qlist, int nlist, int idx_offset,
               int32_t* d_idx2, float* d_dist2, double* d_d2)
{
    int rows = d_qlist ? nlist : nq;
    if (rows <= 0) return ERP_OK;
    knn2_exact_kernel<<<cdiv(rows, QT), 256, 0, ctx->stream>>>(d_q, nq, d_t, nt, dim, d_qlist, 0, rows, nullptr,
                                                              idx_offset, d_idx2, d_dist2, d_d2, nullptr);
    ERP_LAUNCH(ctx, "knn2_exact_kernel");
    return ERP_OK;
}

// Re-scan of a device-resident query list whose length (*d_count <= max_rows) the host does not
// know: the first RESCAN_SMALL entries run T-split (short lists are the normal case and must not
// serialise on one SM), the rest one block per 64 queries.  Blocks past the count exit at once,
// so the launch sequence is fixed and nothing synchronises with the host.
constexpr int RS_SMALL = 1024, RS_SLICES = 128;

int knn2_exact_rescan(erp_ctx* ctx, const float* d_q, int nq, const float* d_t, int nt, int dim,
                      const int32_t* d_list, const int32_t* d_count, int max_rows,
                      int32_t* d_idx2, float* d_dist2, double* d_d2)
{
    if (max_rows <= 0) return ERP_OK;
    int st = ERP_OK;
    int small = max_rows < RS_SMALL ? max_rows : RS_SMALL;
    int slices = cdiv(nt, TT) < RS_SLICES ? cdiv(nt, TT) : RS_SLICES;
    if (slices > 1) {
        Top2* partial = ctx->scratch<Top2>(S_RS_PARTIAL, (size_t)cdiv(small, QT) * QT * slices, &st);
        ERP_TRY(st);
        dim3 grid(cdiv(small, QT), slices);
        knn2_exact_kernel<<<grid, 256, 0, ctx->stream>>>(d_q, nq, d_t, nt, dim, d_list, 0, small, d_count, 0,
                                                         nullptr, nullptr, nullptr, partial);
        ERP_LAUNCH(ctx, "knn2_exact_kernel(split)");
        knn2_merge_kernel<<<cdiv(small, 256), 256, 0, ctx->stream>>>(partial, slices, d_list, small, d_count, 0,
                                                                     d_idx2, d_dist2, d_d2);
        ERP_LAUNCH(ctx, "knn2_merge_kernel");
    } else {
        knn2_exact_kernel<<<cdiv(small, QT), 256, 0, ctx->stream>>>(d_q, nq, d_t, nt, dim, d_list, 0, small, d_count, 0,
                                                                   d_idx2, d_dist2, d_d2, nullptr);
        ERP_LAUNCH(ctx, "knn2_exact_kernel(list)");
    }
    if (max_rows > RS_SMALL) {
        knn2_exact_kernel<<<cdiv(max_rows - RS_SMALL, QT), 256, 0, ctx->stream>>>(d_q, nq, d_t, nt, dim, d_list, RS_SMALL,
                                                                                 max_rows - RS_SMALL, d_count, 0,
                                                                                 d_idx2, d_dist2, d_d2, nullptr);
        ERP_LAUNCH(ctx, "knn2_exact_kernel(overflow)");
    }
    return ERP_OK;
}

// ---------------------------------------------------------------------------------------
// near-tie report (diagnostic, not on the hot path).  The library's 2-NN is exact; an fp32 implementation of the
// reference's matcher (cv::BFMatcher / FLANN compare fp32 distances, src/feature_matcher.cpp:45,52) may legitimately
// return another index where two distances differ by less than its rounding error.  flags[q]:
//   bit 0: |d1 - d0| <= rel_tol * d1                (the ORDER of the two nearest may differ)
//   bit 1: some third train row has d <= d1 (1 + rel_tol)   (the second INDEX may differ)
// d = exact L2 distance (fp64 chain of the engine).  Same tiling as knn2_exact_kernel: 64 queries x 64 train rows.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
near_tie_kernel(const float* __restrict__ q, int nq, const float* __restrict__ t, int nt, int dim,
                const int32_t* __restrict__ idx2, const double* __restrict__ d2, double rel_tol, int32_t* __restrict__ third_count)
{
    __shared__ __align__(16) double Qs[KC][PITCH];
    __shared__ __align__(16) double Ts[KC][PITCH];
    __shared__ int cnt[QT];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.x * QT;
    const int per = ((nt + (int)gridDim.y - 1) / (int)gridDim.y + TT - 1) / TT * TT;
    const int t_begin = min(nt, (int)blockIdx.y * per), t_end = min(nt, t_begin + per);
    if (tid < QT) cnt[tid] = 0;
    double lim[4];
    int i0[4], i1[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int qr = q0 + ty * 4 + i;
        lim[i] = -1.0; i0[i] = i1[i] = -1;
        if (qr < nq && idx2[2 * qr + 1] >= 0) {
            const double f = 1.0 + rel_tol;
            lim[i] = d2[2 * qr + 1] * f * f;
            i0[i] = idx2[2 * qr]; i1[i] = idx2[2 * qr + 1];
        }
    }
    int mine[4] = {0, 0, 0, 0};
    for (int t0 = t_begin; t0 < t_end; t0 += TT) {
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
        for (int kc = 0; kc < dim; kc += KC) {
            __syncthreads();
            for (int v = tid; v < QT * (KC / 4); v += 256) {
                int row = v >> 3, c4 = v & 7, k = kc + c4 * 4;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                int qr = q0 + row;
                if (qr < nq && k < dim) a = *reinterpret_cast<const float4*>(q + (size_t)qr * dim + k);
                int tr = t0 + row;
                if (tr < t_end && k < dim) b = *reinterpret_cast<const float4*>(t + (size_t)tr * dim + k);
                Qs[c4 * 4 + 0][row] = (double)a.x; Qs[c4 * 4 + 1][row] = (double)a.y;
                Qs[c4 * 4 + 2][row] = (double)a.z; Qs[c4 * 4 + 3][row] = (double)a.w;
                Ts[c4 * 4 + 0][row] = (double)b.x; Ts[c4 * 4 + 1][row] = (double)b.y;
                Ts[c4 * 4 + 2][row] = (double)b.z; Ts[c4 * 4 + 3][row] = (double)b.w;
            }
            __syncthreads();
            const int klim = min(KC, dim - kc);
#pragma unroll 4
            for (int k = 0; k < klim; k++) {
                double2 qa = *reinterpret_cast<const double2*>(&Qs[k][ty * 4]);
                double2 qb = *reinterpret_cast<const double2*>(&Qs[k][ty * 4 + 2]);
                double2 ta = *reinterpret_cast<const double2*>(&Ts[k][tx * 4]);
                double2 tb = *reinterpret_cast<const double2*>(&Ts[k][tx * 4 + 2]);
                double qv[4] = {qa.x, qa.y, qb.x, qb.y};
                double tv[4] = {ta.x, ta.y, tb.x, tb.y};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        double diff = __dsub_rn(qv[i], tv[j]);
                        acc[i][j] = __fma_rn(diff, diff, acc[i][j]);
                    }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int tr = t0 + tx * 4 + j;
            if (tr < t_end) {
#pragma unroll
                for (int i = 0; i < 4; i++) mine[i] += (acc[i][j] <= lim[i] && tr != i0[i] && tr != i1[i]) ? 1 : 0;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) if (mine[i]) atomicAdd(&cnt[ty * 4 + i], mine[i]);
    __syncthreads();
    if (tid < QT && q0 + tid < nq && cnt[tid]) atomicAdd(&third_count[q0 + tid], cnt[tid]);
}

__global__ void near_tie_flags_kernel(const int32_t* __restrict__ idx2, const double* __restrict__ d2, const int32_t* __restrict__ third_count,
                                      int nq, double rel_tol, uint8_t* __restrict__ flags, int32_t* __restrict__ n_flagged)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint8_t f = 0;
    if (i < nq && idx2[2 * i + 1] >= 0) {
        const double a = sqrt(d2[2 * i]), b = sqrt(d2[2 * i + 1]);
        if (b - a <= rel_tol * b) f |= 1;
        if (third_count[i] > 0) f |= 2;
    }
    if (i < nq) flags[i] = f;
    const int n = __reduce_add_sync(0xffffffffu, f ? 1 : 0);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(n_flagged, n);
}

int near_ties(erp_ctx* ctx, const float* d_q, int nq, const float* d_t, int nt, int dim, const int32_t* d_idx2, const double* d_d2,
              double rel_tol, int32_t* d_third /* nq + 1 */, uint8_t* d_flags)
{
    ERP_CUDA(cudaMemsetAsync(d_third, 0, sizeof(int32_t) * ((size_t)nq + 1), ctx->stream));
    // few query blocks: split the train rows so that the machine is full
    const int qb = cdiv(nq, QT);
    int slices = 1;
    if (qb < ctx->sm_count * 2) slices = max(1, min(cdiv(nt, TT), (ctx->sm_count * 2) / qb));
    near_tie_kernel<<<dim3(qb, slices), 256, 0, ctx->stream>>>(d_q, nq, d_t, nt, dim, d_idx2, d_d2, rel_tol, d_third);
    ERP_LAUNCH(ctx, "near_tie_kernel");
    near_tie_flags_kernel<<<cdiv(nq, 256), 256, 0, ctx->stream>>>(d_idx2, d_d2, d_third, nq, rel_tol, d_flags, d_third + nq);
    ERP_LAUNCH(ctx, "near_tie_flags_kernel");
    return ERP_OK;
}

// ---------------------------------------------------------------------------------------
// ratio test + cross-check + compaction in ascending queryIdx (feature_matcher.cpp:50-56)
// ---------------------------------------------------------------------------------------
constexpr int FB = 1024;  // queries per filter block

__device__ __forceinline__ bool keep_match(const int32_t* idx2, const float* dist2, int i, float ratio,
                                           const int32_t* rev, int q_offset)
{
    // if (m[0].distance < ratio_thresh * m[1].distance)   -- fp32 product, strict <
    // a row of NaN descriptors has no nearest neighbour (idx -1): it never matches
    if (idx2[2 * i] < 0) return false;
    bool keep = ratio < 0.f ? true : (dist2[2 * i] < __fmul_rn(ratio, dist2[2 * i + 1]));
    if (keep && rev) keep = rev[idx2[2 * i]] == i + q_offset;
    return keep;
}

__global__ void filter_count_kernel(const int32_t* idx2, const float* dist2, int nq, float ratio,
                                    const int32_t* rev, int q_offset, int32_t* block_cnt)
{
    int i = blockIdx.x * FB + threadIdx.x;
    int c = 0;
    for (int k = 0; k < FB / 256; k++, i += 256)
        if (i < nq) c += keep_match(idx2, dist2, i, ratio, rev, q_offset);
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < 8; w++) s += ws[w];
        block_cnt[blockIdx.x] = s;
    }
}

__global__ void filter_scan_kernel(int32_t* block_cnt, int nblocks, int32_t* n_out)
{
    // single block: exclusive scan in place
    __shared__ int carry;
    __shared__ int ws[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nblocks ? block_cnt[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = ws[threadIdx.x], s = w;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (threadIdx.x >= o) s += y; }
            ws[threadIdx.x] = s - w;
        }
        __syncthreads();
        int excl = x - v + ws[threadIdx.x >> 5] + carry;
        if (i < nblocks) block_cnt[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = carry;
}

__global__ void filter_write_kernel(const int32_t* idx2, const float* dist2, int nq, float ratio,
                                    const int32_t* rev, int q_offset, const int32_t* block_off,
                                    erp_dmatch* out)
{
    // each thread owns FB/256 CONSECUTIVE queries so that output order is ascending queryIdx
    constexpr int PER = FB / 256;
    int first = blockIdx.x * FB + threadIdx.x * PER;
    bool k[PER];
    int c = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) {
        int i = first + j;
        k[j] = i < nq && keep_match(idx2, dist2, i, ratio, rev, q_offset);
        c += k[j];
    }
    int x = c;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    __shared__ int ws[8];
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += ws[w];
    int pos = block_off[blockIdx.x] + woff + x - c;
#pragma unroll
    for (int j = 0; j < PER; j++) {
        if (k[j]) {
            int i = first + j;
            erp_dmatch m;
            m.queryIdx = i + q_offset; m.trainIdx = idx2[2 * i]; m.imgIdx = 0; m.distance = dist2[2 * i];
            out[pos++] = m;
        }
    }
}

} // namespace erp

using namespace erp;

ERP_API int erp_match_filter_dev(erp_ctx* ctx, const int32_t* d_idx2, const float* d_dist2, int nq,
                                 float ratio, const int32_t* d_rev_best_q, int q_offset,
                                 erp_dmatch* d_out, int32_t* d_n_out)
{
    ERP_ARG(ctx && d_n_out && nq >= 0, ERP_E_ARG, "erp_match_filter_dev: bad argument");
    DeviceGuard g(ctx->device);
    if (nq == 0) {
        ERP_CUDA(cudaMemsetAsync(d_n_out, 0, sizeof(int32_t), ctx->stream));
        return ERP_OK;
    }
    ERP_ARG(d_idx2 && d_dist2 && d_out, ERP_E_ARG, "erp_match_filter_dev: null buffer");
    int nb = cdiv(nq, FB), st = ERP_OK;
    int32_t* cnt = ctx->scratch<int32_t>(S_BLOCKCNT, nb, &st);
    ERP_TRY(st);
    filter_count_kernel<<<nb, 256, 0, ctx->stream>>>(d_idx2, d_dist2, nq, ratio, d_rev_best_q, q_offset, cnt);
    ERP_LAUNCH(ctx, "filter_count_kernel");
    filter_scan_kernel<<<1, 1024, 0, ctx->stream>>>(cnt, nb, d_n_out);
    ERP_LAUNCH(ctx, "filter_scan_kernel");
    filter_write_kernel<<<nb, 256, 0, ctx->stream>>>(d_idx2, d_dist2, nq, ratio, d_rev_best_q, q_offset, cnt, d_out);
    ERP_LAUNCH(ctx, "filter_write_kernel");
    return ERP_OK;
}
