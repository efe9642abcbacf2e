// common.cuh -- context, error plumbing and scratch management shared by the .cu files.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "erp_b200.h"

#define ERP_API extern "C" __attribute__((visibility("default")))

namespace erp {

void set_error(const char* fmt, ...);

#define ERP_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            erp::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return ERP_E_CUDA;                                                           \
        }                                                                                \
    } while (0)

#define ERP_TRY(call)                 \
    do {                              \
        int s_ = (call);              \
        if (s_ != ERP_OK) return s_;  \
    } while (0)

#define ERP_ARG(cond, code, ...)       \
    do {                               \
        if (!(cond)) {                 \
            erp::set_error(__VA_ARGS__); \
            return (code);             \
        }                              \
    } while (0)

// grow-only device / pinned-host buffer
struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    bool host = false;
    int reserve(size_t bytes);
    void release();
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

enum ScratchId {
    S_Q = 0, S_T, S_IDX2, S_DIST2, S_D2, S_REVQ, S_REVD2, S_OUT, S_NOUT, S_BLOCKCNT,
    S_L3, S_R3, S_L4, S_R4, S_XY, S_SAMPLES, S_GRAM, S_E, S_EF, S_POSE, S_COUNTS, S_PACKED,
    S_MASK, S_PARTIAL, S_CONS, S_MISC,
    S_TC_Q, S_TC_T, S_TC_QN, S_TC_TN, S_TC_CAND, S_TC_LIST, S_TC_MISC,
    S_RS_IDX, S_RS_DIST, S_RS_D2, S_RS_PARTIAL,
    S_SC_E, S_SC_E2, S_SC_K, S_SC_MISC, S_SC_BOUNDS, S_SC_LIST,
    S_IMG_IN, S_IMG_OUT,
    S_COUNT_
};

} // namespace erp

struct erp_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;             // host-buffer calls upload query chunks here while the previous chunk computes
    cudaEvent_t ev_copy[8] = {};                    // "chunk c is on the device" (+ one fork event)
    int tc_chunk = 0;                               // > 0: later query chunk of one host call: train operand and statistics carry over
    int engine = ERP_ENGINE_AUTO;
    uint64_t launches = 0;
    int64_t knn_stats[5] = {0, 0, 0, 0, 0};
    int32_t* sc_misc_dev = nullptr;                 // device words of the last tensor-core best search: ., max c_lo, list length
    int32_t* tc_misc_dev = nullptr;                 // device words of the last tcgen05 call: re-scan count, ., deviation
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;   // around the dominant distance kernel
    std::vector<cudaEvent_t> ev_score;              // pairs around the scoring kernel launches of the last RANSAC call
    int n_ev_score = 0;                             // events used by that call
    erp::Buf dev[erp::S_COUNT_];
    erp::Buf pinned[8];

    template <class T> T* scratch(int id, size_t count, int* status) {
        int s = dev[id].reserve(count * sizeof(T));
        if (s != ERP_OK) *status = s;
        return dev[id].as<T>();
    }
    template <class T> T* host_scratch(int id, size_t count, int* status) {
        pinned[id].host = true;
        int s = pinned[id].reserve(count * sizeof(T));
        if (s != ERP_OK) *status = s;
        return pinned[id].as<T>();
    }
};

namespace erp {

inline int check_launch(erp_ctx* ctx, const char* what)
{
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return ERP_E_CUDA;
    }
    return ERP_OK;
}

#define ERP_LAUNCH(ctx, name) ERP_TRY(erp::check_launch((ctx), (name)))

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// next event of the scoring-kernel timer (created on demand, reused across calls)
inline int score_event(erp_ctx* ctx, cudaEvent_t* ev)
{
    if (ctx->n_ev_score == (int)ctx->ev_score.size()) {
        cudaEvent_t e;
        ERP_CUDA(cudaEventCreate(&e));
        ctx->ev_score.push_back(e);
    }
    *ev = ctx->ev_score[ctx->n_ev_score++];
    ERP_CUDA(cudaEventRecord(*ev, ctx->stream));
    return ERP_OK;
}

// ---- internal device-level entry points shared across translation units ----
int knn2_exact(erp_ctx* ctx, const float* d_q, int nq, const float* d_t, int nt, int dim,
               const int32_t* d_qlist, int nlist, int idx_offset,
               int32_t* d_idx2, float* d_dist2, double* d_d2);
int knn2_tc(erp_ctx* ctx, const float* d_q, int nq, const float* d_t, int nt, int dim,
            int32_t* d_idx2, float* d_dist2, double* d_d2);
bool knn2_tc_supported(int nq, int nt, int dim);
bool knn2_tc_preferred(int nq, int nt, int dim);
bool knn2_tc1_preferred(int nq, int nt, int dim);
int knn2_tc1(erp_ctx* ctx, const float* d_q, int nq, const float* d_t, int nt, int dim,
             int32_t* d_idx2, float* d_dist2, double* d_d2);
int tc_misc_begin(erp_ctx* ctx, int32_t* misc);
int tc_misc_end(erp_ctx* ctx, int32_t* misc);
int refine_launch(erp_ctx* ctx, const float* d_q, int nq, const float* d_t, int nt, int dim, int n_lists, int topk, double kappa,
                  const int32_t* cand, const float* cand_s, const float* cand_thr, unsigned* misc, int32_t* d_idx2, float* d_dist2,
                  double* d_d2, int32_t* rescan_list);
int knn2_exact_rescan(erp_ctx* ctx, const float* d_q, int nq, const float* d_t, int nt, int dim,
                      const int32_t* d_list, const int32_t* d_count, int max_rows,
                      int32_t* d_idx2, float* d_dist2, double* d_d2);
int score_tc_best(erp_ctx* ctx, const double* d_E, int H, const float* d_l4, const float* d_r4, int m, float tau,
                  uint64_t hyp0, int32_t* d_counts_scratch, uint64_t* d_best);
int score_list_best(erp_ctx* ctx, const double* d_E, int H_max, const int32_t* d_list, const int32_t* d_len,
                    const float* d_l4, const float* d_r4, int m, float tau, uint64_t hyp0,
                    int32_t* d_counts, uint64_t* d_best);
bool score_tc_preferred(int H, int m);

} // namespace erp
