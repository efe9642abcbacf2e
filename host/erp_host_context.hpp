// One process-wide B200 runtime behind the drop-in classes.
//
// The reference constructs feature_matcher and eight_point on the stack per call (src/spherical_surf.cpp:96,
// src/automatic.cpp:124) and calls erp_rotation::rotate_pixel from inside `omp parallel for` loops
// (src/spherical_surf.cpp:27-45): the classes must stay cheap and must not create a CUDA context per object or per
// OpenMP thread.  The runtime (one erp_ctx, or one erp_group when $ERP_B200_DEVICES names several GPUs) is created on
// first use and shared; a context is single threaded, so every call into the C ABI holds the runtime's mutex.
//
//   ERP_B200_DEVICES=0,1,2,3   the GPUs one call fans out over (query rows per GPU, NCCL inside liberp_b200.so)
//   ERP_B200_DEVICE=2          a single GPU (default 0)
#pragma once
#include <mutex>
#include <stdexcept>
#include <string>

#include "erp_b200.h"

namespace erp_host {

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& what) : std::runtime_error(what), status(st) {}
};

erp_ctx* context();                              // the process context (rank 0 of the group when there is one)
erp_group* group();                              // nullptr unless $ERP_B200_DEVICES lists more than one device
std::mutex& mutex();                             // held around every C-ABI call made through context() / group()
void check(int status, const char* where);       // throws erp_host::Error on status != 0

struct Lock {
    std::lock_guard<std::mutex> guard;
    Lock() : guard(mutex()) {}
};

} // namespace erp_host
