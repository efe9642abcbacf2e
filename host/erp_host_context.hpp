// One lazily created erp_ctx per host thread (the reference constructs feature_matcher and
// eight_point on the stack per call -- src/spherical_surf.cpp:96, src/automatic.cpp:124 -- so the
// classes themselves must stay cheap; the CUDA context and scratch live here).
#pragma once
#include <stdexcept>
#include <string>

#include "erp_b200.h"

namespace erp_host {

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& what) : std::runtime_error(what), status(st) {}
};

erp_ctx* context();                              // device from $ERP_B200_DEVICE (default 0)
void check(int status, const char* where);       // throws erp_host::Error on status != 0

} // namespace erp_host
