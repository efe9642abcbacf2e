// knn_tc.cu -- tcgen05 3xTF32 distance engine (placeholder until the kernel lands).
#include "common.cuh"

namespace erp {

bool knn2_tc_supported(int, int, int) { return false; }

int knn2_tc(erp_ctx*, const float*, int, const float*, int, int, int32_t*, float*, double*)
{
    set_error("tcgen05 engine not built");
    return ERP_E_DIM;
}

} // namespace erp
