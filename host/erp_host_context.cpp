#include "erp_host_context.hpp"

#include <cstdlib>

namespace erp_host {

namespace {
struct Holder {
    erp_ctx* ctx = nullptr;
    ~Holder() { if (ctx) erp_ctx_destroy(ctx); }
};
}

erp_ctx* context()
{
    static thread_local Holder h;
    if (!h.ctx) {
        const char* env = std::getenv("ERP_B200_DEVICE");
        int dev = env ? std::atoi(env) : 0;
        check(erp_ctx_create(dev, &h.ctx), "erp_ctx_create");
    }
    return h.ctx;
}

void check(int status, const char* where)
{
    if (status != ERP_OK) throw Error(status, std::string(where) + ": " + erp_last_error());
}

} // namespace erp_host
