#include "erp_host_context.hpp"

#include <cstdlib>
#include <vector>

namespace erp_host {

namespace {
struct Runtime {
    erp_ctx* ctx = nullptr;
    erp_group* grp = nullptr;
    std::mutex mu;
    bool ready = false;
    ~Runtime()
    {
        if (grp) erp_group_destroy(grp);
        else if (ctx) erp_ctx_destroy(ctx);
    }
};

Runtime& runtime()
{
    static Runtime rt;             // constructed once (C++11), destroyed at exit before the CUDA runtime goes away
    return rt;
}

void ensure(Runtime& rt)
{
    static std::once_flag once;
    std::call_once(once, [&rt] {
        std::vector<int> devices;
        if (const char* list = std::getenv("ERP_B200_DEVICES")) {
            for (const char* p = list; *p;) {
                char* end = nullptr;
                long v = std::strtol(p, &end, 10);
                if (end == p) break;
                devices.push_back((int)v);
                p = *end == ',' ? end + 1 : end;
            }
        }
        if (devices.size() > 1) {
            check(erp_group_create(devices.data(), (int)devices.size(), &rt.grp), "erp_group_create");
            rt.ctx = erp_group_ctx(rt.grp, 0);
        } else {
            const char* one = std::getenv("ERP_B200_DEVICE");
            const int dev = !devices.empty() ? devices[0] : (one ? std::atoi(one) : 0);
            check(erp_ctx_create(dev, &rt.ctx), "erp_ctx_create");
        }
        rt.ready = true;
    });
    if (!rt.ready) throw Error(ERP_E_NO_DEVICE, "erp_host: the B200 runtime could not be created (see the first error)");
}
} // namespace

erp_ctx* context()
{
    Runtime& rt = runtime();
    ensure(rt);
    return rt.ctx;
}

erp_group* group()
{
    Runtime& rt = runtime();
    ensure(rt);
    return rt.grp;
}

std::mutex& mutex() { return runtime().mu; }

void check(int status, const char* where)
{
    if (status != ERP_OK) throw Error(status, std::string(where) + ": " + erp_last_error());
}

} // namespace erp_host
