// Library-level entry points: version, error text, device check, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace dlio {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace dlio

extern "C" int dlio_abi_version(void) { return DLIO_ABI_VERSION; }
extern "C" const char *dlio_last_error(void) { return dlio::g_err; }
extern "C" long long dlio_launch_count(void) { return dlio::g_launches.load(std::memory_order_relaxed); }

extern "C" int dlio_device_check(int device) {
    cudaDeviceProp p;
    DLIO_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        dlio::set_error("device %d is sm_%d%d; deeplio_b200 is built for sm_100a only", device, p.major, p.minor);
        return DLIO_ERR_UNSUPPORTED;
    }
    return DLIO_OK;
}
