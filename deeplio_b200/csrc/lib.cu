// Library-level entry points: version, error text, device check, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace dlio {

int g_nvtx = -1;   // NVTX ranges (kernel class names) around every profiled launch group; DLIO_NVTX=1 / "nvtx" option
static bool nvtx_on() {
    if (g_nvtx < 0) {
        const char *e = getenv("DLIO_NVTX");
        g_nvtx = (e && e[0] == '1') ? 1 : 0;
    }
    return g_nvtx != 0;
}
static const char *const kProfNames[] = {"dlio/conv_fwd_simt", "dlio/conv_dgrad_simt", "dlio/conv_wgrad_simt",
                                         "dlio/conv_fwd_tc",   "dlio/conv_dgrad_tc",   "dlio/conv_wgrad_tc",
                                         "dlio/elementwise",   "dlio/dense",           "dlio/rnn",
                                         "dlio/optim"};

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional per-kernel-class timing (CUDA events on the launching stream), used by bench.py
struct ProfRec {
    int kind;
    cudaEvent_t a, b;
};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

int prof_begin(int kind, cudaStream_t st) {
    if (nvtx_on()) nvtxRangePushA(kind >= 0 && kind < 10 ? kProfNames[kind] : "dlio");
    if (!g_prof_on.load(std::memory_order_relaxed)) return -1;
    ProfRec r;
    r.kind = kind;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
    cudaEventRecord(r.a, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
    return (int)g_prof.size() - 1;
}
void prof_end(int idx, cudaStream_t st) {
    if (g_nvtx > 0) nvtxRangePop();
    if (idx < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (idx < (int)g_prof.size()) cudaEventRecord(g_prof[idx].b, st);
}

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

extern "C" int dlio_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(dlio::g_prof_mu);
    for (auto &r : dlio::g_prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    dlio::g_prof.clear();
    dlio::g_prof_on.store(on ? 1 : 0);
    return DLIO_OK;
}

extern "C" int dlio_profile_read(int kind, double *total_ms, long long *launches) {
    DLIO_CHECK_ARG(total_ms && launches, "profile_read: null pointer");
    std::lock_guard<std::mutex> lk(dlio::g_prof_mu);
    double ms = 0.0;
    long long n = 0;
    for (auto &r : dlio::g_prof) {
        if (r.kind != kind) continue;
        DLIO_CUDA(cudaEventSynchronize(r.b));
        float t = 0.f;
        DLIO_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
        ms += t;
        ++n;
    }
    *total_ms = ms;
    *launches = n;
    return DLIO_OK;
}
