// Library-level entry points: version, error reporting, architecture gate.
#include <stdarg.h>
#include <atomic>
#include <stdlib.h>

#include "common.cuh"

namespace td {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int require_sm100() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_status = TD_ERR_ARCH;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device: %s (libtinydiff has no CPU fallback)", cudaGetErrorString(e));
        return TD_ERR_ARCH;
    }
    if (dev == cached_dev) {
        if (cached_status != TD_OK) set_error("device %d is not sm_100-class (libtinydiff is sm_100a only)", dev);
        return cached_status;
    }
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cached_dev = dev;
    cached_status = (major == 10) ? TD_OK : TD_ERR_ARCH;
    if (cached_status != TD_OK)
        set_error("device %d is sm_%d%d; libtinydiff is built for sm_100a only and has no fallback", dev, major, minor);
    return cached_status;
}

static int g_pdl = -1;      // -1: not read yet
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool pdl_enabled() {
    if (g_pdl < 0) {
        const char* e = getenv("TD_PDL");
        g_pdl = (e && atoi(e) == 0) ? 0 : 1;
    }
    return g_pdl != 0;
}

static int g_sm_budget = kNumSMs;
int sm_budget() { return g_sm_budget; }

}  // namespace td

extern "C" int td_set_sm_budget(int sms) {
    const int prev = td::g_sm_budget;
    if (sms >= 16 && sms <= td::kNumSMs) td::g_sm_budget = sms;
    return prev;
}

extern "C" int td_set_pdl(int on) {
    const int prev = td::pdl_enabled() ? 1 : 0;
    td::g_pdl = on ? 1 : 0;
    return prev;
}

extern "C" int64_t td_launch_count(void) { return (int64_t)td::g_launches.load(std::memory_order_relaxed); }

extern "C" int td_version(void) { return 100; }
extern "C" const char* td_last_error_string(void) { return td::g_err; }

extern "C" int td_device_check(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        td::set_error("device %d not available (%s)", device, e != cudaSuccess ? cudaGetErrorString(e) : "out of range");
        return TD_ERR_ARCH;
    }
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (major != 10) {
        td::set_error("device %d has compute capability major %d; need 10 (B200)", device, major);
        return TD_ERR_ARCH;
    }
    return TD_OK;
}
