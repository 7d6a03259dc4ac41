// runtime.cu -- error string, launch counter, device queries for libpointrcnn_b200.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace prb {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;  // B200
    }
    return sms;
}
}  // namespace prb

extern "C" int prb_abi_version(void) { return PRB_ABI_VERSION; }
extern "C" const char *prb_last_error(void) { return prb::g_err; }
extern "C" unsigned long long prb_launch_count(void) { return prb::g_launches.load(); }
