// Library-level plumbing of the ts_b200 C ABI: version, thread-local error text, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace tsb {
namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

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
}  // namespace tsb

extern "C" int ts_version(void) { return TS_B200_ABI_VERSION; }
extern "C" const char* ts_last_error(void) { return tsb::g_err; }
extern "C" int64_t ts_launch_count(void) { return tsb::g_launches.load(std::memory_order_relaxed); }
extern "C" void ts_reset_launch_count(void) { tsb::g_launches.store(0, std::memory_order_relaxed); }
