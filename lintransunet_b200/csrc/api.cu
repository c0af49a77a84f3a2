// Process-wide plumbing of libltu_b200.so: error string, SM count cache, launch counter.
#include <atomic>
#include <mutex>
#include <stdarg.h>

#include "common.cuh"

namespace ltu {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static int cache[64];
    static std::once_flag once;
    std::call_once(once, [] { for (int& c : cache) c = 0; });
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev] = n;
    }
    return cache[dev];
}

}  // namespace ltu

extern "C" const char* ltu_last_error(void) { return ltu::g_err; }
extern "C" int ltu_version(void) { return 100; }   // 0.1.0
extern "C" int64_t ltu_launch_count(void) { return (int64_t)ltu::g_launches.load(); }
