// Error plumbing and library information of libtefem_b200.
#include "tefem_common.cuh"

#include <stdarg.h>
#include <atomic>

namespace tefem {

static thread_local char g_last_error[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

}  // namespace tefem

extern "C" int tefem_last_error(char* h_buf, size_t n) {
    if (!h_buf || n == 0) return TEFEM_ERR_INVALID;
    strncpy(h_buf, tefem::g_last_error, n - 1);
    h_buf[n - 1] = '\0';
    return TEFEM_OK;
}

extern "C" int tefem_version(char* h_buf, size_t n) {
    if (!h_buf || n == 0) return TEFEM_ERR_INVALID;
    snprintf(h_buf, n, "tefem_b200 0.1.0 sm_100a cuda %d.%d", CUDART_VERSION / 1000, (CUDART_VERSION % 1000) / 10);
    return TEFEM_OK;
}

extern "C" int64_t tefem_launch_count(void) { return (int64_t)tefem::g_launches.load(); }
