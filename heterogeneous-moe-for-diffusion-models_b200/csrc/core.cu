// Library-level state: last error text, version, launch counter.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace hdmoe {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace hdmoe

extern "C" int hdmoe_version(void) { return 101; }
extern "C" const char* hdmoe_last_error(void) { return hdmoe::g_err; }
extern "C" int64_t hdmoe_launch_count(void) { return hdmoe::g_launches.load(std::memory_order_relaxed); }
