// Library-level state: last error text, version, launch counter.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>

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

// Dynamic-tile-scheduler counters.  A persistent kernel that hands out tiles with atomicAdd needs a zeroed counter
// per launch; the kernels re-arm it themselves (last CTA out), so a counter may only be shared by launches that
// cannot overlap: one slot per (device, stream).  Launches on one stream are serialised, and kernels captured from
// one stream into a CUDA graph keep that dependency chain.  Zero-initialised device memory, no allocation, no
// stream operation: safe to call under stream capture.
__device__ int32_t g_sched[kSchedSlots * 2];
static std::mutex g_sched_mu;
static int g_sched_n = 0;
static struct { int dev; cudaStream_t st; int32_t* ptr; } g_sched_tab[kSchedSlots];
static int32_t* g_sched_base[16] = {nullptr};

int32_t* sched_slot(cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    std::lock_guard<std::mutex> lk(g_sched_mu);
    for (int i = 0; i < g_sched_n; ++i)
        if (g_sched_tab[i].dev == dev && g_sched_tab[i].st == st) return g_sched_tab[i].ptr;
    if (g_sched_n >= kSchedSlots) return nullptr;
    if (!g_sched_base[dev]) {
        void* sym = nullptr;
        if (cudaGetSymbolAddress(&sym, g_sched) != cudaSuccess) return nullptr;
        g_sched_base[dev] = (int32_t*)sym;
    }
    int used = 0;                                   // slots already handed out on this device
    for (int i = 0; i < g_sched_n; ++i) used += g_sched_tab[i].dev == dev;
    g_sched_tab[g_sched_n] = {dev, st, g_sched_base[dev] + 2 * used};
    return g_sched_tab[g_sched_n++].ptr;
}
}  // namespace hdmoe

extern "C" int hdmoe_version(void) { return 101; }
extern "C" const char* hdmoe_last_error(void) { return hdmoe::g_err; }
extern "C" int64_t hdmoe_launch_count(void) { return hdmoe::g_launches.load(std::memory_order_relaxed); }
