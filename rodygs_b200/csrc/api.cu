// C-ABI glue: version and thread-local error string (include/rodygs_b200.h).
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "common.cuh"

static thread_local char g_err[512] = "";

void rdg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int rdg_abi_version(void) { return RDG_ABI_VERSION; }
extern "C" const char* rdg_last_error(void) { return g_err; }

static std::atomic<uint64_t> g_launches{0};
void rdg_count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
extern "C" uint64_t rdg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
