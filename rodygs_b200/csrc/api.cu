// C-ABI glue: version and thread-local error string (include/rodygs_b200.h).
#include <stdarg.h>
#include <stdlib.h>
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

// Tunables: A/B switches and test knobs of the launchers.  Initial values come from the environment (RDG_<NAME>),
// rdg_set_tunable() overrides them at run time (tests use it to drive the multi-chunk paths of the persistent
// kernels at sizes the CPU oracle can check).
static const char* const g_tun_names[RDG_TUN_COUNT] = {"pre_grid_cap", "dtable_v1", "diff_smem", "deterministic", "sm_reserve", "ar_unroll", "l2_prefetch"};
static const char* const g_tun_env[RDG_TUN_COUNT] = {"RDG_PRE_GRID_CAP", "RDG_DTABLE_V1", "RDG_DIFF_SMEM", "RDG_DETERMINISTIC", "RDG_SM_RESERVE", "RDG_AR_UNROLL", "RDG_L2_PREFETCH"};
static const int g_tun_default[RDG_TUN_COUNT] = {0, 0, 1, 0, 0, 4, 0};
static std::atomic<int> g_tun[RDG_TUN_COUNT];
static std::atomic<bool> g_tun_init{false};

static void tun_init() {
    if (g_tun_init.load(std::memory_order_acquire)) return;
    for (int i = 0; i < RDG_TUN_COUNT; ++i) {
        const char* e = getenv(g_tun_env[i]);
        g_tun[i].store(e && e[0] ? atoi(e) : g_tun_default[i], std::memory_order_relaxed);
    }
    g_tun_init.store(true, std::memory_order_release);
}

int rdg_tunable(int id) {
    tun_init();
    return g_tun[id].load(std::memory_order_relaxed);
}

extern "C" int rdg_set_tunable(const char* name, int32_t value) {
    tun_init();
    if (name)
        for (int i = 0; i < RDG_TUN_COUNT; ++i)
            if (strcmp(name, g_tun_names[i]) == 0) {
                g_tun[i].store(value, std::memory_order_relaxed);
                return RDG_OK;
            }
    rdg_set_error("rdg_set_tunable: unknown tunable '%s'", name ? name : "(null)");
    return RDG_E_ARG;
}
