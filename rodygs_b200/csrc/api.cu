// C-ABI glue: version and thread-local error string (include/rodygs_b200.h).
#include <stdarg.h>
#include <string.h>
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
