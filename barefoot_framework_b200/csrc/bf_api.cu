// Version / error plumbing of the C ABI (include/barefoot_sm100.h).
#include <stdarg.h>
#include <string.h>

#include "bf_common.cuh"

static thread_local char g_err[512] = "";

void bf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int bf_version(void) { return BF_VERSION; }
extern "C" const char* bf_last_error(void) { return g_err; }
