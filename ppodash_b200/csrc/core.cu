// Library bookkeeping: ABI version, last-error string, launch counter.
#include <stdarg.h>

#include "ppd_common.cuh"

namespace ppd {
static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }
}  // namespace ppd

extern "C" {
int ppd_abi_version(void) { return PPD_ABI_VERSION; }
const char* ppd_last_error(void) { return ppd::g_err; }
int64_t ppd_launch_count(void) { return ppd::g_launches; }
void ppd_reset_launch_count(void) { ppd::g_launches = 0; }
}
