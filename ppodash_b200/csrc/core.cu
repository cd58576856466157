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

int ppd_upload_rows(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, size_t rows, void* stream) {
    PPD_REQUIRE(dst && src && row_bytes > 0 && rows > 0 && dst_pitch >= row_bytes && src_pitch >= row_bytes, "upload_rows: bad arguments");
    cudaError_t e = cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) { ppd::set_error("upload_rows: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
    return 0;
}
}
