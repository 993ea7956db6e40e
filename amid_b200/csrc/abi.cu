// Error plumbing and version of the C ABI (include/amid_b200.h).
#include "common.cuh"

namespace amid {
static thread_local char g_err[512] = "";
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace amid

extern "C" const char* amid_last_error(void) { return amid::g_err; }
extern "C" int amid_version(void) { return 100; }
