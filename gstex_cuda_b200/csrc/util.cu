// util.cu -- thread-local error string and ABI version.
#include <stdarg.h>

#include "common.cuh"

namespace gstex {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace gstex

extern "C" const char *gstex_last_error(void) { return gstex::g_err; }
extern "C" int gstex_abi_version(void) { return 1; }
