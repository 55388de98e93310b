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

// Opt a kernel in to `bytes` of dynamic shared memory once per (device, kernel).  The "already done" flags are atomics:
// the entry points may be called from several host threads at once, and configuring twice is harmless.
int configure_dynamic_smem(const void *fn, size_t bytes, bool max_carveout, SmemOnceFlags &flags) {
    int dev = 0;
    GSTEX_CUDA_OK(cudaGetDevice(&dev));
    const bool tracked = dev >= 0 && dev < SmemOnceFlags::MAX_DEVICES;
    if (tracked && flags.done[dev].load(std::memory_order_acquire)) return GSTEX_OK;
    GSTEX_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (max_carveout) GSTEX_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (tracked) flags.done[dev].store(true, std::memory_order_release);
    return GSTEX_OK;
}
}  // namespace gstex

extern "C" const char *gstex_last_error(void) { return gstex::g_err; }
extern "C" int gstex_abi_version(void) { return 1; }
