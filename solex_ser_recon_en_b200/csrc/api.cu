// libshg: status, version, device query.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void shg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* shg_last_error(void) { return g_err; }

extern "C" int shg_version(void) { return SHG_VERSION; }

extern "C" int shg_device_info(int device, int64_t* out6) {
    cudaDeviceProp p;
    SHG_CHECK(cudaGetDeviceProperties(&p, device));
    out6[0] = p.multiProcessorCount;
    out6[1] = p.major;
    out6[2] = p.minor;
    out6[3] = (int64_t)p.totalGlobalMem;
    out6[4] = p.l2CacheSize;
    out6[5] = (int64_t)p.sharedMemPerBlockOptin;
    return 0;
}
