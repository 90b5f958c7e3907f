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

// ---- buffers that the other ranks of the box write into (row exchange) ------
extern "C" int shg_ipc_alloc(int64_t bytes, void** d_ptr, unsigned char* handle64) {
    SHG_REQUIRE(d_ptr && handle64 && bytes > 0, "shg_ipc_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == SHG_IPC_HANDLE_BYTES, "IPC handle size");
    SHG_CHECK(cudaMalloc(d_ptr, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *d_ptr);
    if (e != cudaSuccess) {
        cudaFree(*d_ptr);
        shg_set_error("shg_ipc_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
        return 1;
    }
    memcpy(handle64, &h, sizeof(h));
    return 0;
}

extern "C" int shg_ipc_free(void* d_ptr) {
    if (d_ptr) SHG_CHECK(cudaFree(d_ptr));
    return 0;
}

extern "C" int shg_ipc_open(const unsigned char* handle64, void** d_ptr) {
    SHG_REQUIRE(d_ptr && handle64, "shg_ipc_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    SHG_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int shg_ipc_close(void* d_ptr) {
    if (d_ptr) SHG_CHECK(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}
