// libshg: status, version, device query.
#include <stdarg.h>

#include <algorithm>
#include <vector>

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

// ---- host helper: 8-connected components of a sparse, sorted pixel list -----
// (scipy.ndimage.label semantics on the thin-edge pixels of the limb search,
// reference ellipse_to_circle.py:252: labels 1.. numbered in raster order of each
// component's first pixel).  flat = row*cols + col, strictly ascending.  Pure
// host code: the list has ~10^4 entries.
extern "C" int shg_label_points(const int64_t* flat, int64_t n, int64_t cols, int32_t* labels, int32_t* n_labels) {
    SHG_REQUIRE(flat && labels && n_labels && n >= 0 && cols > 0, "shg_label_points: bad arguments");
    std::vector<int32_t> parent((size_t)n);
    auto find = [&](int32_t a) {
        while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; }
        return a;
    };
    auto unite = [&](int32_t a, int32_t b) {
        a = find(a); b = find(b);
        if (a != b) parent[a > b ? a : b] = a > b ? b : a;          // keep the earliest pixel as the root
    };
    int64_t j = 0;                                                   // first point that can be an upper neighbour
    for (int64_t i = 0; i < n; ++i) {
        SHG_REQUIRE(i == 0 || flat[i] > flat[i - 1], "shg_label_points: indices must be strictly ascending");
        parent[i] = (int32_t)i;
        const int64_t col = flat[i] % cols;
        if (i > 0 && col > 0 && flat[i - 1] == flat[i] - 1) unite((int32_t)i, (int32_t)(i - 1));
        const int64_t lo = flat[i] - cols - (col > 0 ? 1 : 0), hi = flat[i] - cols + (col + 1 < cols ? 1 : 0);
        while (j < i && flat[j] < lo) ++j;
        for (int64_t q = j; q < i && flat[q] <= hi; ++q) unite((int32_t)i, (int32_t)q);
    }
    int32_t next = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t r = find((int32_t)i);
        if (r == (int32_t)i) labels[i] = ++next;                     // roots are first pixels: raster order
        else labels[i] = labels[r];
    }
    *n_labels = next;
    return 0;
}

// ---- host helper: scatter matrix of the conic design vectors -----------------
// S = sum over points of d d^T with d = [x^2, xy, y^2, x, y, 1] (the Halir-Flusser direct ellipse fit the reference
// makes through lsq-ellipse, ellipse_to_circle.py:57-59: S1 = S[:3,:3], S2 = S[:3,3:], S3 = S[3:,3:]).  Accumulated
// in long double: at least as accurate as the float64 matrix products it replaces (which took 0.3 ms of NumPy
// temporaries per fit on the critical path of every multi-GPU step; this loop takes ~30 us for 10^4 points).
extern "C" int shg_conic_scatter(const double* h_xy, int64_t n, double* h_s36) {
    SHG_REQUIRE(h_xy && h_s36 && n >= 0, "shg_conic_scatter: bad arguments");
    // blocked summation: partial sums of 256 points in double, totals in long double (x87 arithmetic on every
    // product cost 0.6 ms per fit; this keeps its accuracy where it matters -- the long running sum -- at SSE speed)
    long double acc[21];
    for (int k = 0; k < 21; ++k) acc[k] = 0.0L;
    for (int64_t i0 = 0; i0 < n; i0 += 256) {
        double part[21];
        for (int k = 0; k < 21; ++k) part[k] = 0.0;
        const int64_t i1 = std::min<int64_t>(n, i0 + 256);
        for (int64_t i = i0; i < i1; ++i) {
            const double x = h_xy[2 * i], y = h_xy[2 * i + 1];
            const double d[6] = {x * x, x * y, y * y, x, y, 1.0};     // rounded to double like NumPy's design matrix
            int k = 0;
            for (int a = 0; a < 6; ++a)
                for (int b = a; b < 6; ++b) part[k++] += d[a] * d[b];
        }
        for (int k = 0; k < 21; ++k) acc[k] += (long double)part[k];
    }
    int k = 0;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) {
            h_s36[a * 6 + b] = h_s36[b * 6 + a] = (double)acc[k++];
        }
    return 0;
}
