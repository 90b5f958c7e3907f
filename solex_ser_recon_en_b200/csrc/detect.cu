// Spectral-line detection on the mean frame and the cubic least-squares fit.
// Replaces reference solex_util.py:165-172 (detect_bord), :228-231,242 (blurred
// and sharp per-row minima) and the three np.polyfit calls of :233-255.
// All inputs are the (ih x iw) mean / max images: O(ih*iw) work, negligible next
// to the stack passes, so these kernels favour exactness over tuning.
#include "common.cuh"

namespace {

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// horizontal box sums, uint16 -> uint32
__global__ void __launch_bounds__(256)
hsum_kernel(const uint16_t* __restrict__ img, int rows, int cols, int kw, uint32_t* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)rows * cols) return;
    const int r = (int)(idx / cols), c = (int)(idx % cols);
    const uint16_t* row = img + (int64_t)r * cols;
    const int a = c - kw / 2;
    uint32_t s = 0;
    for (int t = 0; t < kw; ++t) s += row[reflect101(a + t, cols)];
    out[idx] = s;
}

// vertical box sums + OpenCV's normalisation (see shg.h)
__global__ void __launch_bounds__(256)
vsum_scale_kernel(const uint32_t* __restrict__ hs, int rows, int cols, int kh, float scale_f, double scale_d,
                  int simd_cols, uint16_t* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)rows * cols) return;
    const int r = (int)(idx / cols), c = (int)(idx % cols);
    const int a = r - kh / 2;
    uint32_t s = 0;
    for (int t = 0; t < kh; ++t) s += hs[(int64_t)reflect101(a + t, rows) * cols + c];
    int v;
    if (c < simd_cols)
        v = __float2int_rn(__fmul_rn(__int2float_rn((int)s), scale_f));
    else
        v = __double2int_rn(__dmul_rn((double)(int)s, scale_d));
    out[idx] = (uint16_t)min(max(v, 0), 65535);
}

// one warp per row
__global__ void __launch_bounds__(256)
row_sums_kernel(const uint16_t* __restrict__ img, int rows, int cols, unsigned long long* __restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const uint16_t* p = img + (int64_t)row * cols;
    unsigned long long s = 0;
    for (int c = lane; c < cols; c += 32) s += p[c];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = s;
}

// one warp per row; first minimum wins (np.argmin)
__global__ void __launch_bounds__(256)
row_argmin_kernel(const uint16_t* __restrict__ img, int rows, int cols, int c0, int c1, int32_t* __restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const uint16_t* p = img + (int64_t)row * cols;
    uint32_t best = 0xffffffffu;            // (value << 16 | column) orders by value then column
    for (int c = c0 + lane; c < c1; c += 32) best = min(best, ((uint32_t)p[c] << 16) | (uint32_t)(c & 0xffff));
    for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) out[row] = (int32_t)(best & 0xffffu);
}

// ---- block-wide fp64 reductions with warp shuffles -------------------------
constexpr int kFitThreads = 1024;

template <int NV>
__device__ void block_reduce_sum(double (&v)[NV], double* smem /* 32*NV */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j)
        for (int o = 16; o; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int j = 0; j < NV; ++j) smem[j * 32 + warp] = v[j];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            double t = (lane < (int)(blockDim.x >> 5)) ? smem[j * 32 + lane] : 0.0;
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) smem[j * 32] = t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = smem[j * 32];
    __syncthreads();
}

__device__ __forceinline__ double horner3(const double* p, double x) {
    // numpy.polynomial.polynomial.polyval: c0 = c[-1]; c0 = c[-i] + c0*x  (no fma)
    double c = p[3];
    c = __dadd_rn(p[2], __dmul_rn(c, x));
    c = __dadd_rn(p[1], __dmul_rn(c, x));
    c = __dadd_rn(p[0], __dmul_rn(c, x));
    return c;
}

// Masked cubic fit.  Moments are taken in u = (x - xc) / xh (|u| <= 1) so the
// 4x4 normal equations are well conditioned; the solution is expanded back to
// raw-x monomials.  Agrees with np.polyfit (SVD lstsq) to ~1e-10 relative.
__global__ void __launch_bounds__(kFitThreads)
polyfit3_kernel(const int32_t* __restrict__ y, const uint8_t* __restrict__ mask, int x0, int n,
                double* __restrict__ coef, const int32_t* __restrict__ y_resid, double* __restrict__ resid) {
    __shared__ double red[32 * 11];
    __shared__ double sol[4];
    const double xc = x0 + 0.5 * (n - 1);
    const double xh = n > 1 ? 0.5 * (n - 1) : 1.0;
    double m[11];                               // S0..S6, T0..T3
#pragma unroll
    for (int j = 0; j < 11; ++j) m[j] = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (mask && !mask[i]) continue;
        const double u = ((double)(x0 + i) - xc) / xh;
        const double yy = (double)y[i];
        double p = 1.0;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            m[k] += p;
            if (k < 4) m[7 + k] += p * yy;
            p *= u;
        }
    }
    block_reduce_sum<11>(m, red);
    if (threadIdx.x == 0) {
        double A[4][5];
        for (int r = 0; r < 4; ++r) {
            for (int c = 0; c < 4; ++c) A[r][c] = m[r + c];
            A[r][4] = m[7 + r];
        }
        // Gaussian elimination with partial pivoting
        for (int c = 0; c < 4; ++c) {
            int piv = c;
            for (int r = c + 1; r < 4; ++r)
                if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
            for (int k = 0; k < 5; ++k) { double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
            for (int r = c + 1; r < 4; ++r) {
                const double f = A[r][c] / A[c][c];
                for (int k = c; k < 5; ++k) A[r][k] -= f * A[c][k];
            }
        }
        double a[4];
        for (int r = 3; r >= 0; --r) {
            double s = A[r][4];
            for (int k = r + 1; k < 4; ++k) s -= A[r][k] * a[k];
            a[r] = s / A[r][r];
        }
        // p(x) = sum a_k ((x - xc)/xh)^k  ->  raw monomials
        const double g1 = a[1] / xh, g2 = a[2] / (xh * xh), g3 = a[3] / (xh * xh * xh);
        sol[0] = a[0] - g1 * xc + g2 * xc * xc - g3 * xc * xc * xc;
        sol[1] = g1 - 2.0 * g2 * xc + 3.0 * g3 * xc * xc;
        sol[2] = g2 - 3.0 * g3 * xc;
        sol[3] = g3;
        for (int k = 0; k < 4; ++k) coef[k] = sol[k];
    }
    __syncthreads();
    if (resid) {
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            resid[i] = horner3(sol, (double)(x0 + i)) - (double)y_resid[i];
    }
}

// keep = |d / std(d)| < nsigma  (population std, as np.std)
__global__ void __launch_bounds__(kFitThreads)
sigma_mask_kernel(const double* __restrict__ d, int n, double nsigma, uint8_t* __restrict__ keep) {
    __shared__ double red[32];
    double s[1] = {0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[0] += d[i];
    block_reduce_sum<1>(s, red);
    const double mean = s[0] / n;
    double q[1] = {0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const double t = d[i] - mean; q[0] += t * t; }
    block_reduce_sum<1>(q, red);
    const double sd = sqrt(q[0] / n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) keep[i] = fabs(d[i] / sd) < nsigma ? 1 : 0;
}

__global__ void window_mask_kernel(const double* __restrict__ d, int n, double centre, double tol,
                                   uint8_t* __restrict__ good) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) good[i] = fabs(d[i] - centre) < tol ? 1 : 0;
}

__global__ void fit_table_kernel(const double* __restrict__ coef, int ih, double* __restrict__ fit) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= ih) return;
    const double p[4] = {coef[0], coef[1], coef[2], coef[3]};
    const double c = horner3(p, (double)y);
    const double f = floor(c);
    fit[4 * y + 0] = f;
    fit[4 * y + 1] = c - f;
    fit[4 * y + 2] = (double)y;
    fit[4 * y + 3] = c;
}

}  // namespace

extern "C" int shg_box_blur_u16(const uint16_t* d_img, int rows, int cols, int kw, int kh, uint16_t* d_out,
                                uint32_t* d_tmp, void* stream) {
    SHG_REQUIRE(kw >= 1 && kh >= 1, "shg_box_blur_u16: kernel size %dx%d (the reference needs y2-y1 >= 100)", kw, kh);
    SHG_REQUIRE(kw <= cols && kh <= rows && (int64_t)kw * kh * 65535 < 0x7fffffff,
                "shg_box_blur_u16: kernel %dx%d does not fit image %dx%d", kw, kh, cols, rows);
    const int64_t n = (int64_t)rows * cols;
    const unsigned blocks = (unsigned)ceil_div64(n, 256);
    cudaStream_t st = as_stream(stream);
    hsum_kernel<<<blocks, 256, 0, st>>>(d_img, rows, cols, kw, d_tmp);
    SHG_LAUNCH_CHECK();
    const double scale = 1.0 / ((double)kw * kh);
    vsum_scale_kernel<<<blocks, 256, 0, st>>>(d_tmp, rows, cols, kh, (float)scale, scale, cols - cols % 8, d_out);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_row_sums_u16(const uint16_t* d_img, int rows, int cols, uint64_t* d_out, void* stream) {
    row_sums_kernel<<<(rows + 7) / 8, 256, 0, as_stream(stream)>>>(
        d_img, rows, cols, reinterpret_cast<unsigned long long*>(d_out));
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_row_argmin_u16(const uint16_t* d_img, int rows, int cols, int c0, int c1, int32_t* d_out,
                                  void* stream) {
    SHG_REQUIRE(0 <= c0 && c0 < c1 && c1 <= cols && cols <= 65536, "shg_row_argmin_u16: bad column window [%d,%d) of %d", c0, c1, cols);
    row_argmin_kernel<<<(rows + 7) / 8, 256, 0, as_stream(stream)>>>(d_img, rows, cols, c0, c1, d_out);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_polyfit3(const int32_t* d_y, const uint8_t* d_mask, int x0, int n, double* d_coef,
                            const int32_t* d_y_resid, double* d_resid, void* stream) {
    SHG_REQUIRE(n >= 4, "shg_polyfit3: need at least 4 points, got %d", n);
    polyfit3_kernel<<<1, kFitThreads, 0, as_stream(stream)>>>(d_y, d_mask, x0, n, d_coef,
                                                            d_y_resid ? d_y_resid : d_y, d_resid);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_sigma_mask(const double* d_resid, int n, double nsigma, uint8_t* d_keep, void* stream) {
    sigma_mask_kernel<<<1, kFitThreads, 0, as_stream(stream)>>>(d_resid, n, nsigma, d_keep);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_window_mask(const double* d_resid, int n, double centre, double tol, uint8_t* d_good, void* stream) {
    window_mask_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(d_resid, n, centre, tol, d_good);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_fit_table(const double* d_coef, int ih, double* d_fit, void* stream) {
    fit_table_kernel<<<(ih + 255) / 256, 256, 0, as_stream(stream)>>>(d_coef, ih, d_fit);
    SHG_LAUNCH_CHECK();
    return 0;
}
