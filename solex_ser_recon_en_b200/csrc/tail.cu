// image_process tail on the device (reference solex_util.py:527-546): CLAHE on uint16 with
// OpenCV's algorithm (cv2.createCLAHE(clipLimit, tileGridSize).apply, modules/imgproc/src/clahe.cpp) and the
// percentile / brightness rescales.  Each step keeps OpenCV's / NumPy's operation order so the images agree bit for
// bit with the host tail:
//   tile histograms (65536 bins; the image is extended by BORDER_REFLECT_101 to a multiple of the tile grid),
//   clip at clipLimit*tileArea/65536 (>= 1), redistribute the excess (batch + strided residual),
//   lut[i] = saturate_cast<ushort>(cvRound(cumsum[i] * (65535 / tileArea)))      (float),
//   res = (lut11*xa1 + lut12*xa)*ya1 + (lut21*xa1 + lut22*xa)*ya                   (float, no fma), cvRound,
//   rescale_brightness: trunc(clip(65535.0 * (p - lo) / (hi - lo), 0, 65535))      (double).
// PNG / FITS encoding, the protuberance disc (cv2.circle) and rotations stay on the host.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kBins = 65536;

// Size of the image OpenCV cuts tiles from: unchanged when both sides divide by the tile grid; otherwise
// copyMakeBorder(src, 0, tilesY - rows % tilesY, 0, tilesX - cols % tilesX, BORDER_REFLECT_101) -- note that a side
// that DOES divide then still grows by a whole tile count (clahe.cpp: the two paddings are computed independently
// of which side needed one).
inline void clahe_extent(int rows, int cols, int tiles_x, int tiles_y, int* ext_rows, int* ext_cols) {
    if (cols % tiles_x == 0 && rows % tiles_y == 0) {
        *ext_rows = rows;
        *ext_cols = cols;
    } else {
        *ext_rows = rows + (tiles_y - rows % tiles_y);
        *ext_cols = cols + (tiles_x - cols % tiles_x);
    }
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// histogram of every tile of the extended image (+ optionally of the real image)
__global__ void __launch_bounds__(256)
tile_hist_kernel(const uint16_t* __restrict__ img, int rows, int cols, int ext_rows, int ext_cols, int tile_w, int tile_h,
                 int tiles_x, unsigned int* __restrict__ tile_hist, unsigned int* __restrict__ full_hist) {
    const int64_t n = (int64_t)ext_rows * ext_cols;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int y = (int)(i / ext_cols), x = (int)(i % ext_cols);
        const uint32_t v = img[(int64_t)reflect101(y, rows) * cols + reflect101(x, cols)];
        const int t = (y / tile_h) * tiles_x + (x / tile_w);
        atomicAdd(&tile_hist[(int64_t)t * kBins + v], 1u);
        if (full_hist && y < rows && x < cols) atomicAdd(&full_hist[v], 1u);
    }
}

// one CTA per tile: clip, redistribute, cumulative sum, scale
__global__ void __launch_bounds__(1024)
clahe_lut_kernel(unsigned int* __restrict__ tile_hist, int clip_limit, float lut_scale, uint16_t* __restrict__ lut) {
    unsigned int* h = tile_hist + (int64_t)blockIdx.x * kBins;
    uint16_t* out = lut + (int64_t)blockIdx.x * kBins;
    __shared__ unsigned long long s_red[32];
    __shared__ unsigned long long s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int PER = kBins / 1024;                               // 64 consecutive bins per thread
    if (clip_limit > 0) {
        unsigned long long clipped = 0;
        for (int i = tid; i < kBins; i += 1024) {
            const unsigned int v = h[i];
            if (v > (unsigned)clip_limit) { clipped += v - clip_limit; h[i] = clip_limit; }
        }
        for (int o = 16; o; o >>= 1) clipped += __shfl_xor_sync(0xffffffffu, clipped, o);
        if (lane == 0) s_red[warp] = clipped;
        __syncthreads();
        if (tid == 0) {
            unsigned long long t = 0;
            for (int w = 0; w < 32; ++w) t += s_red[w];
            s_total = t;
        }
        __syncthreads();
        const long long total = (long long)s_total;
        const int batch = (int)(total / kBins);
        int residual = (int)(total - (long long)batch * kBins);
        const int step = residual ? max(kBins / residual, 1) : 1;
        for (int i = tid; i < kBins; i += 1024) {
            unsigned int v = h[i] + batch;
            // for (i = 0; i < histSize && residual > 0; i += step, residual--) hist[i]++
            if (residual && i % step == 0 && i / step < residual) ++v;
            h[i] = v;
        }
        __syncthreads();
    }
    // inclusive scan: PER consecutive bins per thread, then across threads
    unsigned int local[PER];
    unsigned long long sum = 0;
#pragma unroll 8
    for (int q = 0; q < PER; ++q) { local[q] = h[tid * PER + q]; sum += local[q]; }
    unsigned long long inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    __syncthreads();
    if (lane == 31) s_red[warp] = inc;
    __syncthreads();
    unsigned long long base = inc - sum;
    for (int w = 0; w < warp; ++w) base += s_red[w];
    unsigned long long run = base;
#pragma unroll 8
    for (int q = 0; q < PER; ++q) {
        run += local[q];
        const float v = __fmul_rn((float)(int)run, lut_scale);      // the reference sums in int and converts
        const int r = __float2int_rn(v);
        out[tid * PER + q] = (uint16_t)min(max(r, 0), 65535);
    }
}

__global__ void __launch_bounds__(256)
clahe_apply_kernel(const uint16_t* __restrict__ img, int rows, int cols, int tiles_x, int tiles_y, float inv_tw,
                   float inv_th, const uint16_t* __restrict__ lut, uint16_t* __restrict__ out,
                   unsigned int* __restrict__ out_hist) {
    const int y = blockIdx.y;
    const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
    int ty1 = (int)floorf(tyf);
    int ty2 = ty1 + 1;
    const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
    ty1 = max(ty1, 0);
    ty2 = min(ty2, tiles_y - 1);
    const uint16_t* p1 = lut + (int64_t)ty1 * tiles_x * kBins;
    const uint16_t* p2 = lut + (int64_t)ty2 * tiles_x * kBins;
    for (int x = blockIdx.x * 256 + threadIdx.x; x < cols; x += gridDim.x * 256) {
        const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
        int tx1 = (int)floorf(txf);
        int tx2 = tx1 + 1;
        const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
        tx1 = max(tx1, 0);
        tx2 = min(tx2, tiles_x - 1);
        const uint32_t v = img[(int64_t)y * cols + x];
        const float l11 = (float)p1[tx1 * kBins + v], l12 = (float)p1[tx2 * kBins + v];
        const float l21 = (float)p2[tx1 * kBins + v], l22 = (float)p2[tx2 * kBins + v];
        const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
        const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
        const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
        const int r = min(max(__float2int_rn(res), 0), 65535);
        out[(int64_t)y * cols + x] = (uint16_t)r;
        if (out_hist) atomicAdd(&out_hist[r], 1u);
    }
}

__global__ void __launch_bounds__(256)
rescale_kernel(const uint16_t* __restrict__ img, int64_t n, double lo, double hi_minus_lo, uint16_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        // float(sat) * alpha * (img - lo) / (hi - lo), evaluated left to right in double
        double v = __ddiv_rn(__dmul_rn(65535.0, __dsub_rn((double)img[i], lo)), hi_minus_lo);
        v = v < 0.0 ? 0.0 : v;
        v = v > 65535.0 ? 65535.0 : v;
        out[i] = (uint16_t)(unsigned int)v;                          // astype: truncation
    }
}

}  // namespace

extern "C" int shg_tile_hist_u16(const uint16_t* d_img, int rows, int cols, int tiles_x, int tiles_y,
                                 uint32_t* d_tile_hist, uint32_t* d_full_hist, void* stream) {
    SHG_REQUIRE(d_img && d_tile_hist && rows > 1 && cols > 1 && tiles_x >= 1 && tiles_y >= 1 && tiles_x * tiles_y <= 64,
                "shg_tile_hist_u16: bad arguments");
    int ext_rows, ext_cols;
    clahe_extent(rows, cols, tiles_x, tiles_y, &ext_rows, &ext_cols);
    cudaStream_t st = as_stream(stream);
    SHG_CHECK(cudaMemsetAsync(d_tile_hist, 0, (size_t)tiles_x * tiles_y * kBins * 4, st));
    if (d_full_hist) SHG_CHECK(cudaMemsetAsync(d_full_hist, 0, (size_t)kBins * 4, st));
    const int64_t n = (int64_t)ext_rows * ext_cols;
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n, 256 * 4), SHG_SM_COUNT_B200 * 16));
    tile_hist_kernel<<<blocks, 256, 0, st>>>(d_img, rows, cols, ext_rows, ext_cols, ext_cols / tiles_x, ext_rows / tiles_y,
                                             tiles_x, d_tile_hist, d_full_hist);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t shg_clahe_tile_area(int rows, int cols, int tiles_x, int tiles_y) {
    int ext_rows, ext_cols;
    clahe_extent(rows, cols, tiles_x, tiles_y, &ext_rows, &ext_cols);
    return (int64_t)(ext_cols / tiles_x) * (ext_rows / tiles_y);
}

extern "C" int shg_clahe_lut(uint32_t* d_tile_hist, int n_tiles, int64_t tile_area, double clip_limit, uint16_t* d_lut,
                             void* stream) {
    SHG_REQUIRE(d_tile_hist && d_lut && n_tiles >= 1 && tile_area >= 1, "shg_clahe_lut: bad arguments");
    SHG_REQUIRE(tile_area < (1LL << 31), "shg_clahe_lut: tile too large");
    int clip = 0;
    if (clip_limit > 0.0) clip = std::max((int)(clip_limit * (double)tile_area / kBins), 1);
    const float lut_scale = (float)(kBins - 1) / (float)tile_area;
    clahe_lut_kernel<<<n_tiles, 1024, 0, as_stream(stream)>>>(d_tile_hist, clip, lut_scale, d_lut);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_clahe_apply(const uint16_t* d_img, int rows, int cols, int tiles_x, int tiles_y, const uint16_t* d_lut,
                               uint16_t* d_out, uint32_t* d_out_hist, void* stream) {
    SHG_REQUIRE(d_img && d_lut && d_out && rows > 1 && cols > 1 && rows <= 65535, "shg_clahe_apply: bad arguments");
    int ext_rows, ext_cols;
    clahe_extent(rows, cols, tiles_x, tiles_y, &ext_rows, &ext_cols);
    const float inv_tw = 1.0f / (float)(ext_cols / tiles_x), inv_th = 1.0f / (float)(ext_rows / tiles_y);
    cudaStream_t st = as_stream(stream);
    if (d_out_hist) SHG_CHECK(cudaMemsetAsync(d_out_hist, 0, (size_t)kBins * 4, st));
    dim3 grid((unsigned)std::max(1, std::min(8, (cols + 1023) / 1024)), rows);
    clahe_apply_kernel<<<grid, 256, 0, st>>>(d_img, rows, cols, tiles_x, tiles_y, inv_tw, inv_th, d_lut, d_out, d_out_hist);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_rescale_u16(const uint16_t* d_img, int64_t n, double lo, double hi, uint16_t* d_out, void* stream) {
    SHG_REQUIRE(d_img && d_out && n >= 0 && hi > lo, "shg_rescale_u16: needs hi > lo");
    if (n == 0) return 0;
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n, 256 * 4), SHG_SM_COUNT_B200 * 16));
    rescale_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_img, n, lo, hi - lo, d_out);
    SHG_LAUNCH_CHECK();
    return 0;
}
