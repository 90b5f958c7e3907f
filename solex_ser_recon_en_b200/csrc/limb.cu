// Limb detection front end of the ellipse fit, on the 4x-downscaled disk image
// (reference ellipse_to_circle.py:148-228 get_flood_image, :231-250 the Canny
// call of get_edge_list).  The image is kept as the integer 4x4 block sums S
// (value = S * 2^-20 of the reference's image/65536), which makes OpenCV's
// double-precision box filter exact in integers; every floating-point step
// below reproduces the operation ORDER of the library routine the reference
// calls, so thresholds and tie-breaks come out the same:
//   cv2.blur (CV_64F)            exact window sum, one multiply by 1/(kw*kh)
//   np.percentile / np.median    exact order statistics by radix select on the integer sums
//   np.histogram (20 uniform bins)  bin = the edge interval that contains the value
//   scipy.ndimage.gaussian_filter   NI_Correlate1D symmetric form: in[0]*w[0] + sum_{j=-r..-1} (in[j]+in[-j])*w[j]
//   scipy.ndimage.sobel          (in[-1]-in[+1])*(-1) then in[0]*2 + (in[-1]+in[+1])*1, mode 'reflect'
//   skimage.feature.canny        magnitude, interpolated non-maximum suppression (_canny_cy)
// Outputs are tiny (order statistics, 20 counts, a list of thin-edge pixels);
// labelling, convex hull and the 6-parameter ellipse fit stay on the host.
// All kernels are O(rows*cols) over a ~5 Mpx image: microseconds each.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}
__device__ __forceinline__ int reflect_edge(int i, int n) {      // scipy 'reflect': d c b a | a b c d | d c b a
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - 1 - i;
    return i;
}

__global__ void __launch_bounds__(256)
hsum_u32_kernel(const uint32_t* __restrict__ img, int rows, int cols, int kw, uint32_t* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)rows * cols) return;
    const int r = (int)(idx / cols), c = (int)(idx % cols);
    const uint32_t* row = img + (int64_t)r * cols;
    const int a = c - kw / 2;
    uint32_t s = 0;
    for (int t = 0; t < kw; ++t) s += row[reflect101(a + t, cols)];
    out[idx] = s;
}

__global__ void __launch_bounds__(256)
vsum_u32_kernel(const uint32_t* __restrict__ hs, int rows, int cols, int kh, uint32_t* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)rows * cols) return;
    const int r = (int)(idx / cols), c = (int)(idx % cols);
    const int a = r - kh / 2;
    uint32_t s = 0;
    for (int t = 0; t < kh; ++t) s += hs[(int64_t)reflect101(a + t, rows) * cols + c];
    out[idx] = s;
}

__global__ void __launch_bounds__(256)
sum_u32_kernel(const uint32_t* __restrict__ v, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) s += v[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// histogram of byte `shift/8` of the keys whose bits above it equal `prefix`
__global__ void __launch_bounds__(256)
radix_hist_kernel(const uint32_t* __restrict__ v, int64_t n, uint32_t prefix, int shift, unsigned int* __restrict__ hist) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int hi_shift = shift + 8;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const uint32_t k = v[i];
        if (hi_shift >= 32 || (k >> hi_shift) == prefix) atomicAdd(&h[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

// blurred(B) = (B * 2^-20) * scale, exactly as cv2.blur computes it from image/65536
__device__ __forceinline__ double blurred_of(uint32_t b, double scale) {
    return __dmul_rn(__dmul_rn((double)b, 9.5367431640625e-07), scale);
}

// out[0] = min B, out[1] = max B among pixels whose blurred value is < ceiling
__global__ void __launch_bounds__(256)
blur_range_kernel(const uint32_t* __restrict__ box, int64_t n, double scale, double ceiling, unsigned int* __restrict__ out) {
    unsigned int lo = 0xffffffffu, hi = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const uint32_t b = box[i];
        if (blurred_of(b, scale) < ceiling) { lo = min(lo, b); hi = max(hi, b); }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(out, lo); atomicMax(out + 1, hi); }
}

struct Edges { double e[33]; int n_bins; };

__global__ void __launch_bounds__(256)
blur_hist_kernel(const uint32_t* __restrict__ box, int64_t n, double scale, double ceiling, const Edges ed,
                 unsigned long long* __restrict__ counts) {
    __shared__ unsigned int h[32];
    if (threadIdx.x < 32) h[threadIdx.x] = 0;
    __syncthreads();
    const double first = ed.e[0], last = ed.e[ed.n_bins];
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double x = blurred_of(box[i], scale);
        if (!(x < ceiling) || x < first || x > last) continue;
        int b = 0;                                              // edges[b] <= x < edges[b+1]; last bin closed
        for (int j = 1; j < ed.n_bins; ++j) b += x >= ed.e[j] ? 1 : 0;
        atomicAdd(&h[b], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32 && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

struct GaussW { double w[33]; int radius; };                   // w[k] = weight at offset -k (symmetric)

// gaussian along `axis` of either the flood image made on the fly from the box
// sums (src == nullptr) or a double image; mode 'constant' (zeros outside)
template <bool FROM_BOX>
__global__ void __launch_bounds__(256)
gauss1d_kernel(const uint32_t* __restrict__ box, double scale, double level, int ones,
               const double* __restrict__ src, int rows, int cols, int axis, const GaussW g, double* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)rows * cols) return;
    const int r = (int)(idx / cols), c = (int)(idx % cols);
    auto at = [&](int rr, int cc) -> double {
        if (rr < 0 || rr >= rows || cc < 0 || cc >= cols) return 0.0;
        if (FROM_BOX) {
            if (ones) return 1.0;
            return blurred_of(box[(int64_t)rr * cols + cc], scale) < level ? 0.0 : 65000.0;
        }
        return src[(int64_t)rr * cols + cc];
    };
    const int dr = axis == 0 ? 1 : 0, dc = axis == 0 ? 0 : 1;
    double acc = __dmul_rn(at(r, c), g.w[0]);
    for (int k = g.radius; k >= 1; --k) {                       // j = -radius .. -1
        const double pair = __dadd_rn(at(r - k * dr, c - k * dc), at(r + k * dr, c + k * dc));
        acc = __dadd_rn(acc, __dmul_rn(pair, g.w[k]));
    }
    out[idx] = acc;
}

__global__ void __launch_bounds__(256)
divide_kernel(double* __restrict__ num, const double* __restrict__ den, double eps, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) num[i] = __ddiv_rn(num[i], __dadd_rn(den[i], eps));
}

// scipy.ndimage.sobel along both axes + magnitude
__global__ void __launch_bounds__(256)
sobel_kernel(const double* __restrict__ s, int rows, int cols, double* __restrict__ gi, double* __restrict__ gj,
             double* __restrict__ mag) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)rows * cols) return;
    const int r = (int)(idx / cols), c = (int)(idx % cols);
    auto S = [&](int rr, int cc) { return s[(int64_t)reflect_edge(rr, rows) * cols + reflect_edge(cc, cols)]; };
    // derivative: in[0]*0 + (in[-1] - in[+1]) * (-1)
    auto d1 = [&](int rr, int cc) {                              // along axis 1 (columns)
        return __dadd_rn(__dmul_rn(S(rr, cc), 0.0), __dmul_rn(__dsub_rn(S(rr, cc - 1), S(rr, cc + 1)), -1.0));
    };
    auto d0 = [&](int rr, int cc) {                              // along axis 0 (rows)
        return __dadd_rn(__dmul_rn(S(rr, cc), 0.0), __dmul_rn(__dsub_rn(S(rr - 1, cc), S(rr + 1, cc)), -1.0));
    };
    // the derivative image is itself extended by reflection before the [1,2,1] pass
    auto D1 = [&](int rr, int cc) { return d1(reflect_edge(rr, rows), reflect_edge(cc, cols)); };
    auto D0 = [&](int rr, int cc) { return d0(reflect_edge(rr, rows), reflect_edge(cc, cols)); };
    // smoothing: in[0]*2 + (in[-1] + in[+1]) * 1
    const double j = __dadd_rn(__dmul_rn(D1(r, c), 2.0), __dmul_rn(__dadd_rn(D1(r - 1, c), D1(r + 1, c)), 1.0));
    const double i = __dadd_rn(__dmul_rn(D0(r, c), 2.0), __dmul_rn(__dadd_rn(D0(r, c - 1), D0(r, c + 1)), 1.0));
    gi[idx] = i;
    gj[idx] = j;
    mag[idx] = __dsqrt_rn(__dadd_rn(__dmul_rn(i, i), __dmul_rn(j, j)));
}

// interpolated non-maximum suppression; survivors are appended to a list
__global__ void __launch_bounds__(256)
nms_kernel(const double* __restrict__ gi, const double* __restrict__ gj, const double* __restrict__ mag, int rows,
           int cols, double low, unsigned int* __restrict__ count, unsigned int cap, uint32_t* __restrict__ list_idx,
           double* __restrict__ list_mag) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)rows * cols) return;
    const int r = (int)(idx / cols), c = (int)(idx % cols);
    if (r < 1 || r >= rows - 1 || c < 1 || c >= cols - 1) return;          // eroded mask
    const double m = mag[idx];
    if (!(m >= low)) return;
    const double a = gi[idx], b = gj[idx];
    const double aa = fabs(a), ab = fabs(b);
    const bool same = (a >= 0 && b >= 0) || (a <= 0 && b <= 0);
    const bool opp = (a <= 0 && b >= 0) || (a >= 0 && b <= 0);
    if (!(same || opp)) return;                                             // nan gradients
    int d1i, d1j, d2i, d2j;
    double w;
    if (same) {
        if (aa > ab) { d1i = 1; d1j = 0; d2i = 1; d2j = 1; w = __ddiv_rn(ab, aa); }
        else         { d1i = 0; d1j = 1; d2i = 1; d2j = 1; w = __ddiv_rn(aa, ab); }
    } else {
        if (aa < ab) { d1i = 0; d1j = 1; d2i = -1; d2j = 1; w = __ddiv_rn(aa, ab); }
        else         { d1i = -1; d1j = 0; d2i = -1; d2j = 1; w = __ddiv_rn(ab, aa); }
    }
    auto M = [&](int di, int dj) { return mag[(int64_t)(r + di) * cols + (c + dj)]; };
    const double omw = __dsub_rn(1.0, w);
    const double plus = __dadd_rn(__dmul_rn(M(d2i, d2j), w), __dmul_rn(M(d1i, d1j), omw));
    const double minus = __dadd_rn(__dmul_rn(M(-d2i, -d2j), w), __dmul_rn(M(-d1i, -d1j), omw));
    if (plus <= m && minus <= m && m > 0.0) {
        const unsigned int slot = atomicAdd(count, 1u);
        if (slot < cap) { list_idx[slot] = (uint32_t)idx; list_mag[slot] = m; }
    }
}

// ---------------------------------------------------------------------------
// Chained front end (shg_limb_front): every data-dependent scalar of the flood
// threshold search -- the four order statistics, the percentile, the histogram
// range, its 21 edges -- stays in a small device-resident block, so the ~20
// kernels queue back to back and the host reads ONE block back at the end
// (the step-by-step entry points above cost ~15 blocking round trips).
struct LimbState {
    unsigned long long total;          // sum of the block sums
    unsigned long long counts[32];     // histogram of the blurred image below the ceiling
    double ceiling;
    double edges[33];
    long long rank[4];                 // remaining rank of each order statistic inside its prefix group
    uint32_t prefix[4];                // value bits decided so far (queries 0,1: box; 2,3: box5)
    uint32_t range[2];                 // min / max box sum below the ceiling
    uint32_t hist[4][4][256];          // [pass][query][byte]
};

// byte histograms of one radix pass for both arrays (blockIdx.y = array); the two queries of an
// array share a histogram while their prefixes agree
__global__ void __launch_bounds__(256)
select_pass_kernel(const uint32_t* __restrict__ v0, const uint32_t* __restrict__ v1, int64_t n, int pass,
                   LimbState* __restrict__ st) {
    __shared__ unsigned int h[2][256];
    const int arr = blockIdx.y;
    const uint32_t* __restrict__ v = arr ? v1 : v0;
    h[0][threadIdx.x] = 0;
    h[1][threadIdx.x] = 0;
    __syncthreads();
    const int shift = 24 - 8 * pass, hi_shift = shift + 8;
    const uint32_t p0 = st->prefix[2 * arr], p1 = st->prefix[2 * arr + 1];
    const bool split = p0 != p1;
    // the blurred background is nearly constant: most keys of a warp fall into ONE bin, so the lanes that share
    // a bin elect a leader that adds their count once (a plain shared atomic would serialise 32 ways)
    const int64_t n_round = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_round; i += (int64_t)gridDim.x * 256) {
        const bool live = i < n;
        const uint32_t k = live ? v[i] : 0u;
        const uint32_t top = hi_shift >= 32 ? 0u : (k >> hi_shift);
        int slot = -1;                                            // 0..255: histogram 0, 256..511: histogram 1
        if (live && top == p0) slot = (int)((k >> shift) & 255u);
        else if (live && split && top == p1) slot = 256 + (int)((k >> shift) & 255u);
        const unsigned peers = __match_any_sync(0xffffffffu, slot);
        if (slot >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&h[0][0] + slot, (unsigned)__popc(peers));
    }
    __syncthreads();
    if (h[0][threadIdx.x]) atomicAdd(&st->hist[pass][2 * arr][threadIdx.x], h[0][threadIdx.x]);
    if (split && h[1][threadIdx.x]) atomicAdd(&st->hist[pass][2 * arr + 1][threadIdx.x], h[1][threadIdx.x]);
}

__global__ void select_step_kernel(int pass, LimbState* __restrict__ st) {
    const int q = threadIdx.x;
    if (q >= 4) return;
    const uint32_t mine = st->prefix[q], mate = st->prefix[q & ~1];
    const unsigned int* h = st->hist[pass][(q & 1) && mine == mate ? q - 1 : q];
    long long r = st->rank[q];
    int b = 0;
    for (; b < 255; ++b) {
        if (r < (long long)h[b]) break;
        r -= h[b];
    }
    __syncwarp(0xf);                                   // every query has read the shared prefixes
    st->rank[q] = r;
    st->prefix[q] = (mine << 8) | (uint32_t)b;
}

// np.percentile(blurred, 99) from its two bracketing order statistics (method 'linear', numpy's _lerp)
__global__ void limb_ceiling_kernel(double scale, double gamma, double one_minus_gamma, LimbState* __restrict__ st) {
    const double a = blurred_of(st->prefix[0], scale), b = blurred_of(st->prefix[1], scale);
    const double diff = __dsub_rn(b, a);
    double out = __dadd_rn(a, __dmul_rn(diff, gamma));
    if (gamma >= 0.5) out = __dsub_rn(b, __dmul_rn(diff, one_minus_gamma));
    st->ceiling = out;
    st->range[0] = 0xffffffffu;
    st->range[1] = 0u;
}

__global__ void __launch_bounds__(256)
blur_range_dev_kernel(const uint32_t* __restrict__ box, int64_t n, double scale, LimbState* __restrict__ st) {
    const double ceiling = st->ceiling;
    unsigned int lo = 0xffffffffu, hi = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const uint32_t b = box[i];
        if (blurred_of(b, scale) < ceiling) { lo = min(lo, b); hi = max(hi, b); }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(&st->range[0], lo); atomicMax(&st->range[1], hi); }
}

// np.histogram's bin edges: np.linspace(first, last, n_bins + 1) = arange * step + first, end point forced
__global__ void limb_edges_kernel(double scale, int n_bins, LimbState* __restrict__ st) {
    const int i = threadIdx.x;
    if (i > n_bins) return;
    double first, last;
    if (st->range[0] > st->range[1]) { first = 0.0; last = 1.0; }             // nothing below the ceiling
    else { first = blurred_of(st->range[0], scale); last = blurred_of(st->range[1], scale); }
    if (first == last) { first = __dsub_rn(first, 0.5); last = __dadd_rn(last, 0.5); }
    const double delta = __dsub_rn(last, first);
    const double step = __ddiv_rn(delta, (double)n_bins);
    double y;
    if (step == 0.0) y = __dmul_rn(__ddiv_rn((double)i, (double)n_bins), delta);
    else y = __dmul_rn((double)i, step);
    y = __dadd_rn(y, first);
    if (i == n_bins) y = last;
    st->edges[i] = y;
}

__global__ void __launch_bounds__(256)
blur_hist_dev_kernel(const uint32_t* __restrict__ box, int64_t n, double scale, int n_bins, LimbState* __restrict__ st) {
    __shared__ unsigned int h[32];
    __shared__ double e[33];
    if (threadIdx.x < 32) h[threadIdx.x] = 0;
    if (threadIdx.x <= n_bins) e[threadIdx.x] = st->edges[threadIdx.x];
    __syncthreads();
    const double ceiling = st->ceiling, first = e[0], last = e[n_bins];
    const int64_t n_round = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_round; i += (int64_t)gridDim.x * 256) {
        int b = -1;
        if (i < n) {
            const double x = blurred_of(box[i], scale);
            if (x < ceiling && !(x < first) && !(x > last)) {
                b = 0;
                for (int j = 1; j < n_bins; ++j) b += x >= e[j] ? 1 : 0;
            }
        }
        const unsigned peers = __match_any_sync(0xffffffffu, b);      // lanes of one bin add once (see the select pass)
        if (b >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&h[b], (unsigned)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < 32 && h[threadIdx.x]) atomicAdd(&st->counts[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// ---------------------------------------------------------------------------
// Fused canny front (shg_limb_canny): two kernels instead of seven, same arithmetic in the same order.
//  K1 flood_gauss_kernel: flood image made on the fly from the box sums, gaussian along axis 0 then axis 1 through
//     shared memory, the same for the all-ones mask, and the division -- one pass over the image instead of five
//     (four 17-tap passes over 40 MB images of doubles + the divide).
//  K2 sobel_nms_kernel: Sobel at the pixel; only where the magnitude reaches `low` (a few per cent of the image:
//     the flood image is flat away from the limb) the four neighbour magnitudes the interpolated non-maximum
//     suppression needs are evaluated -- no gi / gj / magnitude images are written at all.
constexpr int kGT_H = 16, kGT_W = 64;

__global__ void __launch_bounds__(256)
flood_gauss_kernel(const uint32_t* __restrict__ box, int rows, int cols, double scale, double level, const GaussW g,
                   double eps, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int R = g.radius;
    const int fw = kGT_W + 2 * R, fh = kGT_H + 2 * R;
    double* a = reinterpret_cast<double*>(sm_raw);                 // [kGT_H][fw]   axis-0 result of the flood image
    double* a1row = a + kGT_H * fw;                                // [kGT_H]       axis-0 result of the ones mask
    unsigned char* f = reinterpret_cast<unsigned char*>(a1row + kGT_H);   // [fh][fw]  bit0: flood on, bit1: inside
    const int r0 = blockIdx.y * kGT_H, c0 = blockIdx.x * kGT_W;
    for (int i = threadIdx.x; i < fh * fw; i += 256) {
        const int rr = i / fw, cc = i - rr * fw;
        const int gr = r0 - R + rr, gc = c0 - R + cc;
        unsigned char v = 0;
        if (gr >= 0 && gr < rows && gc >= 0 && gc < cols)
            v = 2 | (blurred_of(box[(int64_t)gr * cols + gc], scale) < level ? 0 : 1);
        f[i] = v;
    }
    __syncthreads();
    // axis 0 (rows): in[0]*w[0] + sum_{k = R..1} (in[-k] + in[+k]) * w[k]
    for (int i = threadIdx.x; i < kGT_H * fw; i += 256) {
        const int rr = i / fw, cc = i - rr * fw;
        const unsigned char* col = f + (rr + R) * fw + cc;
        double acc = __dmul_rn((col[0] & 1) ? 65000.0 : 0.0, g.w[0]);
        for (int k = R; k >= 1; --k) {
            const double pair = __dadd_rn((col[-k * fw] & 1) ? 65000.0 : 0.0, (col[k * fw] & 1) ? 65000.0 : 0.0);
            acc = __dadd_rn(acc, __dmul_rn(pair, g.w[k]));
        }
        a[i] = acc;
    }
    if (threadIdx.x < kGT_H) {
        const int gr = r0 + threadIdx.x;
        double acc = __dmul_rn((gr < rows) ? 1.0 : 0.0, g.w[0]);
        for (int k = R; k >= 1; --k) {
            const double pair = __dadd_rn((gr - k >= 0 && gr - k < rows) ? 1.0 : 0.0, (gr + k < rows && gr + k >= 0) ? 1.0 : 0.0);
            acc = __dadd_rn(acc, __dmul_rn(pair, g.w[k]));
        }
        a1row[threadIdx.x] = acc;
    }
    __syncthreads();
    // axis 1 (columns) of both, then the division
    for (int i = threadIdx.x; i < kGT_H * kGT_W; i += 256) {
        const int rr = i / kGT_W, cc = i - rr * kGT_W;
        const int gr = r0 + rr, gc = c0 + cc;
        if (gr >= rows || gc >= cols) continue;
        const double* arow = a + rr * fw + cc + R;
        const double one = a1row[rr];
        double b = __dmul_rn(arow[0], g.w[0]);
        double b1 = __dmul_rn(one, g.w[0]);
        for (int k = R; k >= 1; --k) {
            b = __dadd_rn(b, __dmul_rn(__dadd_rn(arow[-k], arow[k]), g.w[k]));
            const double pair1 = __dadd_rn(gc - k >= 0 ? one : 0.0, gc + k < cols ? one : 0.0);
            b1 = __dadd_rn(b1, __dmul_rn(pair1, g.w[k]));
        }
        out[(int64_t)gr * cols + gc] = __ddiv_rn(b, __dadd_rn(b1, eps));
    }
}

struct SobelOut { double gi, gj, mag; };

__device__ __forceinline__ SobelOut sobel_at(const double* __restrict__ s, int rows, int cols, int r, int c) {
    auto S = [&](int rr, int cc) { return s[(int64_t)reflect_edge(rr, rows) * cols + reflect_edge(cc, cols)]; };
    auto d1 = [&](int rr, int cc) {
        return __dadd_rn(__dmul_rn(S(rr, cc), 0.0), __dmul_rn(__dsub_rn(S(rr, cc - 1), S(rr, cc + 1)), -1.0));
    };
    auto d0 = [&](int rr, int cc) {
        return __dadd_rn(__dmul_rn(S(rr, cc), 0.0), __dmul_rn(__dsub_rn(S(rr - 1, cc), S(rr + 1, cc)), -1.0));
    };
    auto D1 = [&](int rr, int cc) { return d1(reflect_edge(rr, rows), reflect_edge(cc, cols)); };
    auto D0 = [&](int rr, int cc) { return d0(reflect_edge(rr, rows), reflect_edge(cc, cols)); };
    SobelOut o;
    o.gj = __dadd_rn(__dmul_rn(D1(r, c), 2.0), __dmul_rn(__dadd_rn(D1(r - 1, c), D1(r + 1, c)), 1.0));
    o.gi = __dadd_rn(__dmul_rn(D0(r, c), 2.0), __dmul_rn(__dadd_rn(D0(r, c - 1), D0(r, c + 1)), 1.0));
    o.mag = __dsqrt_rn(__dadd_rn(__dmul_rn(o.gi, o.gi), __dmul_rn(o.gj, o.gj)));
    return o;
}

__global__ void __launch_bounds__(256)
sobel_nms_kernel(const double* __restrict__ s, int rows, int cols, double low, unsigned int* __restrict__ count,
                 unsigned int cap, uint32_t* __restrict__ list_idx, double* __restrict__ list_mag) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)rows * cols) return;
    const int r = (int)(idx / cols), c = (int)(idx % cols);
    if (r < 1 || r >= rows - 1 || c < 1 || c >= cols - 1) return;          // eroded mask
    const SobelOut me = sobel_at(s, rows, cols, r, c);
    const double m = me.mag;
    if (!(m >= low)) return;
    const double a = me.gi, b = me.gj;
    const double aa = fabs(a), ab = fabs(b);
    const bool same = (a >= 0 && b >= 0) || (a <= 0 && b <= 0);
    const bool opp = (a <= 0 && b >= 0) || (a >= 0 && b <= 0);
    if (!(same || opp)) return;                                             // nan gradients
    int d1i, d1j, d2i, d2j;
    double w;
    if (same) {
        if (aa > ab) { d1i = 1; d1j = 0; d2i = 1; d2j = 1; w = __ddiv_rn(ab, aa); }
        else         { d1i = 0; d1j = 1; d2i = 1; d2j = 1; w = __ddiv_rn(aa, ab); }
    } else {
        if (aa < ab) { d1i = 0; d1j = 1; d2i = -1; d2j = 1; w = __ddiv_rn(aa, ab); }
        else         { d1i = -1; d1j = 0; d2i = -1; d2j = 1; w = __ddiv_rn(ab, aa); }
    }
    auto M = [&](int di, int dj) { return sobel_at(s, rows, cols, r + di, c + dj).mag; };
    const double omw = __dsub_rn(1.0, w);
    const double plus = __dadd_rn(__dmul_rn(M(d2i, d2j), w), __dmul_rn(M(d1i, d1j), omw));
    const double minus = __dadd_rn(__dmul_rn(M(-d2i, -d2j), w), __dmul_rn(M(-d1i, -d1j), omw));
    if (plus <= m && minus <= m && m > 0.0) {
        const unsigned int slot = atomicAdd(count, 1u);
        if (slot < cap) { list_idx[slot] = (uint32_t)idx; list_mag[slot] = m; }
    }
}

unsigned grid_for(int64_t n) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n, 256), 148 * 16)); }

}  // namespace

extern "C" int shg_box_sum_u32(const uint32_t* d_in, int rows, int cols, int kw, int kh, uint32_t* d_out,
                               uint32_t* d_tmp, void* stream) {
    SHG_REQUIRE(kw >= 1 && kh >= 1 && kw <= cols && kh <= rows, "shg_box_sum_u32: kernel %dx%d does not fit %dx%d", kw,
                kh, cols, rows);
    SHG_REQUIRE((int64_t)kw * kh * (16LL * 65535) < 0xffffffffLL, "shg_box_sum_u32: window too large for 32-bit sums");
    const int64_t n = (int64_t)rows * cols;
    const unsigned blocks = (unsigned)ceil_div64(n, 256);
    hsum_u32_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_in, rows, cols, kw, d_tmp);
    vsum_u32_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_tmp, rows, cols, kh, d_out);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_sum_u32(const uint32_t* d_in, int64_t n, uint64_t* d_out, void* stream) {
    SHG_CHECK(cudaMemsetAsync(d_out, 0, 8, as_stream(stream)));
    sum_u32_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(d_in, n, reinterpret_cast<unsigned long long*>(d_out));
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_select_u32(const uint32_t* d_vals, int64_t n, const int64_t* h_ranks, int n_ranks,
                              uint32_t* h_out, uint32_t* d_work256, void* stream) {
    SHG_REQUIRE(n > 0 && n_ranks >= 0 && n_ranks <= 64, "shg_select_u32: empty input or too many ranks");
    cudaStream_t st = as_stream(stream);
    unsigned int hist[256];
    std::vector<int64_t> rank(h_ranks, h_ranks + n_ranks);
    std::vector<uint32_t> prefix(n_ranks, 0);
    for (int q = 0; q < n_ranks; ++q)
        SHG_REQUIRE(rank[q] >= 0 && rank[q] < n, "shg_select_u32: rank %lld out of range", (long long)rank[q]);
    // MSB-first, one byte per pass; ranks that still share a prefix (the two middle elements of a
    // median, the two neighbours of a percentile) share the histogram of that pass
    for (int shift = 24; shift >= 0; shift -= 8) {
        std::vector<char> done(n_ranks, 0);
        for (int q = 0; q < n_ranks; ++q) {
            if (done[q]) continue;
            SHG_CHECK(cudaMemsetAsync(d_work256, 0, 256 * 4, st));
            radix_hist_kernel<<<grid_for(n), 256, 0, st>>>(d_vals, n, prefix[q], shift, d_work256);
            SHG_LAUNCH_CHECK();
            SHG_CHECK(cudaMemcpyAsync(hist, d_work256, 256 * 4, cudaMemcpyDeviceToHost, st));
            SHG_CHECK(cudaStreamSynchronize(st));
            const uint32_t group = prefix[q];
            for (int p = q; p < n_ranks; ++p) {
                if (done[p] || prefix[p] != group) continue;
                int b = 0;
                for (; b < 256; ++b) {
                    if (rank[p] < (int64_t)hist[b]) break;
                    rank[p] -= hist[b];
                }
                SHG_REQUIRE(b < 256, "shg_select_u32: internal error (histogram does not cover the rank)");
                prefix[p] = (group << 8) | (uint32_t)b;
                done[p] = 1;
            }
        }
    }
    for (int q = 0; q < n_ranks; ++q) h_out[q] = prefix[q];
    return 0;
}

extern "C" int shg_blur_range(const uint32_t* d_box, int64_t n, double scale, double ceiling, uint32_t* d_out2,
                              void* stream) {
    const unsigned int init[2] = {0xffffffffu, 0u};
    SHG_CHECK(cudaMemcpyAsync(d_out2, init, 8, cudaMemcpyHostToDevice, as_stream(stream)));
    blur_range_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(d_box, n, scale, ceiling, d_out2);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_blur_hist(const uint32_t* d_box, int64_t n, double scale, double ceiling, const double* h_edges,
                             int n_bins, uint64_t* d_counts, void* stream) {
    SHG_REQUIRE(n_bins >= 1 && n_bins <= 32, "shg_blur_hist: 1..32 bins");
    Edges ed;
    for (int i = 0; i <= n_bins; ++i) ed.e[i] = h_edges[i];
    ed.n_bins = n_bins;
    SHG_CHECK(cudaMemsetAsync(d_counts, 0, 32 * 8, as_stream(stream)));
    blur_hist_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(d_box, n, scale, ceiling, ed,
                                                               reinterpret_cast<unsigned long long*>(d_counts));
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_flood_smooth(const uint32_t* d_box, int rows, int cols, double scale, double level,
                                const double* h_weights, int radius, double eps, double* d_smoothed,
                                double* d_tmp2 /* 2 images */, void* stream) {
    SHG_REQUIRE(radius >= 0 && radius <= 32, "shg_flood_smooth: radius %d (max 32)", radius);
    GaussW g;
    for (int k = 0; k <= radius; ++k) g.w[k] = h_weights[k];
    g.radius = radius;
    const int64_t n = (int64_t)rows * cols;
    const unsigned blocks = (unsigned)ceil_div64(n, 256);
    cudaStream_t st = as_stream(stream);
    double* t0 = d_tmp2;
    double* t1 = d_tmp2 + n;
    // gaussian of the all-ones mask (the "bleed over" normalisation)
    gauss1d_kernel<true><<<blocks, 256, 0, st>>>(d_box, scale, level, 1, nullptr, rows, cols, 0, g, t0);
    gauss1d_kernel<false><<<blocks, 256, 0, st>>>(nullptr, 0, 0, 0, t0, rows, cols, 1, g, t1);
    // gaussian of the flood image
    gauss1d_kernel<true><<<blocks, 256, 0, st>>>(d_box, scale, level, 0, nullptr, rows, cols, 0, g, t0);
    gauss1d_kernel<false><<<blocks, 256, 0, st>>>(nullptr, 0, 0, 0, t0, rows, cols, 1, g, d_smoothed);
    divide_kernel<<<blocks, 256, 0, st>>>(d_smoothed, t1, eps, n);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_sobel_mag(const double* d_smoothed, int rows, int cols, double* d_gi, double* d_gj, double* d_mag,
                             void* stream) {
    const int64_t n = (int64_t)rows * cols;
    sobel_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(d_smoothed, rows, cols, d_gi, d_gj, d_mag);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_nms_candidates(const double* d_gi, const double* d_gj, const double* d_mag, int rows, int cols,
                                  double low, uint32_t* d_count, uint32_t cap, uint32_t* d_list_idx,
                                  double* d_list_mag, void* stream) {
    const int64_t n = (int64_t)rows * cols;
    SHG_CHECK(cudaMemsetAsync(d_count, 0, 4, as_stream(stream)));
    nms_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(d_gi, d_gj, d_mag, rows, cols, low, d_count,
                                                                          cap, d_list_idx, d_list_mag);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t shg_limb_state_bytes(void) { return (int64_t)sizeof(LimbState); }

extern "C" int shg_limb_front(const uint32_t* d_sums, int rows, int cols, int bw, const int64_t* h_ranks4,
                              double gamma, int n_bins, uint32_t* d_box, uint32_t* d_box5, uint32_t* d_tmp,
                              void* d_state, double* h_out, void* stream) {
    SHG_REQUIRE(d_sums && d_box && d_box5 && d_tmp && d_state && h_out && h_ranks4, "shg_limb_front: null argument");
    SHG_REQUIRE(n_bins >= 1 && n_bins <= 32, "shg_limb_front: 1..32 bins");
    SHG_REQUIRE(bw >= 1 && bw <= rows && bw <= cols && rows >= 5 && cols >= 5,
                "shg_limb_front: image %dx%d too small for the %d px blur", rows, cols, bw);
    SHG_REQUIRE((int64_t)bw * bw * (16LL * 65535) < 0xffffffffLL, "shg_limb_front: window too large for 32-bit sums");
    const int64_t n = (int64_t)rows * cols;
    for (int q = 0; q < 4; ++q)
        SHG_REQUIRE(h_ranks4[q] >= 0 && h_ranks4[q] < n, "shg_limb_front: rank %lld out of range", (long long)h_ranks4[q]);
    cudaStream_t st = as_stream(stream);
    LimbState* S = static_cast<LimbState*>(d_state);
    LimbState init;
    memset(&init, 0, offsetof(LimbState, hist));
    for (int q = 0; q < 4; ++q) init.rank[q] = h_ranks4[q];
    SHG_CHECK(cudaMemsetAsync(S, 0, sizeof(LimbState), st));
    // (the head of the block is small enough for the driver to embed the host copy in the command stream)
    SHG_CHECK(cudaMemcpyAsync(S, &init, offsetof(LimbState, hist), cudaMemcpyHostToDevice, st));
    const unsigned blocks = (unsigned)ceil_div64(n, 256);
    sum_u32_kernel<<<grid_for(n), 256, 0, st>>>(d_sums, n, &S->total);
    hsum_u32_kernel<<<blocks, 256, 0, st>>>(d_sums, rows, cols, bw, d_tmp);
    vsum_u32_kernel<<<blocks, 256, 0, st>>>(d_tmp, rows, cols, bw, d_box);
    hsum_u32_kernel<<<blocks, 256, 0, st>>>(d_sums, rows, cols, 5, d_tmp);
    vsum_u32_kernel<<<blocks, 256, 0, st>>>(d_tmp, rows, cols, 5, d_box5);
    const dim3 sel_grid(std::max(1u, grid_for(n) / 2), 2);
    for (int pass = 0; pass < 4; ++pass) {
        select_pass_kernel<<<sel_grid, 256, 0, st>>>(d_box, d_box5, n, pass, S);
        select_step_kernel<<<1, 32, 0, st>>>(pass, S);
    }
    const double scale = 1.0 / ((double)bw * bw);
    limb_ceiling_kernel<<<1, 1, 0, st>>>(scale, gamma, 1.0 - gamma, S);
    blur_range_dev_kernel<<<grid_for(n), 256, 0, st>>>(d_box, n, scale, S);
    limb_edges_kernel<<<1, 64, 0, st>>>(scale, n_bins, S);
    blur_hist_dev_kernel<<<grid_for(n), 256, 0, st>>>(d_box, n, scale, n_bins, S);
    SHG_LAUNCH_CHECK();
    LimbState head;
    SHG_CHECK(cudaMemcpyAsync(&head, S, offsetof(LimbState, hist), cudaMemcpyDeviceToHost, st));
    SHG_CHECK(cudaStreamSynchronize(st));
    h_out[0] = (double)head.total;
    for (int q = 0; q < 4; ++q) h_out[1 + q] = (double)head.prefix[q];
    h_out[5] = head.ceiling;
    h_out[6] = (double)head.range[0];
    h_out[7] = (double)head.range[1];
    for (int i = 0; i <= n_bins; ++i) h_out[8 + i] = head.edges[i];
    for (int i = 0; i < n_bins; ++i) h_out[8 + 33 + i] = (double)head.counts[i];
    return 0;
}

extern "C" int shg_limb_canny(const uint32_t* d_box, int rows, int cols, double scale, double level,
                              const double* h_weights, int radius, double eps, double low, double* d_buf6,
                              uint32_t* d_count, uint32_t cap, uint32_t* d_list_idx, double* d_list_mag,
                              uint32_t first_chunk, uint32_t* h_count, uint32_t* h_list_idx, double* h_list_mag,
                              void* stream) {
    SHG_REQUIRE(d_buf6 && d_count && d_list_idx && d_list_mag && h_count && h_list_idx && h_list_mag,
                "shg_limb_canny: null argument");
    const int64_t n = (int64_t)rows * cols;
    cudaStream_t st = as_stream(stream);
    double* smoothed = d_buf6;
    SHG_REQUIRE(radius >= 0 && radius <= 32, "shg_limb_canny: radius %d (max 32)", radius);
    if (getenv("SHG_LIMB_UNFUSED")) {                           // the seven-kernel formulation (cross-check / timing)
        int rc = shg_flood_smooth(d_box, rows, cols, scale, level, h_weights, radius, eps, smoothed, d_buf6 + n, stream);
        if (rc) return rc;
        rc = shg_sobel_mag(smoothed, rows, cols, d_buf6 + 3 * n, d_buf6 + 4 * n, d_buf6 + 5 * n, stream);
        if (rc) return rc;
        rc = shg_nms_candidates(d_buf6 + 3 * n, d_buf6 + 4 * n, d_buf6 + 5 * n, rows, cols, low, d_count, cap, d_list_idx,
                                d_list_mag, stream);
        if (rc) return rc;
    } else {
        GaussW g;
        for (int k = 0; k <= radius; ++k) g.w[k] = h_weights[k];
        g.radius = radius;
        const int fw = kGT_W + 2 * radius, fh = kGT_H + 2 * radius;
        const size_t smem = (size_t)(kGT_H * fw + kGT_H) * sizeof(double) + (size_t)fh * fw;
        SHG_CHECK(cudaFuncSetAttribute(flood_gauss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        flood_gauss_kernel<<<dim3((cols + kGT_W - 1) / kGT_W, (rows + kGT_H - 1) / kGT_H), 256, smem, st>>>(
            d_box, rows, cols, scale, level, g, eps, smoothed);
        SHG_CHECK(cudaMemsetAsync(d_count, 0, 4, st));
        sobel_nms_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(smoothed, rows, cols, low, d_count, cap, d_list_idx,
                                                                     d_list_mag);
        SHG_LAUNCH_CHECK();
    }
    // the count and the head of the list travel together: one round trip for the usual ~10^4 candidates
    const uint32_t head = std::min(first_chunk, cap);
    SHG_CHECK(cudaMemcpyAsync(h_count, d_count, 4, cudaMemcpyDeviceToHost, st));
    SHG_CHECK(cudaMemcpyAsync(h_list_idx, d_list_idx, (size_t)head * 4, cudaMemcpyDeviceToHost, st));
    SHG_CHECK(cudaMemcpyAsync(h_list_mag, d_list_mag, (size_t)head * 8, cudaMemcpyDeviceToHost, st));
    SHG_CHECK(cudaStreamSynchronize(st));
    const uint32_t have = std::min(*h_count, cap);
    if (have > head) {
        SHG_CHECK(cudaMemcpyAsync(h_list_idx + head, d_list_idx + head, (size_t)(have - head) * 4,
                                  cudaMemcpyDeviceToHost, st));
        SHG_CHECK(cudaMemcpyAsync(h_list_mag + head, d_list_mag + head, (size_t)(have - head) * 8,
                                  cudaMemcpyDeviceToHost, st));
        SHG_CHECK(cudaStreamSynchronize(st));
    }
    return 0;
}

// ---- host helper: strict convex-hull vertices of integer points -------------
// (scipy.spatial.ConvexHull(points).vertices as a SET, reference ellipse_to_circle.py:263: the
// reference only asks which edge regions own a hull vertex.)  Andrew's monotone chain with an exact
// 64-bit cross product; collinear points on a hull edge are not vertices, as in Qhull.
extern "C" int shg_hull_vertices(const int64_t* h_xy, int64_t n, int64_t* h_vertex_index, int64_t* h_n_vertices) {
    SHG_REQUIRE(h_xy && h_vertex_index && h_n_vertices && n >= 0, "shg_hull_vertices: bad arguments");
    std::vector<int64_t> order((size_t)n);
    for (int64_t i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        if (h_xy[2 * a] != h_xy[2 * b]) return h_xy[2 * a] < h_xy[2 * b];
        if (h_xy[2 * a + 1] != h_xy[2 * b + 1]) return h_xy[2 * a + 1] < h_xy[2 * b + 1];
        return a < b;
    });
    // duplicates: keep the first occurrence
    std::vector<int64_t> uniq;
    uniq.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t a = order[i];
        if (!uniq.empty()) {
            const int64_t b = uniq.back();
            if (h_xy[2 * a] == h_xy[2 * b] && h_xy[2 * a + 1] == h_xy[2 * b + 1]) continue;
        }
        uniq.push_back(a);
    }
    const int64_t m = (int64_t)uniq.size();
    if (m < 3) {
        for (int64_t i = 0; i < m; ++i) h_vertex_index[i] = uniq[i];
        *h_n_vertices = m;
        return 0;
    }
    auto cross = [&](int64_t o, int64_t a, int64_t b) -> __int128 {
        const __int128 ax = h_xy[2 * a] - h_xy[2 * o], ay = h_xy[2 * a + 1] - h_xy[2 * o + 1];
        const __int128 bx = h_xy[2 * b] - h_xy[2 * o], by = h_xy[2 * b + 1] - h_xy[2 * o + 1];
        return ax * by - ay * bx;
    };
    std::vector<int64_t> hull((size_t)(2 * m));
    int64_t k = 0;
    for (int64_t i = 0; i < m; ++i) {
        while (k >= 2 && cross(hull[k - 2], hull[k - 1], uniq[i]) <= 0) --k;
        hull[k++] = uniq[i];
    }
    for (int64_t i = m - 2, t = k + 1; i >= 0; --i) {
        while (k >= t && cross(hull[k - 2], hull[k - 1], uniq[i]) <= 0) --k;
        hull[k++] = uniq[i];
    }
    --k;                                                            // the last point repeats the first
    for (int64_t i = 0; i < k; ++i) h_vertex_index[i] = hull[i];
    *h_n_vertices = k;
    return 0;
}
