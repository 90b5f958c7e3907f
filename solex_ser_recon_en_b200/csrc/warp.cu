// Circularisation warp and the 4x4 downscale that feeds the ellipse fit.
//
// The reference warps with skimage.transform.warp(image, ProjectiveTransform(mat3),
// order 1, constant mode) (ellipse_to_circle.py:94-118).  get_correction_matrix
// (ellipse_to_circle.py:39-50) always yields mat3 rows 1,2 == [0,1,0],[0,0,1]:
// every output row samples only its own input row, so the warp is a 1-D
// linear resample along the frame axis:
//     x = (m00*c + m01*r) + m02
//     out[r][c] = trunc(clip((1-d)*in[r][floor x] + d*in[r][ceil x], lo, hi)),  d = x - floor x
// with taps outside the image reading cval = image[0][0].  The disk arrives
// frame-major (N, ih), so a CTA stages [frame span] x [64 slit rows] in shared
// memory with coalesced loads and writes (ih, Wout) rows with coalesced stores:
// the warp doubles as the transpose to the reference layout.
// Bound: HBM, ih*N*2 bytes read + ih*Wout*2 bytes written per image.
#include <limits.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace {

constexpr int kRows = 64;      // slit rows per tile
constexpr int kPitch = kRows + 2;

// One CTA = 64 slit rows x COLS output columns of one image (blockIdx.z).  The
// tile width shrinks as the stretch m00 grows so that the staged frame span
// stays ~45 KB and several CTAs per SM overlap their loads with the fp64 lerp.
template <int COLS, bool SHARDED>
__global__ void __launch_bounds__(256)
warp_rows_kernel(const uint16_t* __restrict__ disk_base, int64_t disk_stride, const int32_t* __restrict__ sel,
                 int64_t n_frames, int ih, int flip, double m00, double m01, double m02,
                 const uint32_t* __restrict__ minmax /* [n_imgs][2] */,
                 uint16_t* __restrict__ out_base, int64_t out_stride, int out_rows, int out_cols, int span,
                 const uint32_t* __restrict__ cval_arr /* [n_imgs] or null */, int own_lo, int own_hi,
                 const unsigned long long* __restrict__ out_ptrs /* [n_imgs] or null */) {
    extern __shared__ uint16_t tile[];           // [span][kPitch]
    const int img = blockIdx.z;
    // (frame-sharded scans pass a base pointer such that disk + k*ih is frame k of the WHOLE scan.  This rank
    // produces exactly the output pixels whose left tap floor(x) lies in its frames [own_lo, own_hi) -- logical
    // frame order; INT_MIN / INT_MAX = open end -- so only frames [own_lo, own_hi] are touched: its own plus ONE
    // frame of the next rank, whatever the tilt.  image[0][0] then comes from cval_arr and the output image from
    // out_ptrs, possibly on a peer GPU: every pixel of it is written by exactly one rank.)
    const uint16_t* disk = disk_base + (int64_t)(sel ? sel[img] : img) * disk_stride;
    uint16_t* out = out_ptrs ? reinterpret_cast<uint16_t*>(out_ptrs[img]) : out_base + (int64_t)img * out_stride;
    const double cval = cval_arr ? u32_to_double(cval_arr[img])
                                 : u32_to_double(disk[(flip ? (n_frames - 1) : 0) * (int64_t)ih]);   // image[0][0]
    const int r0 = blockIdx.y * kRows;
    const int r1 = min(r0 + kRows, out_rows);
    // columns that can hold a pixel of this rank in rows [r0, r1): x = m00*c + m01*r + m02 in [own_lo, own_hi)
    int cb = 0, ce = out_cols;
    {
        const double s0 = m01 * (double)r0, s1 = m01 * (double)(r1 - 1);
        if (own_lo != INT_MIN) cb = max(0, (int)fmax(floor(((double)own_lo - fmax(s0, s1) - m02) / m00) - 1.0, -1.0e9));
        if (own_hi != INT_MAX) ce = min(out_cols, (int)fmin(ceil(((double)own_hi - fmin(s0, s1) - m02) / m00) + 1.0, 1.0e9));
    }
    const int c0 = cb + blockIdx.x * COLS;
    if (c0 >= ce) return;
    const int c1 = min(c0 + COLS, ce);
    // frame range touched by this tile (x is monotone in c and in r)
    double xa = 1e300, xb = -1e300;
    {
        const double cs[2] = {(double)c0, (double)(c1 - 1)};
        const double rs[2] = {(double)r0, (double)(r1 - 1)};
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
                const double x = __dadd_rn(__dadd_rn(__dmul_rn(m00, cs[a]), __dmul_rn(m01, rs[b])), m02);
                xa = fmin(xa, x);
                xb = fmax(xb, x);
            }
    }
    // clamp before converting: far outside the image everything is cval anyway
    xa = fmax(xa, -4.0);
    xb = fmin(xb, (double)n_frames + 4.0);
    if (own_lo != INT_MIN) xa = fmax(xa, (double)own_lo);           // taps of this rank's pixels: frames [own_lo, own_hi]
    if (own_hi != INT_MAX) xb = fmin(xb, (double)own_hi);
    const int64_t kbase = (int64_t)floor(xa);
    const int64_t kend = min((int64_t)ceil(xb) + 1, kbase + span);   // exclusive
    const int nrow = min(r1, ih) - r0;                                // valid slit rows (may be <= 0)
    // ---- stage: coalesced along the slit axis (64 rows = 128 B per frame) ----
    const int n_stage = (int)(kend - kbase);
    const bool interior = kbase >= 0 && kend <= n_frames && r1 <= ih && r1 - r0 == kRows;
    if (nrow == kRows && (ih & 7) == 0 && (((uintptr_t)disk) & 15) == 0) {
        // 16-byte loads: 8 lanes cover one frame's 64 rows, a warp takes 4 frames per instruction
        const int seg = threadIdx.x & 7;
        const int64_t step = flip ? -(int64_t)ih : (int64_t)ih;
        const uint16_t* src0 = disk + (flip ? (n_frames - 1 - kbase) : kbase) * (int64_t)ih + r0 + seg * 8;
        if (interior) {
            // every staged frame exists: two independent loads in flight per thread, no range tests
            int f = threadIdx.x >> 3;
            for (; f + 32 < n_stage; f += 64) {
                const uint4 v0 = ld_stream_u4(reinterpret_cast<const uint4*>(src0 + f * step));
                const uint4 v1 = ld_stream_u4(reinterpret_cast<const uint4*>(src0 + (f + 32) * step));
                uint32_t* d0 = reinterpret_cast<uint32_t*>(tile + f * kPitch + seg * 8);
                uint32_t* d1 = reinterpret_cast<uint32_t*>(tile + (f + 32) * kPitch + seg * 8);
                d0[0] = v0.x; d0[1] = v0.y; d0[2] = v0.z; d0[3] = v0.w;
                d1[0] = v1.x; d1[1] = v1.y; d1[2] = v1.z; d1[3] = v1.w;
            }
            if (f < n_stage) {
                const uint4 v0 = ld_stream_u4(reinterpret_cast<const uint4*>(src0 + f * step));
                uint32_t* d0 = reinterpret_cast<uint32_t*>(tile + f * kPitch + seg * 8);
                d0[0] = v0.x; d0[1] = v0.y; d0[2] = v0.z; d0[3] = v0.w;
            }
        } else {
            for (int f = threadIdx.x >> 3; f < n_stage; f += 32) {
                const int64_t k = kbase + f;
                if (k < 0 || k >= n_frames) continue;                      // never read (range test below)
                const uint4 v0 = ld_stream_u4(reinterpret_cast<const uint4*>(src0 + f * step));
                uint32_t* d0 = reinterpret_cast<uint32_t*>(tile + f * kPitch + seg * 8);
                d0[0] = v0.x; d0[1] = v0.y; d0[2] = v0.z; d0[3] = v0.w;
            }
        }
    } else {
        for (int64_t kk = kbase + (threadIdx.x >> 5); kk < kend; kk += 8) {
            uint16_t* dst = tile + (kk - kbase) * kPitch;
            if (kk < 0 || kk >= n_frames) continue;                        // never read (range test below)
            const int64_t ksrc = flip ? (n_frames - 1 - kk) : kk;
            const uint16_t* src = disk + ksrc * ih + r0;
            for (int j = (threadIdx.x & 31) * 2; j < nrow; j += 64) {
                if (j + 1 < nrow && (((uintptr_t)(src + j)) & 3) == 0) {
                    const uint32_t v = *reinterpret_cast<const uint32_t*>(src + j);
                    *reinterpret_cast<uint32_t*>(dst + j) = v;
                } else {
                    dst[j] = src[j];
                    if (j + 1 < nrow) dst[j + 1] = src[j + 1];
                }
            }
        }
    }
    __syncthreads();
    constexpr int RGROUPS = 256 / COLS;                                // row groups working side by side
    const int c = c0 + (threadIdx.x % COLS);
    if (c >= c1) return;
    const double mc = __dmul_rn(m00, (double)c);
    const int ilo = (int)minmax[2 * img], ihi = (int)minmax[2 * img + 1];
    const int kb = (int)kbase, nf = (int)n_frames;
    constexpr double kMagic = 6755399441055744.0;                      // 1.5 * 2^52: x + kMagic holds floor(x) in its low word
    const int rg = threadIdx.x / COLS;
    if (interior) {
        // all taps of the tile are staged and valid: no range tests, no int -> double conversions in the loop
        double rd = (double)(r0 + rg);
        const uint16_t* col = tile + rg - kb * kPitch;
        uint16_t* o = out + (int64_t)(r0 + rg) * out_cols + c;
#pragma unroll 4
        for (int r = r0 + rg; r < r1; r += RGROUPS) {
            const double x = __dadd_rn(__dadd_rn(mc, __dmul_rn(m01, rd)), m02);
            const double xm = __dadd_rd(x, kMagic);
            const int kf = __double2loint(xm);
            if (!SHARDED || (kf >= own_lo && kf < own_hi)) {            // (always, unless the scan is frame-sharded)
                const double d = __dsub_rn(x, __dsub_rn(xm, kMagic));
                const uint16_t* p = col + kf * kPitch;
                const double L = u32_to_double(p[0]);
                const double R = u32_to_double(p[d != 0.0 ? kPitch : 0]);
                const double v = __dadd_rn(__dmul_rn(__dsub_rn(1.0, d), L), __dmul_rn(d, R));
                int q = __double2loint(__dadd_rd(v, kMagic));
                q = min(max(q, ilo), ihi);
                *o = (uint16_t)q;
            }
            rd += (double)RGROUPS;
            col += RGROUPS;
            o += (int64_t)RGROUPS * out_cols;
        }
        return;
    }
    for (int r = r0 + rg; r < r1; r += RGROUPS) {
        const double x = __dadd_rn(__dadd_rn(mc, __dmul_rn(m01, (double)r)), m02);
        const double xm = __dadd_rd(x, kMagic);
        const int kf = __double2loint(xm);                              // floor(x) (|x| < 2^31)
        if (SHARDED && (kf < own_lo || kf >= own_hi)) continue;         // another rank's pixel (frame-sharded scans)
        const double d = __dsub_rn(x, __dsub_rn(xm, kMagic));          // x - floor(x), exact subtraction of the integer
        const int kc = kf + (d != 0.0 ? 1 : 0);                         // ceil(x)
        double L = cval, R = cval;
        if (r < ih) {
            const uint16_t* col = tile + (r - r0);
            if (kf >= 0 && kf < nf) L = u32_to_double(col[(kf - kb) * kPitch]);
            if (kc >= 0 && kc < nf) R = u32_to_double(col[(kc - kb) * kPitch]);
        }
        const double v = __dadd_rn(__dmul_rn(__dsub_rn(1.0, d), L), __dmul_rn(d, R));
        // trunc(clip(v, lo, hi)) == clamp(floor(v), lo, hi) because lo and hi are integers
        int q = __double2loint(__dadd_rd(v, kMagic));
        q = min(max(q, ilo), ihi);
        out[(int64_t)r * out_cols + c] = (uint16_t)q;
    }
}

// out[ri][ci] = sum over the 4x4 block of image[r][k] = disk[k][r] (zero padded)
__global__ void __launch_bounds__(256)
downscale4_kernel(const uint16_t* __restrict__ disk, int64_t n_frames, int ih, int flip,
                  uint32_t* __restrict__ out, int out_rows, int out_cols) {
    // thread -> (ci, ri) with ri fastest so the 4 slit rows of one frame are one 8-byte run
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)out_rows * out_cols) return;
    const int ri = (int)(idx % out_rows);
    const int64_t ci = idx / out_rows;
    uint32_t s = 0;
    for (int dk = 0; dk < 4; ++dk) {
        const int64_t k = ci * 4 + dk;
        if (k >= n_frames) break;
        const uint16_t* p = disk + (flip ? (n_frames - 1 - k) : k) * ih + ri * 4;
        for (int dr = 0; dr < 4; ++dr)
            if (ri * 4 + dr < ih) s += p[dr];
    }
    out[(int64_t)ri * out_cols + ci] = s;
}

}  // namespace

extern "C" int shg_warp_rows(const uint16_t* d_disk, int64_t disk_stride, const int32_t* d_sel, int n_imgs,
                             int64_t n_frames, int ih, int flip, double m00, double m01, double m02,
                             const uint32_t* d_minmax, uint16_t* d_out, int64_t out_stride, int out_rows,
                             int out_cols, void* stream) {
    return shg_warp_rows_window(d_disk, disk_stride, d_sel, n_imgs, n_frames, ih, flip, m00, m01, m02, d_minmax, d_out,
                                out_stride, out_rows, out_cols, nullptr, INT_MIN, INT_MAX, nullptr, stream);
}

extern "C" int shg_warp_rows_window(const uint16_t* d_disk, int64_t disk_stride, const int32_t* d_sel, int n_imgs,
                                    int64_t n_frames, int ih, int flip, double m00, double m01, double m02,
                                    const uint32_t* d_minmax, uint16_t* d_out, int64_t out_stride, int out_rows,
                                    int out_cols, const uint32_t* d_cval, int own_lo, int own_hi,
                                    const uint64_t* d_out_ptrs, void* stream) {
    SHG_REQUIRE(n_frames > 0 && ih > 0 && out_rows > 0 && out_cols > 0 && n_imgs > 0, "shg_warp_rows: bad geometry");
    SHG_REQUIRE(own_lo < own_hi, "shg_warp_rows: empty frame range [%d, %d)", own_lo, own_hi);
    SHG_REQUIRE(d_out || d_out_ptrs, "shg_warp_rows: no output");
    SHG_REQUIRE(std::isfinite(m00) && std::isfinite(m01) && std::isfinite(m02) && m00 > 0.0,
                "shg_warp_rows: bad matrix (%g, %g, %g)", m00, m01, m02);
    SHG_REQUIRE(n_imgs <= 65535, "shg_warp_rows: too many images");
    // frames spanned by one tile: COLS columns and kRows rows, + floor/ceil taps
    auto span_for = [&](int cols) { return m00 * (cols - 1) + std::fabs(m01) * (kRows - 1) + 4.0; };
    int cols = 256;
    while (cols > 64 && span_for(cols) * kPitch * 2 > 48.0 * 1024) cols /= 2;
    const double spanf = span_for(cols);
    SHG_REQUIRE(spanf * kPitch * 2 < 200.0 * 1024, "shg_warp_rows: stretch %g / shear %g too large for the tile", m00,
                m01);
    const int span = (int)std::ceil(spanf);
    const size_t smem = (size_t)span * kPitch * sizeof(uint16_t);
    // widest column window over the row tiles (the kernel derives each tile's window with the same formulas)
    int max_width = out_cols;
    if (own_lo != INT_MIN || own_hi != INT_MAX) {
        max_width = 1;
        for (int r0 = 0; r0 < out_rows; r0 += kRows) {
            const int r1 = std::min(r0 + kRows, out_rows);
            const double s0 = m01 * (double)r0, s1 = m01 * (double)(r1 - 1);
            int cb = 0, ce = out_cols;
            if (own_lo != INT_MIN)
                cb = std::max(0, (int)std::max(std::floor(((double)own_lo - std::max(s0, s1) - m02) / m00) - 1.0, -1.0e9));
            if (own_hi != INT_MAX)
                ce = std::min(out_cols, (int)std::min(std::ceil(((double)own_hi - std::min(s0, s1) - m02) / m00) + 1.0, 1.0e9));
            max_width = std::max(max_width, ce - cb);
        }
    }
    dim3 grid((max_width + cols - 1) / cols, (out_rows + kRows - 1) / kRows, n_imgs);
    SHG_REQUIRE(grid.y <= 65535, "shg_warp_rows: too many rows");
    cudaStream_t st = as_stream(stream);
#define SHG_WARP_LAUNCH(C, S)                                                                                         \
    do {                                                                                                            \
        SHG_CHECK(cudaFuncSetAttribute(warp_rows_kernel<C, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        warp_rows_kernel<C, S><<<grid, 256, smem, st>>>(d_disk, disk_stride, d_sel, n_frames, ih, flip, m00, m01, m02,  \
                                                     d_minmax, d_out, out_stride, out_rows, out_cols, span, d_cval, \
                                                     own_lo, own_hi,                                                \
                                                     reinterpret_cast<const unsigned long long*>(d_out_ptrs));      \
    } while (0)
    const bool sharded = own_lo != INT_MIN || own_hi != INT_MAX;     // the ownership test compiles out otherwise
    if (sharded) {
        if (cols == 256) SHG_WARP_LAUNCH(256, true);
        else if (cols == 128) SHG_WARP_LAUNCH(128, true);
        else SHG_WARP_LAUNCH(64, true);
    } else {
        if (cols == 256) SHG_WARP_LAUNCH(256, false);
        else if (cols == 128) SHG_WARP_LAUNCH(128, false);
        else SHG_WARP_LAUNCH(64, false);
    }
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_downscale4_sum(const uint16_t* d_disk, int64_t n_frames, int ih, int flip,
                                  uint32_t* d_out, int out_rows, int out_cols, void* stream) {
    SHG_REQUIRE(out_rows == (ih + 3) / 4 && out_cols == (int)((n_frames + 3) / 4),
                "shg_downscale4_sum: output must be ceil(ih/4) x ceil(N/4)");
    const int64_t n = (int64_t)out_rows * out_cols;
    downscale4_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(d_disk, n_frames, ih, flip, d_out,
                                                                                  out_rows, out_cols);
    SHG_LAUNCH_CHECK();
    return 0;
}
