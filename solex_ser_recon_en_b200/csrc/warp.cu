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
#include <cuda.h>
#include <limits.h>
#include <stdlib.h>

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

// ---------------------------------------------------------------------------
// TMA variant (shg_warp_rows_tma).  Same arithmetic; different data movement:
//  * the [frame span] x [64 slit rows] input tile is staged by TMA (cp.async.bulk.tensor boxes of 64 rows x 64
//    frames, one mbarrier): no loader instructions, dense 128-byte tile rows;
//  * a LANE owns a SLIT ROW (consecutive lanes read consecutive 2-byte words of one tile row: conflict-free on
//    the dense layout) and walks output columns, so m01*r is computed once per thread and eight consecutive
//    output pixels leave as ONE 16-byte store.  Output rows are not 16-byte aligned in general (out_cols is
//    arbitrary), so the column tiles are laid out per row on the row's own 16-byte grid: for row r the tile
//    boundaries sit at columns c with (r*out_cols + c) % 8 == 0;
//  * ~26 instructions per output pixel instead of ~66 (loader + per-pixel 2-byte stores).
constexpr int kBoxK = 32;                 // frames per TMA box
constexpr double kMagicD = 6755399441055744.0;

__device__ __forceinline__ uint32_t warp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kMaxStages = 4;

// Persistent CTAs (one per SM): a CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of the
// (column tile, row tile, image) grid through a ring of STAGES shared-memory stages; thread 0 issues the TMA boxes
// of tile i + STAGES as soon as the block has finished tile i, so the loads of the next tiles are in flight while
// this one is computed (a one-tile-per-CTA version spent 19 % of its stall samples waiting for its own load).
template <int COLS, int NT, bool SHARDED>
__global__ void __launch_bounds__(NT)
warp_tma_kernel(const __grid_constant__ CUtensorMap map, const int32_t* __restrict__ sel, int64_t n_frames,
                int64_t frame_origin, int ih, int flip, double m00, double m01, double m02,
                const uint32_t* __restrict__ minmax, uint16_t* __restrict__ out_base, int64_t out_stride, int out_rows,
                int out_cols, int span_boxes, int n_stages, const uint32_t* __restrict__ cval_arr,
                const uint16_t* __restrict__ first_px /* image [0][0] of image 0 of the map; stride below */,
                int64_t disk_stride, int own_lo, int own_hi, const unsigned long long* __restrict__ out_ptrs,
                int col_base /* multiple of 8: first column of tile 0 (minus the row phase) */, int col_end,
                int ntx, int nry, int n_imgs) {
    extern __shared__ unsigned char smem_raw[];
    uint16_t* ring = reinterpret_cast<uint16_t*>(smem_raw + ((128u - (warp_smem_u32(smem_raw) & 127u)) & 127u));
    __shared__ __align__(8) uint64_t full[kMaxStages];
    __shared__ int s_kb[kMaxStages], s_ke[kMaxStages];       // frame range staged for the tile in each stage
    const int stage_elems = span_boxes * kBoxK * kRows;
    const int cend = min(out_cols, col_end);
    const int64_t total = (int64_t)ntx * nry * n_imgs;

    // thread 0: frame range of tile t, published in s_kb / s_ke, and its TMA boxes
    auto issue = [&](int64_t t, int stage) {
        const int tx = (int)(t % ntx), ry = (int)((t / ntx) % nry), img = (int)(t / ((int64_t)ntx * nry));
        const int r0 = ry * kRows, r1 = min(r0 + kRows, out_rows);
        const int tc0 = col_base + tx * COLS;
        const int cmin = max(tc0 - 7, 0), cmax = min(tc0 + COLS, cend) - 1;
        int kbase = 0, kend = INT_MIN;                   // INT_MIN: no pixel of this rank (or no column) in the tile
        if (cmin <= cmax) {
            double xa = 1e300, xb = -1e300;              // x is monotone in c and in r
            const double cs[2] = {(double)cmin, (double)cmax};
            const double rs[2] = {(double)r0, (double)(r1 - 1)};
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                    const double x = __dadd_rn(__dadd_rn(__dmul_rn(m00, cs[a]), __dmul_rn(m01, rs[b])), m02);
                    xa = fmin(xa, x);
                    xb = fmax(xb, x);
                }
            if (!SHARDED || !(floor(xb) < (double)own_lo || floor(xa) >= (double)own_hi)) {
                xa = fmax(xa, -4.0);
                xb = fmin(xb, (double)n_frames + 4.0);
                if (SHARDED) {
                    if (own_lo != INT_MIN) xa = fmax(xa, (double)own_lo);
                    if (own_hi != INT_MAX) xb = fmin(xb, (double)own_hi);
                }
                kbase = (int)floor(xa);
                // one frame more than the last right tap: the fast path reads frame floor(x) + 1 even when x is an
                // integer (its weight is then exactly 0); a tile beside the image has xb < xa: nothing staged
                kend = max(kbase, min((int)ceil(xb) + 2, kbase + span_boxes * kBoxK));      // exclusive
            }
        }
        s_kb[stage] = kbase;
        s_ke[stage] = kend;
        const int n_box = kend == INT_MIN ? 0 : (kend - kbase + kBoxK - 1) / kBoxK;
        // physical frame of logical frame k: flip ? n-1-k : k; the boxes cover physical [plo, plo + n_box*kBoxK)
        const int64_t plo = flip ? (n_frames - kend) : (int64_t)kbase;
        const int z = sel ? sel[img] : img;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(warp_smem_u32(&full[stage])),
                     "r"((uint32_t)(n_box * kBoxK * kRows * 2)) : "memory");
        uint16_t* dst = ring + (size_t)stage * stage_elems;
        for (int b = 0; b < n_box; ++b) {
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(warp_smem_u32(dst + (size_t)b * kBoxK * kRows)), "l"(&map), "r"(r0),
                "r"((int)(plo - frame_origin) + b * kBoxK), "r"(z), "r"(warp_smem_u32(&full[stage]))
                : "memory");
        }
    };

    if (threadIdx.x == 0) {
        for (int s_ = 0; s_ < n_stages; ++s_)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(warp_smem_u32(&full[s_])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s_ = 0; s_ < n_stages; ++s_) {
            const int64_t t = (int64_t)blockIdx.x + (int64_t)s_ * gridDim.x;
            if (t < total) issue(t, s_);
        }
    }
    __syncthreads();                                  // barriers initialised before anybody polls them

    const int rl = threadIdx.x & (kRows - 1), cq = threadIdx.x >> 6;
    constexpr int CQ = COLS / (NT / kRows);                  // columns per thread and tile: groups of 8
    static_assert(CQ >= 8 && CQ % 8 == 0, "a thread owns whole 16-byte groups");
    const int nf = (int)n_frames;
    const int kstep = flip ? -kRows : kRows;
    int it = 0;
    for (int64_t t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int stage = it % n_stages;
        const uint32_t phase = (uint32_t)((it / n_stages) & 1);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(warp_smem_u32(&full[stage])), "r"(phase) : "memory");
        const int kb = s_kb[stage], ke = s_ke[stage];
        const int tx = (int)(t % ntx), ry = (int)((t / ntx) % nry), img = (int)(t / ((int64_t)ntx * nry));
        const int r = ry * kRows + rl;
        if (ke != INT_MIN && r < out_rows) {
            const int tc0 = col_base + tx * COLS;
            uint16_t* out = out_ptrs ? reinterpret_cast<uint16_t*>(out_ptrs[img]) : out_base + (int64_t)img * out_stride;
            const int ilo = (int)minmax[2 * img], ihi = (int)minmax[2 * img + 1];
            // this row's 16-byte grid: element offset r*out_cols + c is a multiple of 8 at the group starts
            const int phase8 = (int)(((int64_t)r * out_cols + col_base) & 7);
            const int cs = tc0 - phase8 + cq * CQ;
            const double tr = __dmul_rn(m01, (double)r);
            // smem row of logical frame k: its physical frame minus plo, i.e. flip ? ke-1-k : k - kb
            const uint16_t* tile = ring + (size_t)stage * stage_elems;
            const uint16_t* t0 = flip ? tile + (int64_t)(ke - 1) * kRows + rl : tile - (int64_t)kb * kRows + rl;
            const bool interior = ke > kb && kb >= 0 && ke <= nf && r < ih;
            uint16_t* orow = out + (int64_t)r * out_cols;
            const bool vec_ok = (reinterpret_cast<uintptr_t>(out) & 15) == 0;   // (images handed in by address may not be)
#pragma unroll 1
            for (int g = 0; g < CQ / 8; ++g) {
                const int c0 = cs + g * 8;
                if (c0 >= cend || c0 + 8 <= 0) continue;
                bool fast = interior && c0 >= 0 && c0 + 8 <= cend;
                if (fast) {
                    // all eight columns inside the image: are all their taps staged (and, for frame-sharded scans,
                    // the first and last left tap this rank's)?  x is monotone in c.
                    const double xf = __dadd_rn(__dadd_rn(__dmul_rn(m00, (double)c0), tr), m02);
                    const double xl = __dadd_rn(__dadd_rn(__dmul_rn(m00, (double)(c0 + 7)), tr), m02);
                    const int kfirst = __double2loint(__dadd_rd(xf, kMagicD));
                    const int klast = __double2loint(__dadd_rd(xl, kMagicD));
                    fast = kfirst >= kb && klast + 1 < ke;
                    if (SHARDED) fast = fast && kfirst >= own_lo && klast < own_hi;
                }
                if (fast) {
                    uint32_t packed[4];
                    const double cd0 = (double)c0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double x = __dadd_rn(__dadd_rn(__dmul_rn(m00, cd0 + (double)j), tr), m02);
                        const double xm = __dadd_rd(x, kMagicD);
                        const int kf = __double2loint(xm);
                        const double d = __dsub_rn(x, __dsub_rn(xm, kMagicD));
                        const uint16_t* p = t0 + kf * kstep;
                        const double L = u32_to_double(p[0]);
                        const double R = u32_to_double(p[kstep]);      // weight d == 0 exactly when x is an integer
                        const double v = __dadd_rn(__dmul_rn(__dsub_rn(1.0, d), L), __dmul_rn(d, R));
                        int q = __double2loint(__dadd_rd(v, kMagicD));
                        q = min(max(q, ilo), ihi);
                        if (j & 1) packed[j >> 1] |= (uint32_t)q << 16;
                        else packed[j >> 1] = (uint32_t)q;
                    }
                    if (vec_ok) {
                        *reinterpret_cast<uint4*>(orow + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            orow[c0 + 2 * j] = (uint16_t)packed[j];
                            orow[c0 + 2 * j + 1] = (uint16_t)(packed[j] >> 16);
                        }
                    }
                    continue;
                }
                // general path: image / tile / ownership boundaries
                const double cval = cval_arr ? u32_to_double(cval_arr[img])
                                             : u32_to_double(first_px[(int64_t)(sel ? sel[img] : img) * disk_stride]);
#pragma unroll 1
                for (int c = max(c0, 0); c < min(c0 + 8, cend); ++c) {
                    const double x = __dadd_rn(__dadd_rn(__dmul_rn(m00, (double)c), tr), m02);
                    const double xm = __dadd_rd(x, kMagicD);
                    const int kf = __double2loint(xm);
                    if (SHARDED && (kf < own_lo || kf >= own_hi)) continue;
                    const double d = __dsub_rn(x, __dsub_rn(xm, kMagicD));
                    const int kc = kf + (d != 0.0 ? 1 : 0);
                    double L = cval, R = cval;
                    if (r < ih) {
                        if (kf >= 0 && kf < nf && kf >= kb && kf < ke) L = u32_to_double(t0[kf * kstep]);
                        if (kc >= 0 && kc < nf && kc >= kb && kc < ke) R = u32_to_double(t0[kc * kstep]);
                    }
                    const double v = __dadd_rn(__dmul_rn(__dsub_rn(1.0, d), L), __dmul_rn(d, R));
                    int q = __double2loint(__dadd_rd(v, kMagicD));
                    q = min(max(q, ilo), ihi);
                    orow[c] = (uint16_t)q;
                }
            }
        }
        __syncthreads();                                  // everyone is done with this stage
        if (threadIdx.x == 0) {
            const int64_t nxt = t + (int64_t)n_stages * gridDim.x;
            if (nxt < total) issue(nxt, stage);
        }
    }
}

// Copy this rank's pixels of every circularised image from its local full-width buffer into the rank that owns
// the image (peer memory): for row r the owned pixels are the columns whose left tap floor(x) lies in
// [own_lo, own_hi) -- one interval, x being monotone in c.  A warp copies one row interval with 16-byte
// loads / stores (local and remote images share element offsets, hence alignment): 512-byte contiguous runs
// over NVLink instead of the 16- or 64-byte pieces a warp kernel's own stores would be.
__global__ void __launch_bounds__(256)
exchange_rows_kernel(const uint16_t* __restrict__ local, int64_t local_stride, int out_rows, int out_cols, double m00,
                     double m01, double m02, int own_lo, int own_hi,
                     const unsigned long long* __restrict__ out_ptrs) {
    const int img = blockIdx.y;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= out_rows) return;
    const int lane = threadIdx.x & 31;
    auto kf_of = [&](int c) {
        const double x = __dadd_rn(__dadd_rn(__dmul_rn(m00, (double)c), __dmul_rn(m01, (double)r)), m02);
        return __double2loint(__dadd_rd(x, kMagicD));
    };
    // first column with kf >= bound (kf is non-decreasing in c because m00 > 0): estimate, then fix up exactly
    auto first_ge = [&](int bound) {
        if (bound == INT_MIN) return 0;
        if (bound == INT_MAX) return out_cols;
        double est = floor(((double)bound - m01 * (double)r - m02) / m00);
        int c = (int)fmin(fmax(est, -1.0), (double)out_cols);
        c = max(0, min(out_cols, c));
        while (c > 0 && kf_of(c - 1) >= bound) --c;
        while (c < out_cols && kf_of(c) < bound) ++c;
        return c;
    };
    const int ca = first_ge(own_lo), cb = first_ge(own_hi);
    if (ca >= cb) return;
    const uint16_t* src = local + (int64_t)img * local_stride + (int64_t)r * out_cols;
    uint16_t* dst = reinterpret_cast<uint16_t*>(out_ptrs[img]) + (int64_t)r * out_cols;
    if (dst == src) return;                                          // the owner's own pixels are already in place
    const int64_t e0 = (int64_t)r * out_cols;
    int a8 = ca + (int)((8 - ((e0 + ca) & 7)) & 7);                  // first 16-byte aligned column
    if ((out_ptrs[img] & 15) != 0) a8 = cb;                          // remote image not 16-byte aligned: 2-byte copies
    a8 = min(a8, cb);
    const int b8 = a8 + ((cb - a8) & ~7);
    for (int c = ca + lane; c < a8; c += 32) dst[c] = src[c];
    for (int c = a8 + lane * 8; c < b8; c += 256)
        *reinterpret_cast<uint4*>(dst + c) = *reinterpret_cast<const uint4*>(src + c);
    for (int c = b8 + lane; c < cb; c += 32) dst[c] = src[c];
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

typedef CUresult (*WarpEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int warp_encode_fn(WarpEncodeTiledFn* fn) {
    static WarpEncodeTiledFn cached = nullptr;
    if (!cached) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        SHG_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        SHG_REQUIRE(p && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
        cached = reinterpret_cast<WarpEncodeTiledFn>(p);
    }
    *fn = cached;
    return 0;
}

extern "C" int shg_warp_rows_tma_ok(const uint16_t* d_disk, int64_t disk_stride, int ih, const uint16_t* d_out,
                                    int64_t out_stride, const uint64_t* d_out_ptrs) {
    // TMA: 16-byte aligned base and strides; vector stores: 16-byte aligned images (peer images are cudaMalloc'ed)
    if (getenv("SHG_WARP_OLD")) return 0;
    return (ih % 8 == 0) && (disk_stride % 8 == 0) && (((uintptr_t)d_disk) % 16 == 0) &&
           (d_out_ptrs || ((((uintptr_t)d_out) % 16 == 0) && (out_stride % 8 == 0)));
}

extern "C" int shg_warp_rows_tma(const uint16_t* d_disk, int64_t disk_stride, int n_disk_images, int64_t n_local_frames,
                                 int64_t frame_origin, const int32_t* d_sel, int n_imgs, int64_t n_frames, int ih,
                                 int flip, double m00, double m01, double m02, const uint32_t* d_minmax, uint16_t* d_out,
                                 int64_t out_stride, int out_rows, int out_cols, const uint32_t* d_cval, int own_lo,
                                 int own_hi, const uint64_t* d_out_ptrs, void* stream) {
    SHG_REQUIRE(n_frames > 0 && ih > 0 && out_rows > 0 && out_cols > 0 && n_imgs > 0 && n_disk_images > 0 &&
                n_local_frames > 0, "shg_warp_rows_tma: bad geometry");
    SHG_REQUIRE(own_lo < own_hi, "shg_warp_rows_tma: empty frame range [%d, %d)", own_lo, own_hi);
    SHG_REQUIRE(d_out || d_out_ptrs, "shg_warp_rows_tma: no output");
    SHG_REQUIRE(std::isfinite(m00) && std::isfinite(m01) && std::isfinite(m02) && m00 > 0.0,
                "shg_warp_rows_tma: bad matrix (%g, %g, %g)", m00, m01, m02);
    SHG_REQUIRE(n_imgs <= 65535, "shg_warp_rows_tma: too many images");
    SHG_REQUIRE(shg_warp_rows_tma_ok(d_disk, disk_stride, ih, d_out, out_stride, d_out_ptrs) || getenv("SHG_WARP_OLD"),
                "shg_warp_rows_tma: buffers are not 16-byte aligned / ih is not a multiple of 8");
    const bool sharded = own_lo != INT_MIN || own_hi != INT_MAX;
    SHG_REQUIRE(sharded || (frame_origin == 0 && n_local_frames == n_frames),
                "shg_warp_rows_tma: a complete image must be described with frame_origin 0");
    SHG_REQUIRE(d_cval || !sharded, "shg_warp_rows_tma: a frame-sharded call needs d_cval");
    // tile = 64 slit rows x COLS columns; a stage holds its frame span (+ 7 columns of row phase, + 1 tap, + 1 box because
    // a flipped box does not start on a box boundary); at least two stages must fit beside each other
    auto span_for = [&](int cols) { return m00 * (cols + 7 - 1) + std::fabs(m01) * (kRows - 1) + 5.0; };
    int dev = 0, optin = 0, sms = SHG_SM_COUNT_B200;
    SHG_CHECK(cudaGetDevice(&dev));
    SHG_CHECK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SHG_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t budget = (size_t)optin - 2048;
    // a flipped box does not start on a box boundary: one box of slack
    auto stage_bytes_for = [&](int c) {
        return ((size_t)std::ceil(span_for(c) / kBoxK) + (flip ? 1 : 0)) * kBoxK * kRows * 2;
    };
    // Launch shape (SHG_WARP_SHAPE=<cols>,<threads>,<stages>,<ctas per SM; 0 = one tile per CTA> overrides; tuning):
    // default = one 64-column tile per 256-thread CTA, as many CTAs per SM as shared memory allows
    int cols = 64, nt = 256, want_stages = 1, persistent = 0;
    if (const char* e = getenv("SHG_WARP_SHAPE")) sscanf(e, "%d,%d,%d,%d", &cols, &nt, &want_stages, &persistent);
    SHG_REQUIRE((cols == 64 || cols == 128) && (nt == 256 || nt == 512) && cols / (nt / kRows) >= 8 && want_stages >= 1 &&
                want_stages <= kMaxStages, "shg_warp_rows_tma: bad SHG_WARP_SHAPE");
    if ((size_t)want_stages * stage_bytes_for(cols) + 128 > budget) { cols = 64; nt = 256; want_stages = 1; }
    SHG_REQUIRE(stage_bytes_for(cols) + 128 <= budget, "shg_warp_rows_tma: stretch %g / shear %g too large for the tile",
                m00, m01);
    const int span_boxes = (int)(stage_bytes_for(cols) / (kBoxK * kRows * 2));
    const int n_stages = want_stages;
    const size_t smem = (size_t)n_stages * stage_bytes_for(cols) + 128;
    // column window of this rank over all row tiles (complete images: everything)
    int cb_min = 0, ce_max = out_cols;
    if (sharded) {
        cb_min = out_cols;
        ce_max = 0;
        for (int r0 = 0; r0 < out_rows; r0 += kRows) {
            const int r1 = std::min(r0 + kRows, out_rows);
            const double s0 = m01 * (double)r0, s1 = m01 * (double)(r1 - 1);
            int cb = 0, ce = out_cols;
            if (own_lo != INT_MIN)
                cb = std::max(0, (int)std::max(std::floor(((double)own_lo - std::max(s0, s1) - m02) / m00) - 1.0, -1.0e9));
            if (own_hi != INT_MAX)
                ce = std::min(out_cols, (int)std::min(std::ceil(((double)own_hi - std::min(s0, s1) - m02) / m00) + 1.0, 1.0e9));
            cb_min = std::min(cb_min, cb);
            ce_max = std::max(ce_max, ce);
        }
        if (cb_min >= ce_max) return 0;
    }
    const int col_base = cb_min & ~7;
    WarpEncodeTiledFn encode;
    if (int rc = warp_encode_fn(&encode)) return rc;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    cuuint64_t dims[3] = {(cuuint64_t)ih, (cuuint64_t)n_local_frames, (cuuint64_t)n_disk_images};
    cuuint64_t strides[2] = {(cuuint64_t)ih * 2, (cuuint64_t)disk_stride * 2};
    cuuint32_t box[3] = {(cuuint32_t)kRows, (cuuint32_t)kBoxK, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint16_t*>(d_disk), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SHG_REQUIRE(cr == CUDA_SUCCESS, "shg_warp_rows_tma: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    const int ntx = (ce_max - col_base + 7 + cols - 1) / cols, nry = (out_rows + kRows - 1) / kRows;
    const int64_t total = (int64_t)ntx * nry * n_imgs;
    SHG_REQUIRE(total <= 0x7fffffff, "shg_warp_rows_tma: too many tiles");
    const unsigned grid = persistent > 0 ? (unsigned)std::min<int64_t>(total, (int64_t)sms * persistent) : (unsigned)total;
    cudaStream_t st = as_stream(stream);
    // image [0][0] of image 0 of the map (complete images only): physical frame n-1 when flipped
    const uint16_t* first_px = d_disk + (flip ? (n_frames - 1) : 0) * (int64_t)ih;
#define SHG_WARP_TMA_LAUNCH(C, T, S)                                                                                  \
    do {                                                                                                            \
        SHG_CHECK(cudaFuncSetAttribute(warp_tma_kernel<C, T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        warp_tma_kernel<C, T, S><<<grid, T, smem, st>>>(map, d_sel, n_frames, frame_origin, ih, flip, m00, m01, m02,    \
                                                       d_minmax, d_out, out_stride, out_rows, out_cols, span_boxes, \
                                                       n_stages, d_cval, first_px, disk_stride, own_lo, own_hi,     \
                                                       reinterpret_cast<const unsigned long long*>(d_out_ptrs),     \
                                                       col_base, ce_max, ntx, nry, n_imgs);                         \
    } while (0)
#define SHG_WARP_TMA_PICK(S)                                                                                          \
    do {                                                                                                            \
        if (cols == 128 && nt == 512) SHG_WARP_TMA_LAUNCH(128, 512, S);                                              \
        else if (cols == 128) SHG_WARP_TMA_LAUNCH(128, 256, S);                                                      \
        else if (nt == 512) SHG_WARP_TMA_LAUNCH(64, 512, S);                                                         \
        else SHG_WARP_TMA_LAUNCH(64, 256, S);                                                                        \
    } while (0)
    if (sharded) SHG_WARP_TMA_PICK(true);
    else SHG_WARP_TMA_PICK(false);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_exchange_rows(const uint16_t* d_local, int64_t local_stride, int n_imgs, int out_rows, int out_cols,
                                 double m00, double m01, double m02, int own_lo, int own_hi, const uint64_t* d_out_ptrs,
                                 void* stream) {
    SHG_REQUIRE(d_local && d_out_ptrs && n_imgs > 0 && out_rows > 0 && out_cols > 0 && m00 > 0.0,
                "shg_exchange_rows: bad arguments");
    SHG_REQUIRE(((uintptr_t)d_local) % 16 == 0 && local_stride % 8 == 0, "shg_exchange_rows: local images not 16-byte aligned");
    SHG_REQUIRE(n_imgs <= 65535, "shg_exchange_rows: too many images");
    exchange_rows_kernel<<<dim3((out_rows + 7) / 8, n_imgs), 256, 0, as_stream(stream)>>>(
        d_local, local_stride, out_rows, out_cols, m00, m01, m02, own_lo, own_hi,
        reinterpret_cast<const unsigned long long*>(d_out_ptrs));
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
