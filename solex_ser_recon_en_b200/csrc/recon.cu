// Pass 2: per-frame reconstruction along the fitted line at every shift.
// Replaces the reference's frame x shift loop (solex_util.py:111-134):
//   ind_l = clip(int(fit[:,0] + shift), 0, iw-2); ind_r = ind_l + 1
//   disk[s][:, k] = img[arange(ih), ind_l]*lw + img[arange(ih), ind_r]*rw   (float64, truncated)
//
// Two kernels:
//  * recon_generic_kernel -- direct global loads, any geometry (also the
//    non-rotated H >= W case); the first-correct path and the fallback for
//    shapes TMA cannot describe (row pitch not a multiple of 16 bytes).
//  * recon_tma_kernel     -- rotated scans.  In raw coordinates the taps of all
//    shifts for slit column x live in a short run of raw rows around the line,
//    so a persistent CTA keeps one tile of TX columns and walks frames: each
//    frame's [band rows] x [TX columns] is staged into shared memory by one TMA
//    box per run of shifts (cp.async.bulk.tensor, mbarrier completion, 3-4
//    stage ring).  TX columns x G shift groups of threads: a thread walks a
//    contiguous slice of the sorted shifts for its column one band row per
//    shift, converting each tap once.  The image of every shift is addressed
//    through a table of base pointers, which may point into peer GPUs' memory
//    (multi-GPU row exchange fused into the store).  HBM traffic = band rows
//    once + outputs once (ncu: 1.005 x algorithmic).
// Bound: HBM.  Algorithmic bytes / frame = ih*nb*bytes_per_px + n_shifts*ih*2.
#include <cuda.h>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kMaxShifts = 512;
constexpr int kMaxRuns = 8;
constexpr int kMaxBoxRows = 256;

struct ShiftTable {
    int n_shifts;
    int n_runs;
    int run_first[kMaxRuns + 1];   // [run] -> first index into sh/slot (sorted by shift)
    int run_rows[kMaxRuns];        // TMA box height of the run
    int run_off[kMaxRuns];         // element offset of the run inside one stage
    int run_unit[kMaxRuns];        // 1 when the run's shifts are consecutive integers
    short sh[kMaxShifts];
    short slot[kMaxShifts];        // output image index of sh[j]
    short seg_end[kMaxShifts];     // end (exclusive) of the maximal run of consecutive integers that contains sh[j]
};

struct TmaMaps {
    CUtensorMap m[kMaxRuns];
};

template <typename T>
__device__ __forceinline__ double px_to_double(T v);
template <>
__device__ __forceinline__ double px_to_double<uint16_t>(uint16_t v) { return u32_to_double(v); }
template <>
__device__ __forceinline__ double px_to_double<uint8_t>(uint8_t v) { return u32_to_double((uint32_t)v << 8); }

// 16-bit store to a global address held as an integer (local HBM or a peer GPU's)
__device__ __forceinline__ void st_global_u16(unsigned long long addr, uint16_t v) {
    asm volatile("st.global.u16 [%0], %1;" ::"l"(addr), "h"(v) : "memory");
}

__device__ __forceinline__ uint16_t lerp_trunc(double L, double R, double lw, double rw) {
    // (L*lw) + (R*rw), each rounded to nearest: no fma contraction
    return (uint16_t)double_floor_to_u32(__dadd_rn(__dmul_rn(L, lw), __dmul_rn(R, rw)));
}

// ---------------------------------------------------------------------------
template <typename T, bool ROT>
__global__ void __launch_bounds__(256)
recon_generic_kernel(const T* __restrict__ frames, int64_t n_frames, int W, int H,
                     const int* __restrict__ fl, const double* __restrict__ lw, const double* __restrict__ rw,
                     const __grid_constant__ ShiftTable tab, const unsigned long long* __restrict__ out_ptrs,
                     int64_t k0_out) {
    const int ih = ROT ? W : H, iw = ROT ? H : W;
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= ih) return;
    const int i = ROT ? (W - 1 - t) : t;          // ROT: thread t is raw column x = t
    const int f = fl[i];
    const double wl = lw[i], wr = rw[i];
    for (int64_t k = blockIdx.y; k < n_frames; k += gridDim.y) {
        const T* fr = frames + k * (int64_t)H * W;
        const int64_t off = (k0_out + k) * ih + i;
        for (int j = 0; j < tab.n_shifts; ++j) {
            const int il = min(max(f + tab.sh[j], 0), iw - 2);
            T a, b;
            if (ROT) {
                a = fr[(int64_t)il * W + t];
                b = fr[(int64_t)(il + 1) * W + t];
            } else {
                a = fr[(int64_t)t * W + il];
                b = fr[(int64_t)t * W + il + 1];
            }
            st_global_u16(out_ptrs[j] + (unsigned long long)(off * 2), lerp_trunc(px_to_double<T>(a), px_to_double<T>(b), wl, wr));
        }
    }
}

// ---------------------------------------------------------------------------
// TMA / mbarrier plumbing (inline PTX; see the Blackwell guide, TMA + mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2),
        "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// TX columns x G shift groups per CTA: thread (col, grp) walks a contiguous
// slice of each run of shifts for its column, so consecutive shifts reuse the
// previous right tap as their left tap (one shared-memory load + one
// conversion per output instead of two).
template <typename T, int TX, int STAGES, int G>
__global__ void __launch_bounds__(TX * G)
recon_tma_kernel(const __grid_constant__ TmaMaps maps, const __grid_constant__ ShiftTable tab,
                 int64_t n_frames, int W, int H, int n_tx, int stage_elems,
                 const int* __restrict__ fl, const double* __restrict__ lw, const double* __restrict__ rw,
                 const int* __restrict__ row0 /* [n_tx][n_runs] */,
                 const unsigned long long* __restrict__ out_ptrs, int64_t k0_out) {
    extern __shared__ unsigned char smem_raw[];
    // TMA destinations want 128-byte alignment; the launch reserves the slack
    T* stage_buf = reinterpret_cast<T*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    __shared__ __align__(8) uint64_t full[STAGES];
    __shared__ unsigned long long s_out[kMaxShifts];      // per-shift output image (local or a peer GPU's)
    for (int j = threadIdx.x; j < tab.n_shifts; j += blockDim.x) s_out[j] = out_ptrs[j];

    const int tid = threadIdx.x;
    const int col = tid % TX, grp = tid / TX;
    // a CTA keeps ONE column tile (tx) and walks frames: the per-column tables are read once
    const int tx = blockIdx.x % n_tx;
    const int64_t first = blockIdx.x / n_tx, stride = gridDim.x / n_tx;
    const uint32_t stage_bytes = (uint32_t)stage_elems * sizeof(T);

    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int64_t k, int stage) {
        mbar_expect_tx(&full[stage], stage_bytes);
        T* dst = stage_buf + (size_t)stage * stage_elems;
        for (int r = 0; r < tab.n_runs; ++r)
            tma_load_3d(dst + tab.run_off[r], &maps.m[r], &full[stage], tx * TX, row0[tx * tab.n_runs + r], (int)k, policy);
    };

    if (tid == 0)
        for (int s = 0; s < STAGES; ++s)
            if (first + s * stride < n_frames) issue(first + s * stride, s);

    const int iw = H;
    const int x = tx * TX + col;
    const bool live = x < W;
    const int i = W - 1 - x;
    int f = 0;
    double wl = 0.0, wr = 0.0;
    if (live) { f = fl[i]; wl = lw[i]; wr = rw[i]; }
    int it = 0;
    for (int64_t k = first; k < n_frames; k += stride, ++it) {
        const int stage = it % STAGES;
        const uint32_t phase = (it / STAGES) & 1;

        mbar_wait(&full[stage], phase);

        if (live) {
            const T* buf = stage_buf + (size_t)stage * stage_elems + col;
            const unsigned long long off2 = (unsigned long long)(((k0_out + k) * W + i) * 2);   // byte offset
            for (int r = 0; r < tab.n_runs; ++r) {
                const T* rb = buf + tab.run_off[r];
                const int base = row0[tx * tab.n_runs + r];
                const int ja = tab.run_first[r], jb = tab.run_first[r + 1];
                const int per = (jb - ja + G - 1) / G;
                const int j0 = ja + grp * per, j1 = min(jb, j0 + per);
                if (j0 >= j1) continue;
                if (f + tab.sh[j0] >= 0 && f + tab.sh[j1 - 1] <= iw - 2) {
                    // nothing clipped: walk down the band one row per shift through each stretch of
                    // consecutive shifts; every tap is loaded and converted once (right tap of shift s =
                    // left tap of s+1)
                    int j = j0;
                    while (j < j1) {
                        const int je = min(j1, (int)tab.seg_end[j]);
                        const T* p = rb + (f + tab.sh[j] - base) * TX;
                        double Lw = __dmul_rn(px_to_double<T>(p[0]), wl);
#pragma unroll 4
                        for (; j < je; ++j) {
                            p += TX;
                            const double R = px_to_double<T>(p[0]);
                            const double v = __dadd_rn(Lw, __dmul_rn(R, wr));
                            Lw = __dmul_rn(R, wl);
                            st_global_u16(s_out[j] + off2, (uint16_t)double_floor_to_u32(v));
                        }
                    }
                } else {
                    int prev = -0x40000000;
                    double R = 0.0;
                    for (int j = j0; j < j1; ++j) {
                        const int il = min(max(f + tab.sh[j], 0), iw - 2) - base;
                        const double L = (il == prev + 1) ? R : px_to_double<T>(rb[il * TX]);
                        R = px_to_double<T>(rb[(il + 1) * TX]);
                        prev = il;
                        st_global_u16(s_out[j] + off2, lerp_trunc(L, R, wl, wr));
                    }
                }
            }
        }
        __syncthreads();                                  // everyone is done with this stage
        if (tid == 0) {
            const int64_t nxt = k + (int64_t)STAGES * stride;
            if (nxt < n_frames) issue(nxt, stage);
        }
    }
}

// Column-pair variant: a thread owns two adjacent columns and one 32-bit global store writes both
// outputs (they are adjacent in the frame-major image): half the store / address instructions per
// output, and 128-byte instead of 64-byte store requests per warp (matters for peer writes over
// NVLink).  It also keeps the per-image minimum of what it writes (see s_min).
__device__ __forceinline__ void st_global_u32(unsigned long long addr, uint32_t v) {
    asm volatile("st.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}

template <typename T, int TX, int STAGES, int G>
__global__ void __launch_bounds__(TX / 2 * G)
recon_tma_pair_kernel(const __grid_constant__ TmaMaps maps, const __grid_constant__ ShiftTable tab,
                      int64_t n_frames, int W, int H, int n_tx, int stage_elems,
                      const int* __restrict__ fl, const double* __restrict__ lw, const double* __restrict__ rw,
                      const int* __restrict__ row0 /* [n_tx][n_runs] */,
                      const unsigned long long* __restrict__ out_ptrs, int64_t k0_out,
                      uint32_t* __restrict__ gmin /* [n_shifts] by slot, or null */) {
    extern __shared__ unsigned char smem_raw[];
    T* stage_buf = reinterpret_cast<T*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    __shared__ __align__(8) uint64_t full[STAGES];
    __shared__ unsigned long long s_out[kMaxShifts];
    for (int j = threadIdx.x; j < tab.n_shifts; j += blockDim.x) s_out[j] = out_ptrs[j];

    constexpr int HT = TX / 2;
    // running minimum of every (shift, column pair) this CTA produces, packed like the stores (the
    // circularisation clips to the image minimum: folding it in here saves re-reading every image)
    uint32_t* s_min = reinterpret_cast<uint32_t*>(stage_buf + (size_t)STAGES * stage_elems);
    const bool do_min = gmin != nullptr;
    if (do_min)
        for (int q = threadIdx.x; q < tab.n_shifts * HT; q += blockDim.x) s_min[q] = 0xFFFFFFFFu;
    const int tid = threadIdx.x;
    const int cp = tid % HT, grp = tid / HT;
    const int tx = blockIdx.x % n_tx;
    const int64_t first = blockIdx.x / n_tx, stride = gridDim.x / n_tx;
    const uint32_t stage_bytes = (uint32_t)stage_elems * sizeof(T);

    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int64_t k, int stage) {
        mbar_expect_tx(&full[stage], stage_bytes);
        T* dst = stage_buf + (size_t)stage * stage_elems;
        for (int r = 0; r < tab.n_runs; ++r)
            tma_load_3d(dst + tab.run_off[r], &maps.m[r], &full[stage], tx * TX, row0[tx * tab.n_runs + r], (int)k, policy);
    };
    if (tid == 0)
        for (int s = 0; s < STAGES; ++s)
            if (first + s * stride < n_frames) issue(first + s * stride, s);

    const int iw = H;
    const int x0 = tx * TX + 2 * cp;                      // columns x0, x0+1 (W is even: both live or neither)
    const bool live = x0 + 1 < W;
    const int i1 = W - 2 - x0;                            // output index of column x0+1; column x0 is i1+1
    int f0 = 0, f1 = 0;
    double wl0 = 0, wr0 = 0, wl1 = 0, wr1 = 0;
    if (live) {
        f0 = fl[i1 + 1]; wl0 = lw[i1 + 1]; wr0 = rw[i1 + 1];
        f1 = fl[i1];     wl1 = lw[i1];     wr1 = rw[i1];
    }
    // two 16-bit shared-memory loads per band row (one per column).  A single 32-bit load + unpack is
    // one instruction MORE, and is only valid when both columns sit on the same band row (f0 == f1), which
    // cost a predicated copy of every load / move in the loop (32 -> 23 instructions per pair).
    auto taps = [&](const T* p0, const T* p1, double& a, double& b) {
        a = px_to_double<T>(p0[0]);
        b = px_to_double<T>(p1[0]);
    };
    int it = 0;
    for (int64_t k = first; k < n_frames; k += stride, ++it) {
        const int stage = it % STAGES;
        const uint32_t phase = (it / STAGES) & 1;
        mbar_wait(&full[stage], phase);
        if (live) {
            const T* buf = stage_buf + (size_t)stage * stage_elems + 2 * cp;
            const unsigned long long off2 = (unsigned long long)(((k0_out + k) * W + i1) * 2);
            for (int r = 0; r < tab.n_runs; ++r) {
                const T* rb = buf + tab.run_off[r];
                const int base = row0[tx * tab.n_runs + r];
                const int ja = tab.run_first[r], jb = tab.run_first[r + 1];
                const int per = (jb - ja + G - 1) / G;
                const int j0 = ja + grp * per, j1 = min(jb, j0 + per);
                if (j0 >= j1) continue;
                const int lo = min(f0, f1) + tab.sh[j0], hi = max(f0, f1) + tab.sh[j1 - 1];
                if (lo >= 0 && hi <= iw - 2) {
                    int j = j0;
                    while (j < j1) {
                        const int je = min(j1, (int)tab.seg_end[j]);
                        const T* p0 = rb + (f0 + tab.sh[j] - base) * TX;
                        const T* p1 = rb + (f1 + tab.sh[j] - base) * TX + 1;
                        double a, b;
                        taps(p0, p1, a, b);
                        double Lw0 = __dmul_rn(a, wl0), Lw1 = __dmul_rn(b, wl1);
#pragma unroll 4
                        for (; j < je; ++j) {
                            p0 += TX;
                            p1 += TX;
                            taps(p0, p1, a, b);
                            const double v0 = __dadd_rn(Lw0, __dmul_rn(a, wr0));
                            const double v1 = __dadd_rn(Lw1, __dmul_rn(b, wr1));
                            Lw0 = __dmul_rn(a, wl0);
                            Lw1 = __dmul_rn(b, wl1);
                            const uint32_t w = (double_floor_to_u32(v1) & 0xffffu) | (double_floor_to_u32(v0) << 16);
                            st_global_u32(s_out[j] + off2, w);
                            if (do_min) {
                                uint32_t* m = s_min + j * HT + cp;
                                *m = __vminu2(*m, w);
                            }
                        }
                    }
                } else {
                    for (int j = j0; j < j1; ++j) {
                        const int il0 = min(max(f0 + tab.sh[j], 0), iw - 2) - base;
                        const int il1 = min(max(f1 + tab.sh[j], 0), iw - 2) - base;
                        const uint32_t q0 = lerp_trunc(px_to_double<T>(rb[il0 * TX]), px_to_double<T>(rb[(il0 + 1) * TX]), wl0, wr0);
                        const uint32_t q1 = lerp_trunc(px_to_double<T>(rb[il1 * TX + 1]), px_to_double<T>(rb[(il1 + 1) * TX + 1]), wl1, wr1);
                        st_global_u32(s_out[j] + off2, q1 | (q0 << 16));
                        if (do_min) {
                            uint32_t* m = s_min + j * HT + cp;
                            *m = __vminu2(*m, q1 | (q0 << 16));
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            const int64_t nxt = k + (int64_t)STAGES * stride;
            if (nxt < n_frames) issue(nxt, stage);
        }
    }
    if (do_min) {                                         // (the loop ends with a barrier: the table is complete)
        const int lane = tid & 31, nw = (HT * G) >> 5;
        for (int j = tid >> 5; j < tab.n_shifts; j += nw) {
            uint32_t v = 0xFFFFFFFFu;
            for (int q = lane; q < HT; q += 32) v = __vminu2(v, s_min[j * HT + q]);
            v = min(v & 0xffffu, v >> 16);
            v = __reduce_min_sync(0xffffffffu, v);
            if (lane == 0) atomicMin(&gmin[tab.slot[j]], v);
        }
    }
}

// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* fn) {
    static EncodeTiledFn cached = nullptr;
    if (!cached) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        SHG_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        SHG_REQUIRE(p && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
        cached = reinterpret_cast<EncodeTiledFn>(p);
    }
    *fn = cached;
    return 0;
}

struct HostPlan {
    ShiftTable tab;
    std::vector<int> fl;
    std::vector<double> lw, rw;
    std::vector<int> row0;     // [n_tx][n_runs]
    int n_tx = 0, tx = 0, stage_elems = 0;
    bool tma_ok = false;
};

// Sort shifts, group them into runs of nearby shifts (one TMA box each).
void build_runs(const int32_t* shifts, int n, int max_rows, ShiftTable& tab, std::vector<std::pair<int, int>>& order) {
    order.resize(n);
    for (int j = 0; j < n; ++j) order[j] = {shifts[j], j};
    std::stable_sort(order.begin(), order.end());
    tab.n_shifts = n;
    for (int j = 0; j < n; ++j) { tab.sh[j] = (short)order[j].first; tab.slot[j] = (short)order[j].second; }
    // greedy: start a new run when the gap to the previous shift is large or the run gets too tall
    std::vector<int> firsts{0};
    for (int j = 1; j < n; ++j) {
        const int gap = order[j].first - order[j - 1].first;
        const int span = order[j].first - order[firsts.back()].first;
        if (gap > 6 || span + 2 > max_rows) firsts.push_back(j);
    }
    // too many runs: merge the closest neighbours until they fit
    while ((int)firsts.size() > kMaxRuns) {
        int best = 1, best_gap = 1 << 30;
        for (size_t r = 1; r < firsts.size(); ++r) {
            const int gap = order[firsts[r]].first - order[firsts[r] - 1].first;
            if (gap < best_gap) { best_gap = gap; best = (int)r; }
        }
        firsts.erase(firsts.begin() + best);
    }
    tab.n_runs = (int)firsts.size();
    for (int r = 0; r < tab.n_runs; ++r) tab.run_first[r] = firsts[r];
    tab.run_first[tab.n_runs] = n;
    for (int r = 0; r < tab.n_runs; ++r) {
        tab.run_unit[r] = 1;
        for (int j = tab.run_first[r] + 1; j < tab.run_first[r + 1]; ++j)
            if (order[j].first != order[j - 1].first + 1) tab.run_unit[r] = 0;
        // stretches of consecutive integers inside the run (a run may have holes, e.g. -50..9, 11..50)
        int j = tab.run_first[r];
        while (j < tab.run_first[r + 1]) {
            int e = j + 1;
            while (e < tab.run_first[r + 1] && order[e].first == order[e - 1].first + 1) ++e;
            for (int q = j; q < e; ++q) tab.seg_end[q] = (short)e;
            j = e;
        }
    }
}

}  // namespace

extern "C" int64_t shg_recon_workspace_bytes(int ih, int n_shifts) {
    // fl + lw + rw + row0 (worst case one tile per 64 columns, kMaxRuns runs), 256-byte aligned pieces
    const int64_t a = ((int64_t)ih * 4 + 255) / 256 * 256;
    const int64_t b = ((int64_t)ih * 8 + 255) / 256 * 256;
    const int64_t c = (((int64_t)(ih + 63) / 64) * kMaxRuns * 4 + 255) / 256 * 256;
    (void)n_shifts;
    return a + 2 * b + c + kMaxShifts * 8;
}

extern "C" int shg_recon(const void* d_frames, int bytes_per_px, int64_t n_frames, int W, int H,
                         const double* h_fit, const int32_t* h_shifts, int n_shifts,
                         uint16_t* d_disk, int64_t shift_stride, const uint64_t* h_out_ptrs, int64_t k0_out,
                         int impl_and_cap, void* d_work, int64_t work_bytes, uint32_t* d_min, int* h_min_done, void* stream) {
    // bits 0-7: kernel variant (0 auto, 1 direct loads, 2 TMA); bits 8-15: use at most that many SMs (0 = all).  A
    // caller that runs latency-critical small kernels beside this one (the limb search of the ellipse fit while the
    // other shifts are reconstructed, on several GPUs) leaves them a few SMs: the persistent CTAs of this kernel
    // otherwise hold every SM until it ends.
    const int impl = impl_and_cap & 0xff, sm_cap = (impl_and_cap >> 8) & 0xff;
    if (h_min_done) *h_min_done = 0;
    SHG_REQUIRE(bytes_per_px == 1 || bytes_per_px == 2, "shg_recon: bytes_per_px must be 1 or 2");
    SHG_REQUIRE(n_shifts >= 1 && n_shifts <= kMaxShifts, "shg_recon: %d shifts (max %d)", n_shifts, kMaxShifts);
    SHG_REQUIRE(W >= 2 && H >= 2, "shg_recon: bad geometry %dx%d", W, H);
    if (n_frames <= 0) return 0;
    const bool rot = W > H;
    const int ih = rot ? W : H, iw = rot ? H : W;
    SHG_REQUIRE(work_bytes >= shg_recon_workspace_bytes(ih, n_shifts), "shg_recon: workspace too small");
    for (int j = 0; j < n_shifts; ++j)
        SHG_REQUIRE(h_shifts[j] > -30000 && h_shifts[j] < 30000, "shg_recon: shift %d out of range", h_shifts[j]);
    cudaStream_t st = as_stream(stream);

    // ---- host tables (reference solex_util.py:113-123) ----------------------
    HostPlan plan;
    plan.fl.resize(ih); plan.lw.resize(ih); plan.rw.resize(ih);
    for (int i = 0; i < ih; ++i) {
        double f0 = h_fit[4 * i + 0];
        f0 = std::min(std::max(f0, -1.0e9), 1.0e9);
        plan.fl[i] = (int)f0;                          // .astype(int): truncation (value is integral)
        plan.lw[i] = 1.0 - h_fit[4 * i + 1];
        plan.rw[i] = 1.0 - plan.lw[i];
    }
    std::vector<std::pair<int, int>> order;
    build_runs(h_shifts, n_shifts, kMaxBoxRows - 8, plan.tab, order);

    // ---- can TMA describe this stack? ---------------------------------------
    const int64_t row_bytes = (int64_t)W * bytes_per_px;
    bool tma = rot && impl != 1 && row_bytes % 16 == 0 && ((int64_t)H * row_bytes) % 16 == 0 &&
               ((uintptr_t)d_frames % 16 == 0) && n_frames < (1LL << 31);
    int TX = 0, stages = 0, G = 2;
    if (tma) {
        // tile width / shift groups: TX*G threads per CTA.  Default 256 columns x 2 groups (one 512-thread CTA
        // per SM with 3 stages at config-5 band heights: 5.5 ms vs 5.9 ms for 128 x 4 on a B200);
        // SHG_RECON_TX / _G override for tuning.
        int want_tx = 256;
        if (const char* e = getenv("SHG_RECON_TX")) want_tx = atoi(e);
        if (const char* e = getenv("SHG_RECON_G")) G = atoi(e);
        if (G != 1 && G != 2 && G != 4 && G != 8) G = 2;
        int cands[3] = {want_tx, 128, 64};
        for (int cand : cands) {
            if (cand != 64 && cand != 128 && cand != 256) continue;
            if (cand * G > 1024) continue;
            if (cand > 64 && cand / 2 >= W) continue;
            const int n_tx = (W + cand - 1) / cand;
            plan.row0.assign((size_t)n_tx * plan.tab.n_runs, 0);
            bool ok = true;
            int elems = 0;
            for (int r = 0; r < plan.tab.n_runs && ok; ++r) {
                int rows_needed = 0;
                for (int t = 0; t < n_tx; ++t) {
                    int lo = 1 << 30, hi = -(1 << 30);
                    for (int x = t * cand; x < std::min(W, (t + 1) * cand); ++x) {
                        const int f = plan.fl[W - 1 - x];
                        const int a = std::min(std::max(f + plan.tab.sh[plan.tab.run_first[r]], 0), iw - 2);
                        const int b = std::min(std::max(f + plan.tab.sh[plan.tab.run_first[r + 1] - 1], 0), iw - 2);
                        lo = std::min(lo, a);
                        hi = std::max(hi, b + 1);
                    }
                    plan.row0[(size_t)t * plan.tab.n_runs + r] = lo;
                    rows_needed = std::max(rows_needed, hi - lo + 1);
                }
                if (rows_needed > kMaxBoxRows) ok = false;
                plan.tab.run_rows[r] = rows_needed;
                plan.tab.run_off[r] = elems;
                elems += ((rows_needed * cand * bytes_per_px + 127) / 128 * 128) / bytes_per_px;
            }
            if (!ok) continue;
            const int64_t sbytes = (int64_t)elems * bytes_per_px;
            const int max_stages = (int)(200 * 1024 / sbytes);
            if (max_stages >= 3) {
                TX = cand; plan.n_tx = n_tx; plan.stage_elems = elems;
                // prefer two CTAs per SM when four stages of this tile fit in half the shared memory
                stages = (4 * sbytes + 2048 <= 110 * 1024) ? 4 : std::min(max_stages, 4);
                break;
            }
        }
        if (TX == 0) tma = false;
    }
    SHG_REQUIRE(!(impl == 2 && !tma), "shg_recon: TMA path requested but this geometry cannot use it");

    // ---- upload tables --------------------------------------------------------
    char* w = static_cast<char*>(d_work);
    const int64_t a = ((int64_t)ih * 4 + 255) / 256 * 256;
    const int64_t b = ((int64_t)ih * 8 + 255) / 256 * 256;
    int* d_fl = reinterpret_cast<int*>(w);
    double* d_lw = reinterpret_cast<double*>(w + a);
    double* d_rw = reinterpret_cast<double*>(w + a + b);
    int* d_row0 = reinterpret_cast<int*>(w + a + 2 * b);
    const int64_t c_bytes = (((int64_t)(ih + 63) / 64) * kMaxRuns * 4 + 255) / 256 * 256;
    unsigned long long* d_optr = reinterpret_cast<unsigned long long*>(w + a + 2 * b + c_bytes);
    // output image of each shift, in the kernels' sorted shift order
    std::vector<unsigned long long> optr(n_shifts);
    for (int j = 0; j < n_shifts; ++j) {
        const int slot = plan.tab.slot[j];
        optr[j] = h_out_ptrs ? (unsigned long long)h_out_ptrs[slot]
                             : (unsigned long long)(uintptr_t)(d_disk + (int64_t)slot * shift_stride);
        SHG_REQUIRE(optr[j] != 0 && optr[j] % 2 == 0, "shg_recon: bad output pointer for shift slot %d", slot);
    }
    SHG_CHECK(cudaMemcpyAsync(d_optr, optr.data(), (size_t)n_shifts * 8, cudaMemcpyHostToDevice, st));
    SHG_CHECK(cudaMemcpyAsync(d_fl, plan.fl.data(), (size_t)ih * 4, cudaMemcpyHostToDevice, st));
    SHG_CHECK(cudaMemcpyAsync(d_lw, plan.lw.data(), (size_t)ih * 8, cudaMemcpyHostToDevice, st));
    SHG_CHECK(cudaMemcpyAsync(d_rw, plan.rw.data(), (size_t)ih * 8, cudaMemcpyHostToDevice, st));

    if (!tma) {
        dim3 grid((ih + 255) / 256, (unsigned)std::min<int64_t>(n_frames, 65535));
        if (rot && bytes_per_px == 2)
            recon_generic_kernel<uint16_t, true><<<grid, 256, 0, st>>>((const uint16_t*)d_frames, n_frames, W, H, d_fl, d_lw, d_rw, plan.tab, d_optr, k0_out);
        else if (rot)
            recon_generic_kernel<uint8_t, true><<<grid, 256, 0, st>>>((const uint8_t*)d_frames, n_frames, W, H, d_fl, d_lw, d_rw, plan.tab, d_optr, k0_out);
        else if (bytes_per_px == 2)
            recon_generic_kernel<uint16_t, false><<<grid, 256, 0, st>>>((const uint16_t*)d_frames, n_frames, W, H, d_fl, d_lw, d_rw, plan.tab, d_optr, k0_out);
        else
            recon_generic_kernel<uint8_t, false><<<grid, 256, 0, st>>>((const uint8_t*)d_frames, n_frames, W, H, d_fl, d_lw, d_rw, plan.tab, d_optr, k0_out);
        SHG_LAUNCH_CHECK();
        return 0;
    }

    SHG_CHECK(cudaMemcpyAsync(d_row0, plan.row0.data(), plan.row0.size() * 4, cudaMemcpyHostToDevice, st));

    EncodeTiledFn encode;
    if (int rc = get_encode_fn(&encode)) return rc;
    TmaMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int r = 0; r < plan.tab.n_runs; ++r) {
        cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_frames};
        cuuint64_t strides[2] = {(cuuint64_t)row_bytes, (cuuint64_t)H * row_bytes};
        cuuint32_t box[3] = {(cuuint32_t)TX, (cuuint32_t)plan.tab.run_rows[r], 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult cr = encode(&maps.m[r], bytes_per_px == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8,
                             3, const_cast<void*>(d_frames), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SHG_REQUIRE(cr == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for box %dx%d", (int)cr, TX, plan.tab.run_rows[r]);
    }

    int dev = 0, sms = SHG_SM_COUNT_B200;
    SHG_CHECK(cudaGetDevice(&dev));
    SHG_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sm_cap > 0) sms = std::max(plan.n_tx, std::min(sms, sm_cap));
    size_t smem = (size_t)plan.stage_elems * bytes_per_px * stages + 128;
    int pair = (W % 2 == 0) ? 1 : 0;                     // column-pair kernel (default); SHG_RECON_PAIR=0 selects the other
    if (const char* e = getenv("SHG_RECON_PAIR")) pair = pair && atoi(e) != 0;
    bool all_ptrs_even4 = true;                          // 32-bit stores need 4-byte aligned image bases
    for (int j = 0; j < n_shifts; ++j) all_ptrs_even4 = all_ptrs_even4 && (optr[j] % 4 == 0);
    const bool use_pair = pair && all_ptrs_even4 && (TX == 256 || TX == 128);
    // per-image minimum folded into the pair kernel when its (shift x column pair) table fits beside the stages
    uint32_t* gmin = nullptr;
    if (use_pair && d_min && !getenv("SHG_RECON_NO_MIN")) {
        const size_t min_bytes = (size_t)n_shifts * (TX / 2) * 4;
        if (smem + min_bytes + 6 * 1024 <= 227 * 1024) {
            gmin = d_min;
            smem += min_bytes;
            if (h_min_done) *h_min_done = 1;
        }
    }
    const int64_t n_tiles = n_frames * plan.n_tx;
    const int by_smem = std::max<int>(1, (int)std::min<size_t>(8, (220 * 1024) / (smem + 1024)));
    const int by_threads = std::max(1, 2048 / (TX * G));
    const int ctas_per_sm = std::min(by_smem, by_threads);
    // grid = n_tx column tiles x frame lanes (a multiple of n_tx, close to one full wave of resident CTAs)
    const int64_t lanes = std::max<int64_t>(1, std::min<int64_t>(n_frames, ((int64_t)sms * ctas_per_sm) / plan.n_tx));
    const unsigned grid = (unsigned)(lanes * plan.n_tx);
    (void)n_tiles;

#define SHG_LAUNCH_TMA(T, TXV, ST, GV)                                                                     \
    do {                                                                                                   \
        auto kern = recon_tma_kernel<T, TXV, ST, GV>;                                                      \
        SHG_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        kern<<<grid, TXV * GV, smem, st>>>(maps, plan.tab, n_frames, W, H, plan.n_tx, plan.stage_elems, d_fl, \
                                           d_lw, d_rw, d_row0, d_optr, k0_out);                            \
    } while (0)
#define SHG_DISPATCH_G(T, TXV, ST)                                           \
    do {                                                                     \
        if (G == 1) SHG_LAUNCH_TMA(T, TXV, ST, 1);                           \
        else if (G == 2) SHG_LAUNCH_TMA(T, TXV, ST, 2);                      \
        else if (G == 8 && TXV <= 128) SHG_LAUNCH_TMA(T, TXV, ST, (TXV <= 128 ? 8 : 4)); \
        else SHG_LAUNCH_TMA(T, TXV, ST, 4);                                  \
    } while (0)
#define SHG_DISPATCH_ST(T, TXV)                                  \
    do {                                                         \
        if (stages >= 4) SHG_DISPATCH_G(T, TXV, 4);              \
        else SHG_DISPATCH_G(T, TXV, 3);                          \
    } while (0)
#define SHG_DISPATCH_TX(T)                                       \
    do {                                                         \
        if (TX == 256) SHG_DISPATCH_ST(T, 256);                  \
        else if (TX == 128) SHG_DISPATCH_ST(T, 128);             \
        else SHG_DISPATCH_ST(T, 64);                             \
    } while (0)
    if (use_pair) {
        const int GP = (G == 2 || G == 4 || G == 8) ? (getenv("SHG_RECON_G") ? G : 4) : 4;
#define SHG_LAUNCH_PAIR(T, TXV, ST, GV)                                                                    \
    do {                                                                                                   \
        auto kern = recon_tma_pair_kernel<T, TXV, ST, GV>;                                                 \
        SHG_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        kern<<<grid, TXV / 2 * GV, smem, st>>>(maps, plan.tab, n_frames, W, H, plan.n_tx, plan.stage_elems, \
                                               d_fl, d_lw, d_rw, d_row0, d_optr, k0_out, gmin);            \
    } while (0)
#define SHG_PAIR_G(T, TXV, ST)                                   \
    do {                                                         \
        if (GP == 2) SHG_LAUNCH_PAIR(T, TXV, ST, 2);             \
        else if (GP == 8) SHG_LAUNCH_PAIR(T, TXV, ST, 8);        \
        else SHG_LAUNCH_PAIR(T, TXV, ST, 4);                     \
    } while (0)
#define SHG_PAIR_ST(T, TXV)                                      \
    do {                                                         \
        if (stages >= 4) SHG_PAIR_G(T, TXV, 4);                  \
        else SHG_PAIR_G(T, TXV, 3);                              \
    } while (0)
#define SHG_PAIR_TX(T)                                           \
    do {                                                         \
        if (TX == 256) SHG_PAIR_ST(T, 256);                      \
        else SHG_PAIR_ST(T, 128);                                \
    } while (0)
        if (bytes_per_px == 2) SHG_PAIR_TX(uint16_t);
        else SHG_PAIR_TX(uint8_t);
        SHG_LAUNCH_CHECK();
        return 0;
    }
    if (bytes_per_px == 2) SHG_DISPATCH_TX(uint16_t);
    else SHG_DISPATCH_TX(uint8_t);
    SHG_LAUNCH_CHECK();
    return 0;
}
