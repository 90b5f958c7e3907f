// Transversalium correction (reference solex_util.py:76-86, 383-395, 489-516).
//
// Stage 1 (shg_transv_row_stats): for every image row y inside the disk, over
// the chord [xa, xb):   rat = log(img[y] / img[y-1])
//                       out = mean(rat[|rat - median(rat)| / MAD < 2])
// One CTA per (row, image); the two medians are EXACT order statistics.  Two kernels:
//  * transv_row_stats_reg_kernel (further down) takes every ordinary row: 16 ratios per thread in registers
//    (the rest of a long chord in shared memory), a 2048-bin histogram, window selects verified by counts
//    taken from the data, the MAD window derived from the median's histogram; rows it cannot take go to a
//    todo list;
//  * transv_row_stats_kernel (the classic kernel: every row when the other is switched off, else a small
//    persistent grid over the todo list) keeps rat in shared memory (or an L2-resident scratch row for very
//    long chords) and selects with
//      hist_select  -- first choice, rows without zeros: a one-level counting
//        select (1024 bins from sample quartiles, shared-memory atomics, block scan,
//        exact fp64 ranking of the few elements in the target bins); the median's
//        counting pass is fused into the loop that computes the ratios;
//      fast_select  -- fall-back for chords of up to 32 elements per thread: binary
//        radix select over monotone 32-bit keys held bit-sliced in registers;
//      block_select -- longer chords / very many ties: fixed-point 5 bits per level.
// Pixels are uint16 and neighbouring rows differ by noise, so log(a/b) is computed
// (no table gather): 2 atanh((a-b)/(a+b)) by series for |z| <= 2^-6, else
// L(a) - L(b) with L = log_u16; both are good to <= ~2e-15 absolute on a quantity
// of ~1e-2 (the reference's own log(a/b) carries ~1e-16 from the quotient).
// Stage 2 (shg_row_scale_u16): out = trunc(min(img * gain[row], 65535)).
// Bound: stage 1 is instruction-bound (4.0e9 warp instructions for the 101 images of config 5, ~150 per chord
// element, in either kernel; the register-resident one issues them at 68 % of the slots with 8 rows per SM in
// flight against 58 % with 6; its HBM time is 0.5 ms of 4.9 ms); stage 2 is HBM-bound (each image read once,
// written once).
#include <stdlib.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace {

constexpr int kListCap = 256;            // upper bound; a CTA of T threads ranks at most min(T, 256) candidates
constexpr double kQ = 1073741824.0;     // 2^30

struct Shared {
    unsigned int whist[32][33];
    unsigned int hist[32];
    double red[2][32];
    double list[kListCap];
    unsigned long long cnt[4];
    int list_n;
    int ibc[8];
    double dbc[4];
};

constexpr int kHistBins = 1024;           // hist_select: one level, lives in Shared::whist (32*33 words)

__device__ __forceinline__ double warp_min(double v) {
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// min and max over the block; result broadcast to every thread
__device__ void block_minmax(double& mn, double& mx, Shared& S) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    mn = warp_min(mn);
    mx = warp_max(mx);
    __syncthreads();
    if (lane == 0) { S.red[0][warp] = mn; S.red[1][warp] = mx; }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    mn = warp_min(lane < nw ? S.red[0][lane] : INFINITY);
    mx = warp_max(lane < nw ? S.red[1][lane] : -INFINITY);
}

__device__ void block_sum2(double& a, double& b, Shared& S) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();
    if (lane == 0) { S.red[0][warp] = a; S.red[1][warp] = b; }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    a = warp_sum(lane < nw ? S.red[0][lane] : 0.0);
    b = warp_sum(lane < nw ? S.red[1][lane] : 0.0);
}

template <int KIND>
__device__ __forceinline__ double key_of(const double* vals, int i, double med) {
    const double v = vals[i];
    return KIND == 0 ? v : fabs(v - med);
}

__device__ __forceinline__ uint32_t qkey(double x, double lo, double scale) {
    double u = (x - lo) * scale;                       // monotone in x
    u = fmin(fmax(u, 0.0), kQ - 1.0);
    return double_floor_to_u32(u);
}

// Exact order statistics t (and t+1 when need2) of the keys that lie in the
// finite range [lo, hi]; m = number of such keys; 0 <= t (< t+1) < m.
// Results in S.dbc[0], S.dbc[1].
template <int KIND, int kT>
__device__ __noinline__ void block_select(const double* vals, int n, double med, double lo, double hi, int m, int t, bool need2,
                             Shared& S) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        if (!(hi > lo)) {                                    // all candidates equal
            if (threadIdx.x == 0) { S.dbc[0] = lo; S.dbc[1] = lo; }
            __syncthreads();
            return;
        }
        double scale = kQ / (hi - lo);
        if (!(scale < 1e300)) scale = 1e300;
        uint32_t prefix = 0;
        int level = 0;
        constexpr int kCap = kT < kListCap ? kT : kListCap;
        for (; level < 6 && m > kCap; ++level) {
            const int pshift = 30 - 5 * level, bshift = 25 - 5 * level;
            unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;       // 32 bins x 8-bit fields, flushed into acc[]
            // per-thread counts can exceed 255 for very long rows: flush in chunks of 255 elements
            unsigned int acc[32];
#pragma unroll
            for (int b = 0; b < 32; ++b) acc[b] = 0;
            int since = 0;
            auto flush = [&]() {
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    acc[b] += (unsigned int)((c0 >> (8 * b)) & 0xff);
                    acc[8 + b] += (unsigned int)((c1 >> (8 * b)) & 0xff);
                    acc[16 + b] += (unsigned int)((c2 >> (8 * b)) & 0xff);
                    acc[24 + b] += (unsigned int)((c3 >> (8 * b)) & 0xff);
                }
                c0 = c1 = c2 = c3 = 0;
                since = 0;
            };
            for (int i = threadIdx.x; i < n; i += kT) {
                const double x = key_of<KIND>(vals, i, med);
                if (x >= lo && x <= hi) {
                    const uint32_t q = qkey(x, lo, scale);
                    if ((q >> pshift) == prefix) {
                        const uint32_t b = (q >> bshift) & 31u;
                        const unsigned long long inc = 1ull << (8 * (b & 7u));
                        const uint32_t g = b >> 3;
                        c0 += g == 0 ? inc : 0ull;
                        c1 += g == 1 ? inc : 0ull;
                        c2 += g == 2 ? inc : 0ull;
                        c3 += g == 3 ? inc : 0ull;
                    }
                }
                if (++since == 255) flush();
            }
            flush();
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const unsigned int tot = __reduce_add_sync(0xffffffffu, acc[b]);
                if (lane == b) S.whist[warp][b] = tot;
            }
            __syncthreads();
            if (warp == 0) {
                unsigned int h = 0;
                for (int w = 0; w < kT / 32; ++w) h += S.whist[w][lane];
                // inclusive prefix over bins
                unsigned int inc = h;
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                const unsigned int exc = inc - h;
                if (h > 0 && (unsigned)t >= exc && (unsigned)t < inc) { S.ibc[0] = lane; S.ibc[1] = (int)exc; S.ibc[2] = (int)h; }
                if (need2 && h > 0 && (unsigned)(t + 1) >= exc && (unsigned)(t + 1) < inc) S.ibc[3] = lane;
            }
            __syncthreads();
            const int b0 = S.ibc[0], below = S.ibc[1], cnt = S.ibc[2];
            const int b1 = need2 ? S.ibc[3] : b0;
            __syncthreads();
            if (b1 != b0) {
                // ranks t and t+1 straddle two bins: max of bin b0, min of bin b1
                double v0 = -INFINITY, v1 = INFINITY;
                for (int i = threadIdx.x; i < n; i += kT) {
                    const double x = key_of<KIND>(vals, i, med);
                    if (x >= lo && x <= hi) {
                        const uint32_t q = qkey(x, lo, scale);
                        if ((q >> pshift) == prefix) {
                            const int b = (int)((q >> bshift) & 31u);
                            if (b == b0) v0 = fmax(v0, x);
                            if (b == b1) v1 = fmin(v1, x);
                        }
                    }
                }
                block_minmax(v1, v0, S);
                if (threadIdx.x == 0) { S.dbc[0] = v0; S.dbc[1] = v1; }
                __syncthreads();
                return;
            }
            t -= below;
            m = cnt;
            prefix = (prefix << 5) | (uint32_t)b0;
        }
        const int pshift = 30 - 5 * level;
        if (m <= kCap) {
            if (threadIdx.x == 0) S.list_n = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += kT) {
                const double x = key_of<KIND>(vals, i, med);
                if (x >= lo && x <= hi) {
                    const uint32_t q = qkey(x, lo, scale);
                    if (level == 0 || (q >> pshift) == prefix) {
                        const int slot = atomicAdd(&S.list_n, 1);
                        if (slot < kCap) S.list[slot] = x;
                    }
                }
            }
            __syncthreads();
            const int mm = min(S.list_n, kCap);
            if ((int)threadIdx.x < mm) {
                const double c = S.list[threadIdx.x];
                int rank = 0;
                for (int i = 0; i < mm; ++i) {
                    const double o = S.list[i];
                    rank += (o < c || (o == c && i < (int)threadIdx.x)) ? 1 : 0;
                }
                if (rank == t) S.dbc[0] = c;
                if (rank == t + 1) S.dbc[1] = c;
            }
            __syncthreads();
            if (!need2 && threadIdx.x == 0) S.dbc[1] = S.dbc[0];
            __syncthreads();
            return;
        }
        // 30 bits used up and still many candidates in one cell: rebase on their true range
        double cmin = INFINITY, cmax = -INFINITY;
        for (int i = threadIdx.x; i < n; i += kT) {
            const double x = key_of<KIND>(vals, i, med);
            if (x >= lo && x <= hi) {
                const uint32_t q = qkey(x, lo, scale);
                if (q == prefix) { cmin = fmin(cmin, x); cmax = fmax(cmax, x); }
            }
        }
        block_minmax(cmin, cmax, S);
        lo = cmin;
        hi = cmax;
    }
}

// ---------------------------------------------------------------------------
// Fast path for chords of at most 32 elements per thread (n <= 32*kT): the
// order statistic is found by a binary MSB-first radix select on monotone
// 32-bit keys (the float32 rounding of the fp64 key, order-preserving bit
// transform) held BIT-SLICED in registers: after a 32x32 bit transpose,
// slice[b] holds bit b of each of the thread's 32 keys, so one level is two
// logic ops + a popcount per thread and one block-wide sum.  float32 rounding is
// monotone, so keys strictly below the target key are strictly below in fp64;
// the (usually single) elements that share the final key are ranked exactly in
// fp64.  Ranks t and t+1 (even n) share the descent until they part ways.
// w[e] (e < 16) carries the 16-bit keys of elements e (low half) and e+16 (high
// half); afterwards w[b] is the 32-element slice of key bit b (bit j = element j).
__device__ __forceinline__ void transpose16x2(uint32_t (&w)[16]) {
    uint32_t m = 0x00FF00FFu;
#pragma unroll
    for (int j = 8; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
        for (int k = 0; k < 16; k = (k + j + 1) & ~j) {
            const uint32_t t = ((w[k] >> j) ^ w[k + j]) & m;
            w[k] ^= t << j;
            w[k + j] ^= t;
        }
    }
}

__device__ __forceinline__ uint32_t ordered_key32(double x) {
    const uint32_t b = __float_as_uint(__double2float_rn(x));
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// Block-wide integer sum with ONE barrier: the per-warp partials alternate
// between two halves of S.whist (parity = call counter), so a warp can only
// overwrite a half after every thread has passed the barrier of the call in
// between, i.e. after all reads of that half.
template <int kT>
__device__ __forceinline__ int block_sum_int(int v, int parity, Shared& S) {
    constexpr int NW = kT / 32;
    v = __reduce_add_sync(0xffffffffu, v);
    unsigned int* slot = &S.whist[parity & 1][0];
    if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = (unsigned int)v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) t += (int)slot[w];
    return t;
}

// Returns true when handled (results in S.dbc[0..1]); false -> caller falls back to block_select.
// Two rounds of 16 binary levels: the high half of the 32-bit key first (usually enough: the few
// elements that share it are ranked exactly in fp64), the low half only when many elements tie.
// A thread owns up to 32*WORDS elements (element e of word q is vals[(32q + e)*kT + tid]).
template <int KIND, int kT, int WORDS>
__device__ __forceinline__ bool fast_select(const double* vals, int n, double med, int t, bool need2, Shared& S) {
    uint32_t cand[WORDS];
    uint32_t w[WORDS][16];
    int cnt = 0;
    // one pass builds the candidate masks (finite keys; +-inf are ranked by the caller) and the
    // bit slices of the HIGH key halves
#pragma unroll
    for (int q = 0; q < WORDS; ++q) {
        cand[q] = 0;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int i0 = (32 * q + e) * kT + (int)threadIdx.x, i1 = (32 * q + e + 16) * kT + (int)threadIdx.x;
            uint32_t lo = 0, hi = 0;
            if (i0 < n) {
                const double x = key_of<KIND>(vals, i0, med);
                if (fabs(x) < INFINITY) { lo = ordered_key32(x); cand[q] |= 1u << e; }
            }
            if (i1 < n) {
                const double x = key_of<KIND>(vals, i1, med);
                if (fabs(x) < INFINITY) { hi = ordered_key32(x); cand[q] |= 1u << (e + 16); }
            }
            w[q][e] = (lo >> 16) | (hi & 0xFFFF0000u);
        }
        transpose16x2(w[q]);
        cnt += __popc(cand[q]);
    }
    int m = block_sum_int<kT>(cnt, 0, S);
    int parity = 1;
    for (int round = 0; round < 2 && m > kListCap / 2; ++round) {
        if (round == 1) {                                      // many ties in the high half: slices of the low halves
#pragma unroll
            for (int q = 0; q < WORDS; ++q) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    uint32_t lo = 0, hi = 0;
                    const int i0 = (32 * q + e) * kT + (int)threadIdx.x, i1 = (32 * q + e + 16) * kT + (int)threadIdx.x;
                    if ((cand[q] >> e) & 1u) lo = ordered_key32(key_of<KIND>(vals, i0, med));
                    if ((cand[q] >> (e + 16)) & 1u) hi = ordered_key32(key_of<KIND>(vals, i1, med));
                    w[q][e] = (lo & 0xFFFFu) | (hi << 16);
                }
                transpose16x2(w[q]);
            }
        }
#pragma unroll
        for (int L = 0; L < 16; ++L) {
            int z = 0;
#pragma unroll
            for (int q = 0; q < WORDS; ++q) z += __popc(cand[q] & ~w[q][15 - L]);
            const int Z = block_sum_int<kT>(z, parity++, S);
            if (need2 && t == Z - 1) {                         // rank t is the largest "0", rank t+1 the smallest "1"
                double v0 = -INFINITY, v1 = INFINITY;
#pragma unroll
                for (int q = 0; q < WORDS; ++q) {
                    const uint32_t lo_set = cand[q] & ~w[q][15 - L], hi_set = cand[q] & w[q][15 - L];
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int i = (32 * q + e) * kT + (int)threadIdx.x;
                        if ((lo_set >> e) & 1u) v0 = fmax(v0, key_of<KIND>(vals, i, med));
                        if ((hi_set >> e) & 1u) v1 = fmin(v1, key_of<KIND>(vals, i, med));
                    }
                }
                block_minmax(v1, v0, S);
                if (threadIdx.x == 0) { S.dbc[0] = v0; S.dbc[1] = v1; }
                __syncthreads();
                return true;
            }
            if (t < Z) {
#pragma unroll
                for (int q = 0; q < WORDS; ++q) cand[q] &= ~w[q][15 - L];
                m = Z;
            } else {
                t -= Z;
#pragma unroll
                for (int q = 0; q < WORDS; ++q) cand[q] &= w[q][15 - L];
                m -= Z;
            }
            if (m <= 32) break;                                // few enough to rank directly
        }
    }
    // rank the remaining candidates exactly in fp64
    __syncthreads();
    if (threadIdx.x == 0) S.list_n = 0;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < WORDS; ++q) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            if ((cand[q] >> e) & 1u) {
                const int slot = atomicAdd(&S.list_n, 1);
                if (slot < kListCap) S.list[slot] = key_of<KIND>(vals, (32 * q + e) * kT + (int)threadIdx.x, med);
            }
        }
    }
    __syncthreads();
    const int mm = S.list_n;
    if (mm > kListCap) return false;                           // very many near-identical values: generic path
    for (int c = threadIdx.x; c < mm; c += kT) {
        const double x = S.list[c];
        int rank = 0;
        for (int i = 0; i < mm; ++i) {
            const double o = S.list[i];
            rank += (o < x || (o == x && i < c)) ? 1 : 0;
        }
        if (rank == t) S.dbc[0] = x;
        if (rank == t + 1) S.dbc[1] = x;
    }
    __syncthreads();
    if (!need2 && threadIdx.x == 0) S.dbc[1] = S.dbc[0];
    __syncthreads();
    return true;
}

// ---------------------------------------------------------------------------
// First choice for rows whose values are all finite (every real scan): ONE
// 1024-bin counting pass (shared-memory atomics), a block scan to find the
// bin(s) that hold ranks t0 <= t1 (t1 - t0 <= 1), and an exact fp64 ranking of
// the handful of elements in those bins.  The bin index is a monotone
// non-decreasing function of the key, so every element of a lower bin is <=
// every element of a higher one and the order statistic is exact whatever the
// bin edges are.  The edges only decide how many elements share the target bin:
// they come from the quartiles of a 32-element sample of the row (q_lo, q_hi;
// limb pixels make min / max useless), 1022 bins over the quartile range widened
// by 4 IQR each side plus one open bin at each end; for the MAD the keys
// |rat - med| get 1023 bins over [0, 4 IQR) plus one open bin.
// ~8 instructions per element and 5 barriers per select (the bit-sliced select
// below costs ~16 block reductions and a 32x32 bit transpose per 32 elements).
// Returns false (nothing decided) when the target bins hold more than kListCap
// elements (heavy duplication, unrepresentative sample): the caller falls back.
// bin = clamp(floor((key - lo) * scale) + first, 0, 1023) through the 1.5*2^52 trick (no conversion pipe);
// |(key - lo) * scale| < 2^31 because |key| <= 23 (log ratios of uint16) and iqr >= 1e-5
struct HistBins {
    double lo, scale, magic;
    template <int KIND>
    __device__ __forceinline__ void set(double qlo, double qhi) {
        const double iqr = qhi - qlo;
        lo = KIND == 0 ? qlo - 4.0 * iqr : 0.0;
        scale = KIND == 0 ? (double)(kHistBins - 2) / (9.0 * iqr) : (double)(kHistBins - 1) / (4.0 * iqr);
        magic = 6755399441055744.0 + (KIND == 0 ? 1.0 : 0.0);
    }
    __device__ __forceinline__ int operator()(double x) const {
        const int b = __double2loint(__fma_rd(x - lo, scale, magic));
        return min(max(b, 0), kHistBins - 1);
    }
};

__device__ __forceinline__ bool hist_usable(int n, double qlo, double qhi) {
    const double iqr = qhi - qlo;
    return n > kListCap && iqr >= 1e-5 && iqr < 1e30;
}

// `counted`: the caller has already zeroed H and counted every key into it (the median's
// counting pass is fused into the loop that computes the ratios).
template <int KIND, int kT>
__device__ __forceinline__ bool hist_select(const double* vals, int n, double med, double qlo, double qhi, int t0, int t1,
                                            bool counted, Shared& S) {
    constexpr int BPT = kHistBins / kT;
    constexpr int NW = kT / 32;
    static_assert(BPT >= 1 && BPT * kT == kHistBins, "thread count must divide the bin count");
    unsigned int* H = &S.whist[0][0];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int b0 = 0, b1 = kHistBins - 1, below = 0, mm = n;
    HistBins bin_of;
    bin_of.set<KIND>(qlo, qhi);
    if (n > kListCap) {
        if (!hist_usable(n, qlo, qhi)) return false;          // (uniform across the block)
        if (!counted) {
#pragma unroll
            for (int q = 0; q < BPT; ++q) H[q * kT + threadIdx.x] = 0;
            if (threadIdx.x == 0) S.list_n = 0;
            __syncthreads();
#pragma unroll 4
            for (int i = threadIdx.x; i < n; i += kT) atomicAdd(&H[bin_of(key_of<KIND>(vals, i, med))], 1u);
            __syncthreads();
        }
        // exclusive scan over the bins: BPT consecutive bins per thread
        unsigned int c[BPT], s = 0;
        if constexpr (BPT % 4 == 0) {
#pragma unroll
            for (int q = 0; q < BPT; q += 4) {
                const uint4 v = *reinterpret_cast<const uint4*>(H + threadIdx.x * BPT + q);
                c[q] = v.x; c[q + 1] = v.y; c[q + 2] = v.z; c[q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < BPT; ++q) c[q] = H[threadIdx.x * BPT + q];
        }
#pragma unroll
        for (int q = 0; q < BPT; ++q) s += c[q];
        unsigned int inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) S.hist[warp] = inc;
        __syncthreads();
        unsigned int exc = inc - s;
#pragma unroll
        for (int w = 0; w < NW; ++w) exc += w < warp ? S.hist[w] : 0u;
        if ((unsigned)t1 >= exc && (unsigned)t0 < exc + s) {
#pragma unroll
            for (int q = 0; q < BPT; ++q) {
                if (c[q]) {
                    if ((unsigned)t0 >= exc && (unsigned)t0 < exc + c[q]) { S.ibc[0] = threadIdx.x * BPT + q; S.ibc[1] = (int)exc; }
                    if ((unsigned)t1 >= exc && (unsigned)t1 < exc + c[q]) { S.ibc[2] = threadIdx.x * BPT + q; S.ibc[3] = (int)(exc + c[q]); }
                    exc += c[q];
                }
            }
        }
        __syncthreads();
        b0 = S.ibc[0]; below = S.ibc[1]; b1 = S.ibc[2];
        mm = S.ibc[3] - below;                                // bins strictly between b0 and b1 are empty
        if (mm > kListCap) {
            __syncthreads();                                  // H (aliases whist) is reused by the caller's fallback
            return false;
        }
#pragma unroll 4
        for (int i = threadIdx.x; i < n; i += kT) {
            const double x = key_of<KIND>(vals, i, med);
            const int b = bin_of(x);
            if (b >= b0 && b <= b1) S.list[atomicAdd(&S.list_n, 1)] = x;
        }
    } else {
        // short chord: rank everything
        for (int i = threadIdx.x; i < n; i += kT) S.list[i] = key_of<KIND>(vals, i, med);
    }
    __syncthreads();
    for (int ci = threadIdx.x; ci < mm; ci += kT) {
        const double x = S.list[ci];
        int rank = below;
        for (int i = 0; i < mm; ++i) {
            const double o = S.list[i];
            rank += (o < x || (o == x && i < ci)) ? 1 : 0;
        }
        if (rank == t0) S.dbc[0] = x;
        if (rank == t1) S.dbc[1] = x;
    }
    __syncthreads();
    return true;
}

// value of rank t in the full key set: nneg keys are -inf, then nfin finite keys in [lo, hi], then +inf
template <int KIND, int kT>
__device__ void ranked_pair(const double* vals, int n, double med, int nneg, int nfin,
                            int t0, int t1, double& v0, double& v1, Shared& S, bool hist_ok, double klo, double khi,
                            bool counted) {
    auto group = [&](int t) { return t < nneg ? -1 : (t < nneg + nfin ? 0 : 1); };
    const int g0 = group(t0), g1 = group(t1);
    if (hist_ok && nfin == n && hist_select<KIND, kT>(vals, n, med, klo, khi, t0, t1, counted, S)) {
        v0 = S.dbc[0];
        v1 = S.dbc[1];
        __syncthreads();
        return;
    }
    if (g0 == 0 && g1 == 0) {
        // the common case: both middle ranks are finite values
        constexpr int kWords = kT <= 64 ? 2 : 1;              // 64-thread CTAs own up to 64 elements per thread
        if (n <= 32 * kWords * kT && fast_select<KIND, kT, kWords>(vals, n, med, t0 - nneg, t1 != t0, S)) {
            v0 = S.dbc[0];
            v1 = S.dbc[1];
            __syncthreads();
            return;
        }
    }
    // generic path (long chords, very many ties, or a middle rank that is +-inf): needs the finite key range
    double lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < n; i += kT) {
        const double x = key_of<KIND>(vals, i, med);
        if (fabs(x) < INFINITY) { lo = fmin(lo, x); hi = fmax(hi, x); }
    }
    block_minmax(lo, hi, S);
    __syncthreads();
    if (g0 == 0 && g1 == 0) {
        block_select<KIND, kT>(vals, n, med, lo, hi, nfin, t0 - nneg, t1 != t0, S);
        v0 = S.dbc[0];
        v1 = S.dbc[1];
        __syncthreads();
        return;
    }
    // rows with so many zeros that a middle rank is +-inf: rare, generic path
    if (g0 == 0) {
        block_select<KIND, kT>(vals, n, med, lo, hi, nfin, t0 - nneg, false, S);
        v0 = S.dbc[0];
        __syncthreads();
    } else {
        v0 = g0 < 0 ? -INFINITY : INFINITY;
    }
    if (g1 == 0) {
        block_select<KIND, kT>(vals, n, med, lo, hi, nfin, t1 - nneg, false, S);
        v1 = S.dbc[0];
        __syncthreads();
    } else {
        v1 = g1 < 0 ? -INFINITY : INFINITY;
    }
}

// ---------------------------------------------------------------------------
// log(v) for a 16-bit pixel without a memory gather (a 65536-entry fp64 table costs one divergent
// 8-byte L1 / L2 gather per pixel: the row-statistics kernel spent 2/3 of its time waiting on them).
//   v = m16 * 2^(e-15), m16 in [2^15, 2^16);  hi = top 7 mantissa bits;  c = fl(1 / (1 + (hi + 1/2)/128))
//   log(v) = e*ln2 - log(c) + log1p(r),  r = m*c - 1 (one fma: exact to 2^-62), |r| <= 2^-8,
//   log1p by its degree-6 Taylor polynomial (remainder < 2e-18).  {c / 2^15, -log(c)} come from a
// 128-entry (2 KB) table staged in shared memory; -log(c) is computed on the host in long double.
// Absolute error <= ~2e-15 (two roundings at the magnitude of the result, <= 11.1), the same as a table
// of correctly rounded logs.  v == 0 -> -inf, as np.log(0).
struct LogSeg { double c, t; };
static LogSeg g_logseg_host[128];
__device__ LogSeg g_logseg[128];

static void fill_logseg_host() {
    static const bool once = [] {                            // thread-safe one-time initialisation
        for (int h = 0; h < 128; ++h) {
            const double c = 1.0 / (1.0 + (h + 0.5) / 128.0);
            g_logseg_host[h].c = c / 32768.0;
            g_logseg_host[h].t = (double)(-logl((long double)c));
        }
        return true;
    }();
    (void)once;
}

__device__ __forceinline__ double log_u16(uint32_t v, const LogSeg* seg /* shared */) {
    const int e = 31 - __clz((int)v);
    const uint32_t m16 = v << ((15 - e) & 31);
    const double2 ct = *reinterpret_cast<const double2*>(&seg[(m16 >> 8) & 127u]);
    const double r = fma(u32_to_double(m16), ct.x, -1.0);
    double p = fma(r, -1.0 / 6.0, 0.2);
    p = fma(r, p, -0.25);
    p = fma(r, p, 1.0 / 3.0);
    p = fma(r, p, -0.5);
    const double ed = u32_to_double((uint32_t)e);
    const double hi = fma(ed, 6.93147180369123816490e-01, ct.y);       // e * ln2_hi is exact (ln2_hi has 32 bits)
    const double lo = fma(ed, 1.90821492927058770002e-10, fma(r * r, p, r));
    return v ? hi + lo : -INFINITY;
}

constexpr size_t kSharedBytes = (sizeof(Shared) + 15) / 16 * 16;
constexpr size_t kHeadBytes = kSharedBytes + 128 * sizeof(LogSeg);

// log(a / b) for two 16-bit pixels.  Neighbouring rows differ by noise, so almost every pair has
// |z| <= 2^-6 with z = (a - b)/(a + b), where log(a/b) = 2 atanh(z) = 2 (z + z^3/3 + ... + z^9/9)
// (remainder < 2^-60 relative) costs one reciprocal and five fma -- about half of two log_u16 -- and
// is good to a few ulp of the RESULT.  Other pairs (limb, zeros) take L(a) - L(b).
__device__ __forceinline__ bool log_ratio_small(uint32_t a, uint32_t b) {
    const uint32_t den = a + b;
    return ((uint32_t)abs((int)a - (int)b) << 6) <= den && den != 0u;
}
// the series; only meaningful when log_ratio_small(a, b) (garbage, but no trap, otherwise)
__device__ __forceinline__ double log_ratio_series(uint32_t a, uint32_t b) {
    const double dn = u32_to_double(a + b);
    const double nn = u32_to_double(a + 65536u - b) - 65536.0;
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(dn));           // ~20 bits; den in [1, 131070]
    double e = fma(-dn, x, 1.0);
    x = fma(x, e, x);
    e = fma(-dn, x, 1.0);
    x = fma(x, e, x);
    const double z = nn * x;
    const double z2 = z * z;
    double p = fma(z2, 1.0 / 9.0, 1.0 / 7.0);
    p = fma(z2, p, 0.2);
    p = fma(z2, p, 1.0 / 3.0);
    const double h = fma(z * z2, p, z);
    return h + h;
}
__device__ __forceinline__ double log_ratio_u16(uint32_t a, uint32_t b, const LogSeg* seg) {
    return log_ratio_small(a, b) ? log_ratio_series(a, b) : log_u16(a, seg) - log_u16(b, seg);
}

__global__ void __launch_bounds__(256)
log_u16_eval_kernel(double* __restrict__ out) {
    __shared__ __align__(16) LogSeg seg[128];
    if (threadIdx.x < 128) seg[threadIdx.x] = g_logseg[threadIdx.x];
    __syncthreads();
    const int v = blockIdx.x * 256 + threadIdx.x;
    if (v < 65536) out[v] = log_u16((uint32_t)v, seg);
}

template <int kT>
__global__ void __launch_bounds__(kT, (kT == 64 ? 6 : (kT == 128 ? 6 : (kT == 256 ? 3 : 1))))
transv_row_stats_kernel(const uint16_t* __restrict__ img_base, int64_t img_stride, int cols,
                        const int32_t* __restrict__ rows, const int32_t* __restrict__ xa_list,
                        const int32_t* __restrict__ xb_list, int n_list,
                        double* __restrict__ out,
                        double* __restrict__ gscratch, int64_t scratch_pitch, int smem_cap, int use_hist,
                        const uint32_t* __restrict__ todo, const unsigned int* __restrict__ todo_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared& S = *reinterpret_cast<Shared*>(smem_raw);
    LogSeg* seg = reinterpret_cast<LogSeg*>(smem_raw + kSharedBytes);
    double* const vals_shared = reinterpret_cast<double*>(smem_raw + kHeadBytes);
    for (int h = threadIdx.x; h < 128; h += kT) seg[h] = g_logseg[h];
    __syncthreads();
    // Two launch shapes: grid (n_list, n_imgs), one CTA per (row, image); or, with a todo list (the rows the
    // register-resident kernel below handed back), a small persistent grid that walks the list.
    const bool listed = todo != nullptr;
    int64_t work = listed ? (int64_t)blockIdx.x : (int64_t)blockIdx.y * n_list + blockIdx.x;
    const int64_t work_end = listed ? (int64_t)*todo_count : work + 1;
    const int64_t work_step = listed ? (int64_t)gridDim.x : 1;
    auto one_row = [&](const int64_t slot) {
    const int j = (int)(slot % n_list);
    const uint16_t* img = img_base + (slot / n_list) * img_stride;
    double* vals = vals_shared;
    const int y = rows[j], xa = xa_list[j], xb = xb_list[j];
    const int n = xb - xa;
    if (n <= 0) {                                   // np.mean of an empty slice
        if (threadIdx.x == 0) out[slot] = NAN;
        return;
    }
    if constexpr (kT == 1024) {                     // only chords > 16384 px can exceed shared memory; the other
        if (n > smem_cap) vals = gscratch + slot * scratch_pitch;   // variants keep a provably-shared pointer (LDS / STS)
    }
    const uint16_t* ry = img + (int64_t)y * cols + xa;
    const uint16_t* rp = img + (int64_t)(y - 1) * cols + xa;

    // ---- bin edges of the counting select from 32 sample ratios (warp 0), before the main pass ----
    const bool want_hist = (use_hist & 1) && n > kListCap;
    if (want_hist) {
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            const int i = (int)(((int64_t)(2 * lane + 1) * n) >> 6);
            const double x = log_ratio_u16(ry[i], rp[i], seg);
            S.list[lane] = x;
            __syncwarp();
            int rank = 0;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const double o = S.list[q];
                rank += (o < x || (o == x && q < lane)) ? 1 : 0;
            }
            if (rank == 8) S.dbc[2] = x;
            if (rank == 23) S.dbc[3] = x;
        }
        unsigned int* H = &S.whist[0][0];
        for (int q = threadIdx.x; q < kHistBins; q += kT) H[q] = 0;
        if (threadIdx.x == 0) S.list_n = 0;
        __syncthreads();
    }
    const double qlo = want_hist ? S.dbc[2] : 0.0, qhi = want_hist ? S.dbc[3] : 0.0;
    const bool counted = want_hist && hist_usable(n, qlo, qhi);     // nan / inf samples -> not usable
    HistBins bin0;
    bin0.set<0>(qlo, qhi);

    // ---- rat = log(img[y]/img[y-1]) (+ the counting pass of the median select) -------------------
    unsigned int nnan = 0, nneg = 0, npos = 0;
    // U elements per thread and trip: the series of all U are evaluated unconditionally (independent
    // dependency chains the scheduler can interleave; a branch per element serialised them), the rare
    // pairs outside its range are redone with two logs; the next trip's pixels are loaded meanwhile.
    constexpr int U = 4;                             // (eight chains per thread measured slower: 6.9 vs 5.8 ms)
    uint32_t pa[U], pb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int i = (int)threadIdx.x + u * kT;
        pa[u] = i < n ? ry[i] : 1u;
        pb[u] = i < n ? rp[i] : 1u;
    }
    for (int i0 = threadIdx.x; i0 < n; i0 += kT * U) {
        uint32_t ca[U], cb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { ca[u] = pa[u]; cb[u] = pb[u]; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + (U + u) * kT;
            pa[u] = i < n ? ry[i] : 1u;
            pb[u] = i < n ? rp[i] : 1u;
        }
        double r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = log_ratio_series(ca[u], cb[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!log_ratio_small(ca[u], cb[u])) {               // limb, dust, zeros
                r[u] = log_u16(ca[u], seg) - log_u16(cb[u], seg);
                if (!(fabs(r[u]) < INFINITY) && i0 + u * kT < n) {   // a zero pixel in either row
                    if (r[u] != r[u]) ++nnan;
                    else if (r[u] < 0) ++nneg;
                    else ++npos;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * kT;
            if (i < n) {
                vals[i] = r[u];
                if (counted) atomicAdd(&S.whist[0][0] + bin0(r[u]), 1u);
            }
        }
    }
    if (threadIdx.x < 4) S.cnt[threadIdx.x] = 0;
    __syncthreads();
    {
        const unsigned int a = __reduce_add_sync(0xffffffffu, nnan);
        const unsigned int b = __reduce_add_sync(0xffffffffu, nneg);
        const unsigned int c = __reduce_add_sync(0xffffffffu, npos);
        if ((threadIdx.x & 31) == 0) {
            if (a) atomicAdd(&S.cnt[0], (unsigned long long)a);
            if (b) atomicAdd(&S.cnt[1], (unsigned long long)b);
            if (c) atomicAdd(&S.cnt[2], (unsigned long long)c);
        }
    }
    __syncthreads();                                 // counts complete; also orders the writes of vals[]
    const int t_nan = (int)S.cnt[0], t_neg = (int)S.cnt[1], t_pos = (int)S.cnt[2];
    const int nfin = n - t_neg - t_pos;
    if (t_nan > 0) {                                 // np.median -> nan -> everything rejected -> mean([]) = nan
        if (threadIdx.x == 0) out[slot] = NAN;
        return;
    }
    // ---- median (np.median: mean of the two middle values for even n) ---------
    const int t0 = (n - 1) / 2, t1 = n / 2;
    // counting select (hist_select) for rows without zeros
    const bool hist_ok = (use_hist & 1) && nfin == n;
    if (use_hist & 4) return;                        // diagnostics: cost of the rat phase alone
    double a0, a1;
    ranked_pair<0, kT>(vals, n, 0.0, t_neg, nfin, t0, t1, a0, a1, S, hist_ok, qlo, qhi, counted);
    const double med = t0 == t1 ? a0 : (a0 + a1) / 2.0;
    if (!(fabs(med) < INFINITY)) {                   // |rat - med| contains nan -> mean([]) = nan
        if (threadIdx.x == 0) out[slot] = NAN;
        return;
    }
    // ---- MAD -------------------------------------------------------------------
    const int ninf = t_neg + t_pos;
    double b0, b1;
    if (use_hist & 8) return;                        // diagnostics: rat phase + median
    ranked_pair<1, kT>(vals, n, med, 0, n - ninf, t0, t1, b0, b1, S, hist_ok, qlo, qhi, false);
    const double mdev = t0 == t1 ? b0 : (b0 + b1) / 2.0;
    // ---- mean of the inliers ---------------------------------------------------
    double sum = 0.0;
    int kept = 0;
    if (mdev != 0.0) {                               // (nan is truthy in the reference too, but cannot occur here)
        // fl(d / mdev) < 2  <=>  d < 2*mdev for positive normal doubles (2*mdev is exact, and the largest
        // double below it divides to at most 2 - 2^-52), so the per-element division is not needed
        const double thr = 2.0 * mdev;
        if (mdev > 1e-300 && mdev < 1e300) {
#pragma unroll 4
            for (int i = threadIdx.x; i < n; i += kT) {
                const double r = vals[i];
                const bool keep = fabs(r - med) < thr;
                sum += keep ? r : 0.0;
                kept += keep ? 1 : 0;
            }
        } else {
            for (int i = threadIdx.x; i < n; i += kT) {
                const double r = vals[i];
                if (fabs(r - med) / mdev < 2.0) { sum += r; ++kept; }
            }
        }
    } else {
        for (int i = threadIdx.x; i < n; i += kT) { sum += vals[i]; ++kept; }
    }
    double cnt = (double)kept;
    block_sum2(sum, cnt, S);
    if (threadIdx.x == 0) out[slot] = sum / cnt;
    };
    for (; work < work_end; work += work_step) {
        one_row(listed ? (int64_t)todo[work] : work);
        __syncthreads();                             // shared state is reused by the next row
    }
}

// ---------------------------------------------------------------------------
// Register-resident row statistics (the launch every real scan takes).
// A thread keeps 16 ratios in REGISTERS (element e of thread t is chord position e*kT + t) and the rest of a chord
// longer than 16 kT in shared memory, so the passes after the first are straight-line code over registers plus a
// short loop, and residency is 8 CTAs of 128 threads per SM (64 registers, 10 KB static + 16 KB dynamic shared
// memory) against 6 for the classic kernel.  Throughput here is rows in flight / latency of one row: a version with
// 256 threads and 4 rows per SM executed fewer instructions and was slower.
//   pass 1  ratios + 2048-bin counts (bins from sample quartiles); block scan -> prefix table P
//   pass 2  WINDOW SELECT for the median: the value window [lo, hi) of the bin(s) holding the middle ranks;
//           every thread counts its values below the window and appends those inside to a list; the list is
//           ranked exactly in fp64.  The counts are taken from the data, so the result does not depend on
//           the bin arithmetic being consistent with the window edges: if the middle ranks do not fall in
//           the list (an edge rounding, a list overflow) the row is handed back.
//   pass 3  WINDOW SELECT for the MAD without a second histogram: with u = position of the median in bin
//           units, the values with |r - med| < k bins lie in bins floor(u)-k .. floor(u)+k, and all values of
//           bins floor(u)-k+1 .. floor(u)+k-1 are that close, so P brackets the count for every k: the
//           largest k_lo with upper(k_lo) <= t0 and the smallest k_hi with lower(k_hi) > t1 (a three-level
//           search, two ballots each) give a distance window [k_lo w, k_hi w) that holds both middle ranks;
//           same count / collect / verify / rank as pass 2, on the keys |r - med|.
//           (tests/test_numerics_cpu.py: test_window_select_with_histogram_brackets_is_exact models passes 2-3.)
//   pass 4  mean of the inliers, in the same per-thread order and with the same block reduction as the
//           classic kernel (identical bits: test_row_stats_counting_select_equals_bitsliced).
// Rows this cannot take (a zero pixel, chords <= 256 or > 32 kT, ties that overflow a list, an
// unrepresentative sample, MAD == 0) are appended to a todo list and done by the classic kernel above.
// The values must stay in registers: anything that makes the compiler spill them (four interleaved chains per
// group instead of two, a noinline call for the rare log path) turns every later pass into a wait on local memory.
constexpr int kBins2 = 2048;
constexpr int kCap2 = 128;            // candidate list (a few dozen with bins this fine)
constexpr int kMinLen2 = 256;         // shorter chords: the classic kernel ranks everything

template <int kT>
struct Shared2 {
    // counts; after the scan the exclusive prefix, entry kBins2 = n.  Bin b lives at H[b + b / (bins per thread)]:
    // the scan gives a thread consecutive bins, and the padding word makes its stride odd (no bank conflicts)
    unsigned int H[kBins2 + kT + 8];
    double list[kCap2];
    double red[32];
    unsigned int wsum[32];
    double dbc[4];
    int ibc[4];
    int list_n, below, kept;
};

// Bins of the register-resident kernel: kBins2 - 2 regular bins over the sample's inter-quartile range widened by
// 3 IQR each side (a 32-value sample can misjudge the spread by a factor of two and med +- MAD still lies inside;
// whatever falls outside goes to the two open bins), i.e. a bin is IQR/292 wide: a handful of values per bin at the
// centre of a 3000-pixel chord.
struct HistBins2 {
    double lo, scale, magic, width;
    __device__ __forceinline__ void set(double qlo, double qhi) {
        const double iqr = qhi - qlo;
        lo = qlo - 3.0 * iqr;
        scale = (double)(kBins2 - 2) / (7.0 * iqr);
        width = iqr * (7.0 / (double)(kBins2 - 2));
        magic = 6755399441055744.0 + 1.0;
    }
    __device__ __forceinline__ int operator()(double x) const {
        const int b = __double2loint(__fma_rd(x - lo, scale, magic));
        return min(max(b, 0), kBins2 - 1);
    }
};

// ranks t0, t1 of the multiset {`below` smaller values} + list[0..mm) -> dbc[0], dbc[1].  P = 2^k threads share one
// candidate (each counts a slice of the list, the partial ranks meet by shuffles), so a list of 100 costs a few
// dozen iterations on every warp instead of 100 on three of them while the others wait.
template <int kT>
__device__ __forceinline__ void rank_list2(Shared2<kT>& S, int mm, int below, int t0, int t1) {
    if (mm <= 32) {                                    // the usual case: a lane per list entry, a ballot per candidate
        const int lane = threadIdx.x & 31;
        const double mine = lane < mm ? S.list[lane] : INFINITY;
        for (int c = threadIdx.x >> 5; c < mm; c += kT / 32) {          // (uniform per warp)
            const double x = __shfl_sync(0xffffffffu, mine, c);
            const bool before = lane < mm && (mine < x || (mine == x && lane < c));
            const int rank = below + __popc(__ballot_sync(0xffffffffu, before));
            if (lane == 0) {
                if (rank == t0) S.dbc[0] = x;
                if (rank == t1) S.dbc[1] = x;
            }
        }
        return;
    }
    int lg = 0;
    while (lg < 5 && (mm << (lg + 1)) <= kT) ++lg;     // (uniform)
    const int P = 1 << lg;
    const int p = threadIdx.x & (P - 1);
    const int chunk = (mm + P - 1) >> lg;
    const int j0 = p * chunk, j1 = min(mm, j0 + chunk);
    for (int c0 = 0; c0 < mm; c0 += kT >> lg) {        // (uniform trip count)
        const int c = c0 + ((int)threadIdx.x >> lg);
        const double x = S.list[min(c, mm - 1)];
        int rank = 0;
#pragma unroll 4
        for (int i = j0; i < j1; ++i) {
            const double o = S.list[i];
            rank += (o < x || (o == x && i < c)) ? 1 : 0;
        }
        for (int o = P >> 1; o; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
        rank += below;
        if (c < mm && p == 0) {
            if (rank == t0) S.dbc[0] = x;
            if (rank == t1) S.dbc[1] = x;
        }
    }
}

// the rare pairs outside the series' range (inlined: a call makes the compiler spill the register-resident values)
__device__ __forceinline__ double log_ratio_far(uint32_t a, uint32_t b) {
    return log_u16(a, g_logseg) - log_u16(b, g_logseg);
}

template <int kT>
__device__ __forceinline__ void list_append2(Shared2<kT>& S, double k) {
    const int at = atomicAdd(&S.list_n, 1);
    if (at < kCap2) S.list[at] = k;
}

// ER chord positions per thread in registers, up to ES more in shared memory (position e of thread t is chord
// index e*kT + t; spill[(e - ER)*kT + t]).
template <int kT, int ER, int ES>
__global__ void __launch_bounds__(kT, (1024 / kT > 0 ? 1024 / kT : 1))
transv_row_stats_reg_kernel(const uint16_t* __restrict__ img_base, int64_t img_stride, int cols,
                            const int32_t* __restrict__ rows, const int32_t* __restrict__ xa_list,
                            const int32_t* __restrict__ xb_list, int n_list, double* __restrict__ out,
                            uint32_t* __restrict__ todo, unsigned int* __restrict__ todo_count, int stop) {
    __shared__ Shared2<kT> S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* const spill = reinterpret_cast<double*>(smem_raw);
    constexpr int G = 2;                              // elements per group: one uniform guard, two interleaved chains
                                                      // (four spill: the values must stay in registers)
    constexpr int NG = ER / G;
    constexpr int NW = kT / 32;
    constexpr int kDump = kBins2 + 2;                 // counts of the padding lanes go here (beyond the scan)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = blockIdx.x;
    const int64_t slot = (int64_t)blockIdx.y * n_list + j;
    const int y = rows[j], xa = xa_list[j];
    const int n = xb_list[j] - xa;
    auto defer = [&](int why) {                       // (uniform) hand the row back; todo_count[why] counts the reasons
        if (tid == 0) {
            todo[atomicAdd(todo_count, 1u)] = (uint32_t)slot;
            atomicAdd(todo_count + why, 1u);
        }
    };
    if (n <= kMinLen2 || n > kT * (ER + ES)) { defer(1); return; }
    const uint16_t* img = img_base + (int64_t)blockIdx.y * img_stride;
    const uint16_t* ry = img + (int64_t)y * cols + xa;
    const uint16_t* rp = img + (int64_t)(y - 1) * cols + xa;
    const int n_groups = (n + G * kT - 1) / (G * kT);          // groups of G positions per thread that hold any pixel

    // pixels of the first group are in flight while warp 0 looks at the sample
    uint32_t na[G], nb[G];
#pragma unroll
    for (int u = 0; u < G; ++u) {
        const int i = u * kT + tid;
        na[u] = i < n ? (uint32_t)__ldg(ry + i) : 1u;
        nb[u] = i < n ? (uint32_t)__ldg(rp + i) : 1u;
    }
    // bin edges from the quartiles of 32 sample ratios (warp 0)
    if (warp == 0) {
        const int i = (int)(((int64_t)(2 * lane + 1) * n) >> 6);
        const double x = log_ratio_u16(__ldg(ry + i), __ldg(rp + i), g_logseg);
        S.list[lane] = x;
        __syncwarp();
        int rank = 0;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const double o = S.list[q];
            rank += (o < x || (o == x && q < lane)) ? 1 : 0;
        }
        if (rank == 8) S.dbc[2] = x;
        if (rank == 23) S.dbc[3] = x;
    }
    constexpr int BPT = kBins2 / kT;                  // bins per thread in the scan
    static_assert(BPT >= 2 && BPT * kT == kBins2 && (BPT & (BPT - 1)) == 0, "thread count must divide the bin count");
    constexpr int kPadShift = BPT == 2 ? 1 : (BPT == 4 ? 2 : (BPT == 8 ? 3 : (BPT == 16 ? 4 : 5)));
    static_assert((1 << kPadShift) == BPT, "2 to 32 bins per thread");
    auto pad = [](int b) { return b + (b >> kPadShift); };
    for (int q = tid; q < kBins2 + kT; q += kT) S.H[q] = 0;
    if (tid == 0) { S.list_n = 0; S.below = 0; S.kept = 0; }
    __syncthreads();
    const double qlo = S.dbc[2], qhi = S.dbc[3];
    if (!hist_usable(n, qlo, qhi)) { defer(2); return; }
    if (stop == 1) return;                            // (timing knob SHG_TRANSV_STOP: cost of the phases)
    HistBins2 bin0;
    bin0.set(qlo, qhi);

    // ---- pass 1: ratios + counts ----------------------------------------------------------------------
    double v[ER];
    bool bad = false;
    // one group: the pixels that were in flight, the next group's loads, four series, the rare redo, counts
    auto ratios_of_group = [&](const int g, double (&r)[G]) {
        uint32_t ca[G], cb[G];
#pragma unroll
        for (int u = 0; u < G; ++u) { ca[u] = na[u]; cb[u] = nb[u]; }
        if (g + 1 < n_groups) {                                // (uniform)
#pragma unroll
            for (int u = 0; u < G; ++u) {
                const int i = ((g + 1) * G + u) * kT + tid;
                na[u] = i < n ? (uint32_t)__ldg(ry + i) : 1u;
                nb[u] = i < n ? (uint32_t)__ldg(rp + i) : 1u;
            }
        }
#pragma unroll
        for (int u = 0; u < G; ++u) r[u] = log_ratio_series(ca[u], cb[u]);
#pragma unroll
        for (int u = 0; u < G; ++u) {
            if (!log_ratio_small(ca[u], cb[u])) {              // limb, dust, zeros
                r[u] = log_ratio_far(ca[u], cb[u]);
                bad |= !(fabs(r[u]) < INFINITY);               // a zero pixel: the classic kernel knows the rules
            }
        }
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const bool valid = (g * G + u) * kT + tid < n;
            atomicAdd(&S.H[pad(valid ? bin0(r[u]) : kDump)], 1u);
            r[u] = valid ? r[u] : NAN;                         // padding fails every comparison below
        }
    };
#pragma unroll
    for (int g = 0; g < NG; ++g) {
#pragma unroll
        for (int u = 0; u < G; ++u) v[g * G + u] = NAN;
        if (g < n_groups) {                                    // (uniform)
            double r[G];
            ratios_of_group(g, r);
#pragma unroll
            for (int u = 0; u < G; ++u) v[g * G + u] = r[u];
        }
    }
    if constexpr (ES > 0) {
#pragma unroll 1
        for (int g = NG; g < n_groups; ++g) {
            double r[G];
            ratios_of_group(g, r);
#pragma unroll
            for (int u = 0; u < G; ++u) spill[((g - NG) * G + u) * kT + tid] = r[u];
        }
    }
    if (__syncthreads_or(bad)) { defer(3); return; }
    if (stop == 2) { if (v[0] == 1.25) out[slot] = v[ER - 1]; return; }
    // every value of this thread, in chord order: f(value)
    auto for_each_value = [&](auto f) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            if (g < n_groups) {                                // (uniform)
#pragma unroll
                for (int u = 0; u < G; ++u) f(v[g * G + u]);
            }
        }
        if constexpr (ES > 0) {
#pragma unroll 1
            for (int g = NG; g < n_groups; ++g) {
                double r[G];
#pragma unroll
                for (int u = 0; u < G; ++u) r[u] = spill[((g - NG) * G + u) * kT + tid];
#pragma unroll
                for (int u = 0; u < G; ++u) f(r[u]);
            }
        }
    };

    // ---- exclusive scan of the counts, in place; the bins of the middle ranks ---------------------------
    const int t0 = (n - 1) / 2, t1 = n / 2;
    {
        unsigned int* mine_h = &S.H[tid * (BPT + 1)];           // == &S.H[pad(tid * BPT)]
        unsigned int s = 0;
#pragma unroll
        for (int q = 0; q < BPT; ++q) s += mine_h[q];
        unsigned int inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) S.wsum[warp] = inc;
        __syncthreads();
        unsigned int run = inc - s;
#pragma unroll
        for (int w = 0; w < NW; ++w) run += w < warp ? S.wsum[w] : 0u;
        const bool mine = (unsigned)t1 >= run && (unsigned)t0 < run + s;   // a middle rank in this thread's bins
#pragma unroll
        for (int q = 0; q < BPT; ++q) {
            const unsigned int c = mine_h[q];
            if (mine && c) {
                if ((unsigned)t0 >= run && (unsigned)t0 < run + c) S.ibc[0] = tid * BPT + q;
                if ((unsigned)t1 >= run && (unsigned)t1 < run + c) S.ibc[1] = tid * BPT + q;
            }
            mine_h[q] = run;
            run += c;
        }
        if (tid == kT - 1) S.H[pad(kBins2)] = run;
        __syncthreads();
    }
    if (stop == 3) return;
    const double w = bin0.width;
    // one window select: count keys below [lo, hi), list those inside, verify, rank -> S.dbc[0..1]
    auto window_select = [&](auto key, const double lo, const double hi) -> bool {
        int below = 0;
        for_each_value([&](const double x) {
            const double k = key(x);
            const bool lt_lo = k < lo, lt_hi = k < hi;         // (nan padding: both false)
            below += lt_lo ? 1 : 0;
            if (lt_hi && !lt_lo) list_append2(S, k);           // a few values per row
        });
        below = __reduce_add_sync(0xffffffffu, below);
        if (lane == 0 && below) atomicAdd(&S.below, below);
        __syncthreads();
        const int mm = S.list_n, nb_ = S.below;
        if (mm > kCap2 || t0 < nb_ || t1 >= nb_ + mm) {                     // (uniform)
            if (tid == 0) atomicAdd(todo_count + (mm > kCap2 ? 8 : (t0 < nb_ ? 9 : 10)), 1u);
            return false;
        }
        rank_list2<kT>(S, mm, nb_, t0, t1);
        __syncthreads();
        return true;
    };

    // ---- pass 2: median ---------------------------------------------------------------------------------
    const int b0 = S.ibc[0], b1 = S.ibc[1];
    const double elo = b0 <= 0 ? -INFINITY : bin0.lo + (double)(b0 - 1) * w;
    const double ehi = b1 >= kBins2 - 1 ? INFINITY : bin0.lo + (double)b1 * w;
    if (!window_select([](double r) { return r; }, elo, ehi)) { defer(4); return; }
    const double med = t0 == t1 ? S.dbc[0] : (S.dbc[0] + S.dbc[1]) / 2.0;
    __syncthreads();                                  // dbc / list / counters are reused
    if (stop == 4) return;

    // ---- pass 3: MAD ------------------------------------------------------------------------------------
    // brackets of #{|r - med| < k w} from the prefix table (warp 0)
    if (warp == 0) {
        if (lane == 0) { S.list_n = 0; S.below = 0; }
        const double uu = fmin(fmax((med - bin0.lo) * bin0.scale + 1.0, -4096.0), 8192.0);
        const int fl = __double2int_rd(uu);
        auto P = [&](int idx) { return (int)S.H[pad(min(max(idx, 0), kBins2))]; };
        auto upper = [&](int k) { return k == 0 ? 0 : P(fl + k + 1) - P(fl - k); };
        auto lower = [&](int k) {                     // whole regular bins only (bins 0 and kBins2-1 are open-ended)
            const int hi = min(max(fl + k, 1), kBins2 - 1), lo = min(max(fl - k + 1, 1), kBins2 - 1);
            return hi > lo ? P(hi) - P(lo) : 0;
        };
        // three levels (64, 2, 1) cover k < 2048; both predicates are monotone in k
        static_assert(kBins2 <= 2048, "the bracket search covers 2048 bins");
        unsigned int m = __ballot_sync(0xffffffffu, upper(64 * lane) <= t0);      // lane 0 always votes
        int base = 64 * (31 - __clz(m));
        m = __ballot_sync(0xffffffffu, upper(base + 2 * lane) <= t0);
        int k_lo = base + 2 * (31 - __clz(m));
        if (upper(k_lo + 1) <= t0) ++k_lo;
        m = __ballot_sync(0xffffffffu, lower(64 * lane + 63) > t1);
        int k_hi = -1;
        if (m != 0) {
            base = 64 * (__ffs(m) - 1);
            m = __ballot_sync(0xffffffffu, lower(base + 2 * lane + 1) > t1);
            k_hi = base + 2 * (__ffs(m) - 1) + 1;
            if (lower(k_hi - 1) > t1) --k_hi;
        }
        if (lane == 0) { S.ibc[2] = k_lo; S.ibc[3] = k_hi; }
    }
    __syncthreads();
    const int k_lo = S.ibc[2], k_hi = S.ibc[3];
    if (k_hi < 0) { defer(5); return; }
    if (stop == 5) return;
    if (!window_select([med](double r) { return fabs(r - med); }, (double)k_lo * w, (double)k_hi * w)) {
        defer(6);
        return;
    }
    const double mdev = t0 == t1 ? S.dbc[0] : (S.dbc[0] + S.dbc[1]) / 2.0;
    if (!(mdev > 1e-300 && mdev < 1e300)) { defer(7); return; }
    if (stop == 6) return;

    // ---- pass 4: mean of the inliers (fl(d / mdev) < 2  <=>  d < 2 mdev, see the classic kernel) --------
    // same per-thread order as the classic kernel; the nan padding adds +0.0, which leaves the sum as it is
    const double thr = 2.0 * mdev;
    double sum = 0.0;
    int kept = 0;
    for_each_value([&](const double x) {
        const bool keep = fabs(x - med) < thr;
        sum += keep ? x : 0.0;
        kept += keep ? 1 : 0;
    });
    // the sum as block_sum2 of the classic kernel adds it (identical bits); the count is an integer either way
    kept = __reduce_add_sync(0xffffffffu, kept);
    sum = warp_sum(sum);
    if (lane == 0) { S.red[warp] = sum; atomicAdd(&S.kept, kept); }
    __syncthreads();
    if (warp == 0) {
        sum = warp_sum(lane < NW ? S.red[lane] : 0.0);
        if (lane == 0) out[slot] = sum / (double)S.kept;
    }
}

// Per-row gain from the per-row statistics (reference solex_util.py:400-404, 456-479):
//   trend = savgol_filter(ratios, window, 3)  (mode 'interp': cubic fitted to the first / last window at the ends)
//   detrended = ratios - trend; detrended -= mean; correction = exp(-cumsum(detrended))
//   gain[y1:y2] = 1 + (correction - 1) * taper, 1 elsewhere.
// One CTA per image; ratios / detrended live in shared memory.
constexpr int kGT = 256;

__device__ void cubic_fit_window(const double* x, int w, double* sol /* smem[4]: coefficients in u */, double* red) {
    // least squares cubic in u = (t - c)/c, c = (w-1)/2, by moment sums + 4x4 elimination
    const double c = 0.5 * (w - 1);
    double mom[11];
#pragma unroll
    for (int j = 0; j < 11; ++j) mom[j] = 0.0;
    for (int t = threadIdx.x; t < w; t += kGT) {
        const double u = ((double)t - c) / c;
        double p = 1.0;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            mom[k] += p;
            if (k < 4) mom[7 + k] += p * x[t];
            p *= u;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 11; ++j) mom[j] = warp_sum(mom[j]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int j = 0; j < 11; ++j) red[j * 8 + warp] = mom[j];
    __syncthreads();
    if (threadIdx.x == 0) {
        double M[11];
        for (int j = 0; j < 11; ++j) {
            double t = 0.0;
            for (int q = 0; q < kGT / 32; ++q) t += red[j * 8 + q];
            M[j] = t;
        }
        double A[4][5];
        for (int r = 0; r < 4; ++r) {
            for (int q = 0; q < 4; ++q) A[r][q] = M[r + q];
            A[r][4] = M[7 + r];
        }
        for (int q = 0; q < 4; ++q) {
            int piv = q;
            for (int r = q + 1; r < 4; ++r)
                if (fabs(A[r][q]) > fabs(A[piv][q])) piv = r;
            for (int k = 0; k < 5; ++k) { const double t = A[q][k]; A[q][k] = A[piv][k]; A[piv][k] = t; }
            for (int r = q + 1; r < 4; ++r) {
                const double f = A[r][q] / A[q][q];
                for (int k = q; k < 5; ++k) A[r][k] -= f * A[q][k];
            }
        }
        for (int r = 3; r >= 0; --r) {
            double t = A[r][4];
            for (int k = r + 1; k < 4; ++k) t -= A[r][k] * sol[k];
            sol[r] = t / A[r][r];
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kGT)
transv_gain_kernel(const double* __restrict__ stats /* [n_imgs][n-1] */, int n, int window,
                   const double* __restrict__ coeffs /* [window] */, const double* __restrict__ taper /* [n] */,
                   int y1, int n_rows, double* __restrict__ gains /* [n_imgs][n_rows] */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* r = reinterpret_cast<double*>(smem_raw);          // ratios, then cumsum
    double* d = r + n;                                          // detrended
    double* red = d + n;                                        // 11*8 scratch
    double* sol = red + 96;
    const int img = blockIdx.x;
    const double* st = stats + (int64_t)img * (n - 1);
    double* g = gains + (int64_t)img * n_rows;
    for (int i = threadIdx.x; i < n; i += kGT) r[i] = i == 0 ? 0.0 : st[i - 1];
    for (int i = threadIdx.x; i < n_rows; i += kGT)
        if (i < y1 || i >= y1 + n) g[i] = 1.0;
    __syncthreads();
    const int m = window / 2;
    // interior of the filter
    for (int i = m + threadIdx.x; i < n - m; i += kGT) {
        double acc = 0.0;
        const double* x = r + (i - m);
        for (int k = 0; k < window; ++k) acc += coeffs[window - 1 - k] * x[k];
        d[i] = r[i] - acc;
    }
    // ends: cubic through the first / last window
    const double c = 0.5 * (window - 1);
    cubic_fit_window(r, window, sol, red);
    for (int i = threadIdx.x; i < m; i += kGT) {
        const double u = ((double)i - c) / c;
        d[i] = r[i] - (((sol[3] * u + sol[2]) * u + sol[1]) * u + sol[0]);
    }
    __syncthreads();
    cubic_fit_window(r + (n - window), window, sol, red);
    for (int i = threadIdx.x; i < m; i += kGT) {
        const int t = window - m + i;
        const double u = ((double)t - c) / c;
        d[n - m + i] = r[n - m + i] - (((sol[3] * u + sol[2]) * u + sol[1]) * u + sol[0]);
    }
    __syncthreads();
    // mean
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += kGT) s += d[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    double mean = 0.0;
    for (int q = 0; q < kGT / 32; ++q) mean += red[q];
    mean /= n;
    __syncthreads();
    // cumsum: contiguous chunk per thread, then offsets
    const int per = (n + kGT - 1) / kGT;
    const int a = min(n, (int)threadIdx.x * per), b = min(n, a + per);
    double run = 0.0;
    for (int i = a; i < b; ++i) { run += d[i] - mean; r[i] = run; }
    __syncthreads();
    double* totals = sol + 4;                                   // kGT chunk totals
    totals[threadIdx.x] = run;
    __syncthreads();
    double off = 0.0;
    for (int q = 0; q < (int)threadIdx.x; ++q) off += totals[q];
    for (int i = a; i < b; ++i) {
        const double corr = exp(-(r[i] + off));
        g[y1 + i] = 1.0 + (corr - 1.0) * taper[i];
    }
}

// One warp per image row (8 rows per CTA): the row's gain is read once, the 16-byte body sits on the row's own
// 16-byte grid (rows of an odd width start misaligned), head and tail go by pixel.
__global__ void __launch_bounds__(256)
row_scale_kernel(const uint16_t* __restrict__ img_base, int64_t img_stride, int rows, int cols,
                 const double* __restrict__ gain /* [n_imgs][rows] */, uint16_t* __restrict__ out_base) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    const double g = gain[(int64_t)blockIdx.y * rows + r];
    const uint16_t* src = img_base + (int64_t)blockIdx.y * img_stride + (int64_t)r * cols;
    uint16_t* dst = out_base + (int64_t)blockIdx.y * img_stride + (int64_t)r * cols;
    auto one = [&](uint32_t v) -> uint32_t {
        double p = __dmul_rn(u32_to_double(v), g);
        p = p > 65535.0 ? 65535.0 : p;               // ret[ret > 65535] = 65535
        p = p > 0.0 ? p : 0.0;
        return double_floor_to_u32(p) & 0xffffu;
    };
    const int mis = (int)(((uintptr_t)src & 15) >> 1);                         // pixels past a 16-byte boundary
    const bool same = (((uintptr_t)src ^ (uintptr_t)dst) & 15) == 0 && (((uintptr_t)src & 1) == 0);
    const int head = same ? min(cols, (8 - mis) & 7) : cols;
    const int nv = same ? (cols - head) / 8 : 0;
    for (int c = lane; c < head; c += 32) dst[c] = (uint16_t)one(src[c]);
    const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
#pragma unroll 2
    for (int v = lane; v < nv; v += 32) {
        const uint4 q = ld_stream_u4(s4 + v);
        uint4 o;
        o.x = one(q.x & 0xffffu) | (one(q.x >> 16) << 16);
        o.y = one(q.y & 0xffffu) | (one(q.y >> 16) << 16);
        o.z = one(q.z & 0xffffu) | (one(q.z >> 16) << 16);
        o.w = one(q.w & 0xffffu) | (one(q.w >> 16) << 16);
        d4[v] = o;
    }
    for (int c = head + nv * 8 + lane; c < cols; c += 32) dst[c] = (uint16_t)one(src[c]);
}

}  // namespace

extern "C" int shg_log_table(double* d_tab65536, void* stream) {
    fill_logseg_host();
    SHG_CHECK(cudaMemcpyToSymbolAsync(g_logseg, g_logseg_host, sizeof(g_logseg_host), 0, cudaMemcpyHostToDevice,
                                      as_stream(stream)));
    log_u16_eval_kernel<<<256, 256, 0, as_stream(stream)>>>(d_tab65536);
    SHG_LAUNCH_CHECK();
    return 0;
}

static int transv_threads(int max_len) {
    // 16 chord positions per thread (measured: 32 per thread in half as many threads is slower, 124 registers); both
    // kernels of one call use the same count, so their inlier sums associate identically
    int t = max_len <= 4096 ? 128 : (max_len <= 8192 ? 256 : (max_len <= 16384 ? 512 : 1024));
    if (const char* e = getenv("SHG_TRANSV_T")) {            // tuning knob
        const int v = atoi(e);
        if (v == 64 && max_len <= 4096) t = 64;
        if ((v == 128 || v == 256 || v == 512 || v == 1024) && max_len <= 32 * v) t = v;
    }
    return t;
}

static int64_t transv_smem_cap(int optin) {
    return ((int64_t)optin - (int64_t)kHeadBytes) / 8;
}

// workspace: [64 bytes: todo counter, then how many rows came back for reason 1..7][todo list, one uint32 per (image, row)][scratch rows for chords beyond shared memory]
static int64_t transv_todo_bytes(int n_list, int n_imgs) {
    return 64 + ((int64_t)n_imgs * n_list * 4 + 15) / 16 * 16;
}

extern "C" int64_t shg_transv_workspace_bytes(int n_list, int max_len, int n_imgs) {
    int dev = 0, optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return -1;
    const int64_t todo = transv_todo_bytes(n_list, n_imgs);
    if (max_len <= transv_smem_cap(optin)) return todo;
    return todo + (int64_t)n_imgs * n_list * (((int64_t)max_len + 15) / 16 * 16) * 8;
}

extern "C" int shg_transv_row_stats(const uint16_t* d_img, int rows, int cols, int n_imgs, int64_t img_stride,
                                    const int32_t* d_rows, const int32_t* d_xa, const int32_t* d_xb, int n_list,
                                    int max_len, double* d_out, void* d_work, int64_t work_bytes, void* stream) {
    (void)rows;
    if (n_list <= 0 || n_imgs <= 0) return 0;
    SHG_REQUIRE(max_len >= 0 && max_len <= cols, "shg_transv_row_stats: max_len %d out of range", max_len);
    SHG_REQUIRE(n_imgs <= 65535, "shg_transv_row_stats: too many images");
    SHG_REQUIRE((int64_t)n_imgs * n_list < (int64_t)1 << 32, "shg_transv_row_stats: too many rows");
    int dev = 0, optin = 0;
    SHG_CHECK(cudaGetDevice(&dev));
    SHG_CHECK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int64_t head = (int64_t)kHeadBytes;
    const int64_t cap = transv_smem_cap(optin);
    const int64_t todo_bytes = transv_todo_bytes(n_list, n_imgs);
    int64_t pitch = 0;
    size_t smem = (size_t)head;
    if (max_len <= cap) {
        smem += (size_t)max_len * 8;
    } else {
        pitch = ((int64_t)max_len + 15) / 16 * 16;
        SHG_REQUIRE(d_work && work_bytes >= todo_bytes + (int64_t)n_imgs * n_list * pitch * 8,
                    "shg_transv_row_stats: chords of %d px need %lld bytes of workspace", max_len,
                    (long long)(todo_bytes + (int64_t)n_imgs * n_list * pitch * 8));
    }
    const int smem_cap = max_len <= cap ? max_len : 0;
    SHG_REQUIRE(max_len <= cap || transv_threads(max_len) == 1024, "shg_transv_row_stats: internal: %d px chords", max_len);
    cudaStream_t st = as_stream(stream);
    fill_logseg_host();
    SHG_CHECK(cudaMemcpyToSymbolAsync(g_logseg, g_logseg_host, sizeof(g_logseg_host), 0, cudaMemcpyHostToDevice, st));
    const int threads = transv_threads(max_len);
    int use_hist = 1;                // SHG_TRANSV_HIST=0: bit-sliced select only; +4 / +8: stop after the rat phase / median (timing)
    if (const char* e = getenv("SHG_TRANSV_HIST")) use_hist = atoi(e);
    // the register-resident kernel first (every row of a real scan), the classic kernel for what it hands back;
    // SHG_TRANSV_REG=0 or SHG_TRANSV_HIST != 1: the classic kernel on every row
    bool reg = use_hist == 1 && threads >= 128 && max_len <= 32 * threads && d_work && work_bytes >= todo_bytes;
    if (const char* e = getenv("SHG_TRANSV_REG")) reg = reg && atoi(e) != 0;
    int stop = 0;                    // SHG_TRANSV_STOP=k: the register-resident kernel returns after phase k (timing only)
    if (const char* e = getenv("SHG_TRANSV_STOP")) stop = atoi(e);
    // 16 chord positions per thread in registers; longer chords keep up to 16 more per thread in shared memory
    const bool spill = max_len > 16 * threads;
    const size_t reg_smem = spill ? (size_t)16 * threads * sizeof(double) : 0;
    unsigned int* todo_count = static_cast<unsigned int*>(d_work);
    uint32_t* todo = reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(d_work) + 64);
    double* scratch = d_work ? reinterpret_cast<double*>(static_cast<unsigned char*>(d_work) + todo_bytes) : nullptr;
    dim3 grid(n_list, n_imgs);
    if (reg) {
        SHG_CHECK(cudaMemsetAsync(todo_count, 0, 64, st));
#define SHG_TRANSV_REG_LAUNCH(T)                                                                                    \
        do {                                                                                                        \
            if (!spill) {                                                                                           \
                transv_row_stats_reg_kernel<T, 16, 0><<<grid, T, 0, st>>>(d_img, img_stride, cols, d_rows, d_xa,    \
                                                                           d_xb, n_list, d_out, todo, todo_count,   \
                                                                           stop);                                   \
            } else {                                                                                                \
                SHG_CHECK(cudaFuncSetAttribute(transv_row_stats_reg_kernel<T, 16, 16>,                              \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)reg_smem));        \
                transv_row_stats_reg_kernel<T, 16, 16><<<grid, T, reg_smem, st>>>(d_img, img_stride, cols, d_rows,  \
                                                                                   d_xa, d_xb, n_list, d_out, todo, \
                                                                                   todo_count, stop);               \
            }                                                                                                       \
        } while (0)
        if (threads == 128) SHG_TRANSV_REG_LAUNCH(128);
        else if (threads == 256) SHG_TRANSV_REG_LAUNCH(256);
        else if (threads == 512) SHG_TRANSV_REG_LAUNCH(512);
        else SHG_TRANSV_REG_LAUNCH(1024);
        SHG_LAUNCH_CHECK();
        grid = dim3((unsigned)std::min<int64_t>((int64_t)n_list * n_imgs, 2 * SHG_SM_COUNT_B200), 1);
    }
    const uint32_t* todo_arg = reg ? todo : nullptr;
#define SHG_TRANSV_LAUNCH(T)                                                                                        \
    do {                                                                                                            \
        SHG_CHECK(cudaFuncSetAttribute(transv_row_stats_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                       (int)smem));                                                                 \
        transv_row_stats_kernel<T><<<grid, T, smem, st>>>(d_img, img_stride, cols, d_rows, d_xa, d_xb, n_list,      \
                                                          d_out, scratch, pitch, smem_cap, use_hist, todo_arg,      \
                                                          todo_count);                                              \
    } while (0)
    if (threads == 64) SHG_TRANSV_LAUNCH(64);
    else if (threads == 128) SHG_TRANSV_LAUNCH(128);
    else if (threads == 256) SHG_TRANSV_LAUNCH(256);
    else if (threads == 512) SHG_TRANSV_LAUNCH(512);
    else SHG_TRANSV_LAUNCH(1024);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_row_scale_u16(const uint16_t* d_img, int rows, int cols, int n_imgs, int64_t img_stride,
                                 const double* d_gain, uint16_t* d_out, void* stream) {
    if (rows <= 0 || cols <= 0 || n_imgs <= 0) return 0;
    SHG_REQUIRE(n_imgs <= 65535, "shg_row_scale_u16: too many images");
    row_scale_kernel<<<dim3((rows + 7) / 8, n_imgs), 256, 0, as_stream(stream)>>>(d_img, img_stride, rows, cols, d_gain,
                                                                                 d_out);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_transv_gain(const double* d_stats, int n_imgs, int n, int window, const double* d_coeffs,
                               const double* d_taper, int y1, int n_rows, double* d_gains, void* stream) {
    if (n_imgs <= 0) return 0;
    SHG_REQUIRE(window >= 5 && (window & 1) && window <= n, "shg_transv_gain: window %d for %d rows", window, n);
    SHG_REQUIRE(y1 >= 0 && y1 + n <= n_rows, "shg_transv_gain: rows [%d, %d) outside the image", y1, y1 + n);
    const size_t smem = ((size_t)2 * n + 96 + 4 + kGT) * sizeof(double);
    int dev = 0, optin = 0;
    SHG_CHECK(cudaGetDevice(&dev));
    SHG_CHECK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SHG_REQUIRE((int)smem <= optin, "shg_transv_gain: %d rows do not fit shared memory", n);
    SHG_CHECK(cudaFuncSetAttribute(transv_gain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    transv_gain_kernel<<<n_imgs, kGT, smem, as_stream(stream)>>>(d_stats, n, window, d_coeffs, d_taper, y1, n_rows,
                                                                d_gains);
    SHG_LAUNCH_CHECK();
    return 0;
}
