// Pass 1: integer mean / max frame of the raw stack.
// Replaces the reference's per-frame NumPy loop (solex_util.py:174-188):
//   my_data(uint64) += img ; max_data = np.maximum(max_data, img)
// Bound: HBM read of the whole stack, frame_px * bytes_per_px bytes per frame.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 8;

__device__ __forceinline__ void acc16(uint32_t w, uint32_t& lo, uint32_t& hi, uint32_t& m) {
    lo = __dp2a_lo(w, 0x00000001u, lo);   // += low  16-bit lane
    hi = __dp2a_lo(w, 0x00000100u, hi);   // += high 16-bit lane
    m = __vmaxu2(m, w);
}

// One thread owns 8 consecutive uint16 pixels (one 16-byte vector) and walks a
// range of frames; 32-bit partial sums are exact for <= 65536 frames per split.
__global__ void __launch_bounds__(kThreads)
accumulate_u16_kernel(const uint4* __restrict__ frames, int64_t n_frames, int64_t vec_per_frame,
                      int64_t frames_per_split, unsigned long long* __restrict__ sum,
                      unsigned int* __restrict__ mx) {
    const int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (v >= vec_per_frame) return;
    const int64_t k0 = (int64_t)blockIdx.y * frames_per_split;
    const int64_t k1 = min(n_frames, k0 + frames_per_split);
    uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0};
    const uint4* p = frames + k0 * vec_per_frame + v;
    int64_t k = k0;
    for (; k + kUnroll <= k1; k += kUnroll) {
        uint4 r[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) r[u] = ld_stream_u4(p + u * vec_per_frame);
        p += kUnroll * vec_per_frame;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            acc16(r[u].x, lo[0], hi[0], m[0]);
            acc16(r[u].y, lo[1], hi[1], m[1]);
            acc16(r[u].z, lo[2], hi[2], m[2]);
            acc16(r[u].w, lo[3], hi[3], m[3]);
        }
    }
    for (; k < k1; ++k) {
        uint4 r = ld_stream_u4(p);
        p += vec_per_frame;
        acc16(r.x, lo[0], hi[0], m[0]);
        acc16(r.y, lo[1], hi[1], m[1]);
        acc16(r.z, lo[2], hi[2], m[2]);
        acc16(r.w, lo[3], hi[3], m[3]);
    }
    const int64_t base = v * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        atomicAdd(sum + base + 2 * j, (unsigned long long)lo[j]);
        atomicAdd(sum + base + 2 * j + 1, (unsigned long long)hi[j]);
        atomicMax(mx + base + 2 * j, m[j] & 0xffffu);
        atomicMax(mx + base + 2 * j + 1, m[j] >> 16);
    }
}

// 8-bit stack: one thread owns 16 pixels; dp4a splits the byte lanes.
__global__ void __launch_bounds__(kThreads)
accumulate_u8_kernel(const uint4* __restrict__ frames, int64_t n_frames, int64_t vec_per_frame,
                     int64_t frames_per_split, unsigned long long* __restrict__ sum,
                     unsigned int* __restrict__ mx) {
    const int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (v >= vec_per_frame) return;
    const int64_t k0 = (int64_t)blockIdx.y * frames_per_split;
    const int64_t k1 = min(n_frames, k0 + frames_per_split);
    uint32_t acc[16];
    uint32_t m[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0;
    const uint4* p = frames + k0 * vec_per_frame + v;
    for (int64_t k = k0; k < k1; k += 4) {
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            r[u] = (k + u < k1) ? ld_stream_u4(p + u * vec_per_frame) : make_uint4(0, 0, 0, 0);
        p += 4 * vec_per_frame;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t w[4] = {r[u].x, r[u].y, r[u].z, r[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[4 * j + 0] = __dp4a(w[j], 0x00000001u, acc[4 * j + 0]);
                acc[4 * j + 1] = __dp4a(w[j], 0x00000100u, acc[4 * j + 1]);
                acc[4 * j + 2] = __dp4a(w[j], 0x00010000u, acc[4 * j + 2]);
                acc[4 * j + 3] = __dp4a(w[j], 0x01000000u, acc[4 * j + 3]);
                m[j] = __vmaxu4(m[j], w[j]);
            }
        }
    }
    const int64_t base = v * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        atomicAdd(sum + base + j, (unsigned long long)acc[j]);
        atomicMax(mx + base + j, (m[j >> 2] >> (8 * (j & 3))) & 0xffu);
    }
}

// Any geometry (frame bytes not a multiple of 16): one thread per pixel.
template <typename T>
__global__ void __launch_bounds__(kThreads)
accumulate_generic_kernel(const T* __restrict__ frames, int64_t n_frames, int64_t frame_px,
                          int64_t frames_per_split, unsigned long long* __restrict__ sum,
                          unsigned int* __restrict__ mx) {
    const int64_t px = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (px >= frame_px) return;
    const int64_t k0 = (int64_t)blockIdx.y * frames_per_split;
    const int64_t k1 = min(n_frames, k0 + frames_per_split);
    unsigned long long s = 0;
    unsigned int m = 0;
    for (int64_t k = k0; k < k1; ++k) {
        unsigned int v = frames[k * frame_px + px];
        s += v;
        m = max(m, v);
    }
    atomicAdd(sum + px, s);
    atomicMax(mx + px, m);
}

// mean = floor(scale*sum / N), written in image orientation.  For rotated
// scans img[i][j] = raw[j][W-1-i]: 32x32 tiles through shared memory so both
// the raw reads (along x) and the image writes (along j) are contiguous.
__global__ void __launch_bounds__(256)
finalize_kernel(const unsigned long long* __restrict__ sum, const unsigned int* __restrict__ mx,
                unsigned long long n_total, int W, int H, unsigned int scale, int rotated,
                uint16_t* __restrict__ mean_img, uint16_t* __restrict__ max_img) {
    __shared__ uint16_t t_mean[32][33];
    __shared__ uint16_t t_max[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    for (int r = ty; r < 32; r += 8) {
        const int x = x0 + tx, y = y0 + r;
        uint16_t a = 0, b = 0;
        if (x < W && y < H) {
            const int64_t idx = (int64_t)y * W + x;
            a = (uint16_t)((sum[idx] * scale) / n_total);
            b = (uint16_t)(mx[idx] * scale);
            if (!rotated) {
                mean_img[idx] = a;
                max_img[idx] = b;
            }
        }
        t_mean[r][tx] = a;
        t_max[r][tx] = b;
    }
    if (!rotated) return;
    __syncthreads();
    // image row i = W-1-x has H contiguous entries j = y
    for (int c = ty; c < 32; c += 8) {
        const int x = x0 + c, y = y0 + tx;
        if (x < W && y < H) {
            const int64_t o = (int64_t)(W - 1 - x) * H + y;
            mean_img[o] = t_mean[tx][c];
            max_img[o] = t_max[tx][c];
        }
    }
}

// per-frame sum of the raw pixels (one CTA per frame): np.mean(frame) of all_video_reader.means
template <typename T>
__global__ void __launch_bounds__(256)
frame_sums_kernel(const T* __restrict__ frames, int64_t frame_px, unsigned long long* __restrict__ out) {
    const T* f = frames + (int64_t)blockIdx.x * frame_px;
    unsigned long long s = 0;
    for (int64_t i = threadIdx.x; i < frame_px; i += 256) s += f[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ unsigned long long part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += part[w];
        out[blockIdx.x] = t;
    }
}

}  // namespace

extern "C" int shg_frame_sums(const void* d_frames, int bytes_per_px, int64_t n_frames, int64_t frame_px,
                              uint64_t* d_out, void* stream) {
    SHG_REQUIRE(bytes_per_px == 1 || bytes_per_px == 2, "shg_frame_sums: bytes_per_px must be 1 or 2");
    if (n_frames <= 0 || frame_px <= 0) return 0;
    SHG_REQUIRE(n_frames <= 0x7fffffff, "shg_frame_sums: too many frames");
    auto* out = reinterpret_cast<unsigned long long*>(d_out);
    if (bytes_per_px == 2)
        frame_sums_kernel<uint16_t><<<(unsigned)n_frames, 256, 0, as_stream(stream)>>>((const uint16_t*)d_frames, frame_px, out);
    else
        frame_sums_kernel<uint8_t><<<(unsigned)n_frames, 256, 0, as_stream(stream)>>>((const uint8_t*)d_frames, frame_px, out);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_accumulate(const void* d_frames, int bytes_per_px, int64_t n_frames, int64_t frame_px,
                              uint64_t* d_sum, uint32_t* d_max, void* stream) {
    SHG_REQUIRE(bytes_per_px == 1 || bytes_per_px == 2, "shg_accumulate: bytes_per_px must be 1 or 2");
    if (n_frames <= 0 || frame_px <= 0) return 0;
    const int64_t frame_bytes = frame_px * bytes_per_px;
    const bool vec = (frame_bytes % 16 == 0) && ((uintptr_t)d_frames % 16 == 0);
    const int64_t items = vec ? frame_bytes / 16 : frame_px;
    const int64_t bx = ceil_div64(items, kThreads);
    // enough CTAs for ~2 waves of 8 CTAs/SM; splits also cap 32-bit partial sums
    int64_t splits = ceil_div64((int64_t)SHG_SM_COUNT_B200 * 16, bx);
    splits = max((int64_t)1, min(splits, ceil_div64(n_frames, 16)));
    int64_t fps = ceil_div64(n_frames, splits);
    fps = min(fps, (int64_t)65536);
    splits = ceil_div64(n_frames, fps);
    SHG_REQUIRE(splits <= 65535 && bx <= 0x7fffffff, "shg_accumulate: launch too large");
    dim3 grid((unsigned)bx, (unsigned)splits);
    auto* sum = reinterpret_cast<unsigned long long*>(d_sum);
    cudaStream_t st = as_stream(stream);
    if (vec && bytes_per_px == 2)
        accumulate_u16_kernel<<<grid, kThreads, 0, st>>>((const uint4*)d_frames, n_frames, items, fps, sum, d_max);
    else if (vec)
        accumulate_u8_kernel<<<grid, kThreads, 0, st>>>((const uint4*)d_frames, n_frames, items, fps, sum, d_max);
    else if (bytes_per_px == 2)
        accumulate_generic_kernel<uint16_t><<<grid, kThreads, 0, st>>>((const uint16_t*)d_frames, n_frames, frame_px, fps, sum, d_max);
    else
        accumulate_generic_kernel<uint8_t><<<grid, kThreads, 0, st>>>((const uint8_t*)d_frames, n_frames, frame_px, fps, sum, d_max);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_finalize_mean_max(const uint64_t* d_sum, const uint32_t* d_max, int64_t n_total,
                                     int W, int H, int eight_bit, uint16_t* d_mean_img, uint16_t* d_max_img,
                                     void* stream) {
    SHG_REQUIRE(n_total > 0 && W > 0 && H > 0, "shg_finalize_mean_max: bad geometry");
    dim3 grid((W + 31) / 32, (H + 31) / 32);
    finalize_kernel<<<grid, 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const unsigned long long*>(d_sum), d_max, (unsigned long long)n_total, W, H,
        eight_bit ? 256u : 1u, W > H ? 1 : 0, d_mean_img, d_max_img);
    SHG_LAUNCH_CHECK();
    return 0;
}
