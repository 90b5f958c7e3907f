// Layout helpers: frame-major disk <-> reference (ih, N) layout, min/max.
// The reference builds disk images as (ih, N) arrays with the frame index
// fastest (solex_util.py:96-97,134) and flips them with np.flip(axis=1)
// (Solex_recon.py:75-76); the device keeps them frame-major (N, ih).
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kTile = 64;

// out[c][r'] = in[r][c]; 64x64 tiles, 32-bit global accesses on both sides.
__global__ void __launch_bounds__(256)
transpose_u16_kernel(const uint16_t* __restrict__ in, int64_t rows, int64_t cols, uint16_t* __restrict__ out,
                     int flip) {
    __shared__ uint16_t tile[kTile][kTile + 2];
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;      // 32 x 8
    const int64_t r0 = (int64_t)blockIdx.y * kTile, c0 = (int64_t)blockIdx.x * kTile;
    const bool vec_in = (cols % 2 == 0), vec_out = (rows % 2 == 0) && !flip;
    for (int r = wy; r < kTile; r += 8) {
        const int64_t rr = r0 + r, cc = c0 + 2 * lane;
        if (rr >= rows) continue;
        const uint16_t* p = in + rr * cols + cc;
        if (vec_in && cc + 1 < cols) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
            tile[r][2 * lane] = (uint16_t)(v & 0xffffu);
            tile[r][2 * lane + 1] = (uint16_t)(v >> 16);
        } else {
            if (cc < cols) tile[r][2 * lane] = p[0];
            if (cc + 1 < cols) tile[r][2 * lane + 1] = p[1];
        }
    }
    __syncthreads();
    for (int c = wy; c < kTile; c += 8) {
        const int64_t cc = c0 + c;
        if (cc >= cols) continue;
        const int64_t ra = r0 + 2 * lane;
        if (vec_out) {
            if (ra + 1 < rows) {
                const uint32_t v = (uint32_t)tile[2 * lane][c] | ((uint32_t)tile[2 * lane + 1][c] << 16);
                *reinterpret_cast<uint32_t*>(out + cc * rows + ra) = v;
            } else if (ra < rows) {
                out[cc * rows + ra] = tile[2 * lane][c];
            }
        } else {
            for (int j = 0; j < 2; ++j) {
                const int64_t rr = ra + j;
                if (rr < rows) out[cc * rows + (flip ? rows - 1 - rr : rr)] = tile[2 * lane + j][c];
            }
        }
    }
}

// min / max of every image of a batch (blockIdx.y = image)
__global__ void __launch_bounds__(256)
minmax_u16_kernel(const uint16_t* __restrict__ base, int64_t n, int64_t img_stride, const int32_t* __restrict__ sel,
                  unsigned int* __restrict__ out /* [n_imgs][2], preset to {65535, 0} */) {
    const uint16_t* in = base + (int64_t)(sel ? sel[blockIdx.y] : blockIdx.y) * img_stride;
    unsigned int* out2 = out + 2 * blockIdx.y;
    unsigned int lo = 0xffffu, hi = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nvec = ((uintptr_t)in % 16 == 0) ? n / 8 : 0;
    const uint4* v4 = reinterpret_cast<const uint4*>(in);
    uint32_t mn = 0xffffffffu, mx = 0;
    for (int64_t v = i; v < nvec; v += stride) {
        const uint4 q = ld_stream_u4(v4 + v);
        mn = __vminu2(mn, __vminu2(__vminu2(q.x, q.y), __vminu2(q.z, q.w)));
        mx = __vmaxu2(mx, __vmaxu2(__vmaxu2(q.x, q.y), __vmaxu2(q.z, q.w)));
    }
    lo = min(mn & 0xffffu, mn >> 16);
    hi = max(mx & 0xffffu, mx >> 16);
    for (int64_t j = nvec * 8 + i; j < n; j += stride) {
        lo = min(lo, (unsigned int)in[j]);
        hi = max(hi, (unsigned int)in[j]);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out2, lo);
        atomicMax(out2 + 1, hi);
    }
}

__global__ void minmax_init_kernel(unsigned int* out, int n_imgs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_imgs) { out[2 * i] = 0xffffu; out[2 * i + 1] = 0; }
}

// position-sensitive checksum of one image: sum over pixels of (v + 1) * odd(mix(index)) mod 2^64.
// Integer and order-independent, so every launch geometry (and every sharding of the work that produced
// the image) gives the same value for the same pixels.
__global__ void __launch_bounds__(256)
checksum_u16_kernel(const uint16_t* __restrict__ in, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        unsigned long long m = (unsigned long long)(i + 1) * 0x9E3779B97F4A7C15ull;
        m ^= m >> 29;
        acc += ((unsigned long long)in[i] + 1ull) * (m | 1ull);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

}  // namespace

extern "C" int shg_checksum_u16(const uint16_t* d_in, int64_t n, uint64_t* d_out, void* stream) {
    SHG_REQUIRE(d_out && (d_in || n == 0) && n >= 0, "shg_checksum_u16: bad arguments");
    SHG_CHECK(cudaMemsetAsync(d_out, 0, 8, as_stream(stream)));
    if (n == 0) return 0;
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n, 256 * 8), SHG_SM_COUNT_B200 * 8));
    checksum_u16_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_in, n, reinterpret_cast<unsigned long long*>(d_out));
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_transpose_u16(const uint16_t* d_in, int64_t rows, int64_t cols, uint16_t* d_out, int flip,
                                 void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    const int64_t gx = ceil_div64(cols, kTile), gy = ceil_div64(rows, kTile);
    SHG_REQUIRE(gy <= 65535 && gx <= 0x7fffffff, "shg_transpose_u16: image too large (%lld x %lld)", (long long)rows,
                (long long)cols);
    transpose_u16_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, as_stream(stream)>>>(d_in, rows, cols, d_out, flip);
    SHG_LAUNCH_CHECK();
    return 0;
}

extern "C" int shg_minmax_u16(const uint16_t* d_in, int64_t n, int64_t img_stride, const int32_t* d_sel, int n_imgs,
                              uint32_t* d_out, void* stream) {
    if (n <= 0 || n_imgs <= 0) return 0;
    SHG_REQUIRE(n_imgs <= 65535, "shg_minmax_u16: too many images");
    cudaStream_t st = as_stream(stream);
    minmax_init_kernel<<<(n_imgs + 255) / 256, 256, 0, st>>>(d_out, n_imgs);
    SHG_LAUNCH_CHECK();
    const int64_t want = std::max<int64_t>(1, (int64_t)SHG_SM_COUNT_B200 * 8 / n_imgs);
    const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n, 256 * 8), std::max<int64_t>(want, 8)));
    minmax_u16_kernel<<<dim3((unsigned)blocks, n_imgs), 256, 0, st>>>(d_in, n, img_stride, d_sel, d_out);
    SHG_LAUNCH_CHECK();
    return 0;
}
