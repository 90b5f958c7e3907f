// Synthetic spectroheliograph scans generated directly in HBM (bench and
// full-size property tests: the 84 GB config-5 stack never exists on a host).
// Same recipe family as solex_ser_recon_en_b200/synth.py: a dark
// absorption line bending quadratically along the slit, times the limb-darkened
// solar disk at this frame's slit position, plus pedestal and noise.  The noise
// is a counter-based hash (sum of four uniform bytes), not NumPy's generator, so
// these scans are their own family: parity tests run the oracle on frames read
// back from the device.
#include <algorithm>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

struct SynthParams {
    int W, H, rotated, n_slit, n_disp;
    float sigma, depth, amplitude, pedestal, noise_scale, top;
    float ex, ey;          // disk semi-axes as fractions of n_total / n_slit
    uint32_t seed;
};

__device__ __forceinline__ float synth_pixel(const SynthParams& p, int64_t k, int64_t n_total, int x, int y) {
    const int s = p.rotated ? x : y;
    const int d = p.rotated ? y : x;
    const float hs = 0.5f * p.n_slit;
    const float u = ((float)s - hs) / hs;
    const float centre = 0.5f * p.n_disp + 6.0f * u * u + 1.5f * u;
    const float z = ((float)d - centre) / p.sigma;
    const float profile = 1.0f - p.depth * __expf(-0.5f * z * z);
    const float a = ((float)k - 0.5f * (float)n_total) / (p.ex * (float)n_total);
    const float b = ((float)s - hs) / (p.ey * (float)p.n_slit);
    const float rho2 = a * a + b * b;
    const float disk = rho2 < 1.0f ? sqrtf(fmaxf(1.0f - 0.6f * rho2, 0.0f)) : 0.02f;
    const float tex = 1.0f + 0.05f * __sinf(0.37f * (float)(k % 4096)) * __cosf(0.11f * (float)s);
    const float dust = (s % 997) == 400 ? 0.97f : 1.0f;
    const uint32_t h = mix32(mix32((uint32_t)k * 0x9e3779b9u + p.seed) ^ (uint32_t)(y * p.W + x));
    const float noise = ((float)__dp4a(h, 0x01010101u, 0u) - 510.0f) * p.noise_scale;
    const float v = p.amplitude * disk * tex * dust * profile + p.pedestal + noise;
    return fminf(fmaxf(rintf(v), 0.0f), p.top);
}

template <typename T>
__global__ void __launch_bounds__(256)
synth_kernel(T* __restrict__ frames, int64_t k0, int64_t n, int64_t n_total, const SynthParams p) {
    constexpr int PX = 16 / sizeof(T);
    const int64_t frame_px = (int64_t)p.W * p.H;
    const int64_t groups = (frame_px + PX - 1) / PX;
    for (int64_t f = blockIdx.y; f < n; f += gridDim.y) {
        T* fr = frames + f * frame_px;
        for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < groups; g += (int64_t)gridDim.x * 256) {
            const int64_t base = g * PX;
            T v[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                const int64_t px = base + j;
                const int y = (int)(px / p.W), x = (int)(px % p.W);
                v[j] = px < frame_px ? (T)synth_pixel(p, k0 + f, n_total, x, y) : (T)0;
            }
            if (base + PX <= frame_px && ((uintptr_t)(fr + base) % 16) == 0) {
                *reinterpret_cast<uint4*>(fr + base) = *reinterpret_cast<const uint4*>(v);
            } else {
                for (int j = 0; j < PX && base + j < frame_px; ++j) fr[base + j] = v[j];
            }
        }
    }
}

}  // namespace

extern "C" int shg_synth_fill(void* d_frames, int bytes_per_px, int64_t k0, int64_t n, int64_t n_total,
                              int W, int H, uint64_t seed, void* stream) {
    SHG_REQUIRE(bytes_per_px == 1 || bytes_per_px == 2, "shg_synth_fill: bytes_per_px must be 1 or 2");
    SHG_REQUIRE(W > 0 && H > 0 && n_total > 0, "shg_synth_fill: bad geometry");
    if (n <= 0) return 0;
    SynthParams p;
    p.W = W; p.H = H; p.rotated = W > H ? 1 : 0;
    p.n_slit = p.rotated ? W : H;
    p.n_disp = p.rotated ? H : W;
    if (bytes_per_px == 2) { p.sigma = 3.0f; p.depth = 0.75f; p.amplitude = 30000.0f; p.pedestal = 300.0f; p.noise_scale = 50.0f / 147.8f; p.top = 65535.0f; }
    else { p.sigma = 8.0f; p.depth = 0.9f; p.amplitude = 110.0f; p.pedestal = 4.0f; p.noise_scale = 1.2f / 147.8f; p.top = 255.0f; }
    p.ex = 0.42f; p.ey = 0.40f;
    p.seed = (uint32_t)(seed * 0x9e3779b97f4a7c15ull >> 32) ^ (uint32_t)seed;
    const int64_t frame_px = (int64_t)W * H;
    const int px = 16 / bytes_per_px;
    const unsigned gx = (unsigned)std::min<int64_t>(ceil_div64(ceil_div64(frame_px, px), 256), 4096);
    const unsigned gy = (unsigned)std::min<int64_t>(n, 65535);
    if (bytes_per_px == 2)
        synth_kernel<uint16_t><<<dim3(gx, gy), 256, 0, as_stream(stream)>>>((uint16_t*)d_frames, k0, n, n_total, p);
    else
        synth_kernel<uint8_t><<<dim3(gx, gy), 256, 0, as_stream(stream)>>>((uint8_t*)d_frames, k0, n, n_total, p);
    SHG_LAUNCH_CHECK();
    return 0;
}
