// Shared helpers for libshg (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "shg.h"

#define SHG_SM_COUNT_B200 148

void shg_set_error(const char* fmt, ...);

#define SHG_CHECK(expr)                                                              \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            shg_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,              \
                          cudaGetErrorString(_e));                                   \
            return 1;                                                                \
        }                                                                            \
    } while (0)

#define SHG_REQUIRE(cond, ...)                                                       \
    do {                                                                             \
        if (!(cond)) {                                                               \
            shg_set_error(__VA_ARGS__);                                              \
            return 2;                                                                \
        }                                                                            \
    } while (0)

#define SHG_LAUNCH_CHECK() SHG_CHECK(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// 128-bit streaming load that does not allocate in L1 (data is touched once).
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// uint16 -> double without the conversion pipe: 2^52 + v has v in its low
// mantissa bits; subtracting 2^52 is exact.
__device__ __forceinline__ double u32_to_double(uint32_t v) {
    return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}

// trunc for 0 <= v < 2^31 without the conversion pipe: add 2^52 rounding
// toward -inf, the integer sits in the low word.
__device__ __forceinline__ uint32_t double_floor_to_u32(double v) {
    return (uint32_t)__double2loint(__dadd_rd(v, 4503599627370496.0));
}
