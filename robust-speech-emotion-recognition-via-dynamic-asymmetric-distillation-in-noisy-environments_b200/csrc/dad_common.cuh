// Common device/host helpers for libdad_sm100a.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "dad_sm100a.h"
#include <stdio.h>
#include <stdarg.h>

#define DAD_OK 0
#define DAD_ERR_ARG (-1)
#define DAD_ERR_CUDA (-2)
#define DAD_ERR_WS (-3)
#define DAD_ERR_UNSUPPORTED (-4)

int dad_set_error(int code, const char* fmt, ...);  // returns code

#define DAD_REQUIRE(cond, ...)                                    \
    do {                                                          \
        if (!(cond)) return dad_set_error(DAD_ERR_ARG, __VA_ARGS__); \
    } while (0)

#define DAD_CHECK_LAUNCH(name)                                                          \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess)                                                         \
            return dad_set_error(DAD_ERR_CUDA, "%s launch: %s", name, cudaGetErrorString(e__)); \
    } while (0)

// DAD_D = 768, DAD_H = 256, DAD_C = 4 come from dad_sm100a.h
constexpr int DAD_HW = DAD_H / 32;  // ReLU bitmask words per frame

static __host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static __host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// Augmentation arithmetic shared by the materialising kernel and the fused encoder producers.
// Reference: I/utils.py:328-375.  x + (n*std) is two fp32 roundings: no FMA contraction.
// ---------------------------------------------------------------------------------------------
typedef dad_aug AugParams;  // see include/dad_sm100a.h

__device__ __forceinline__ unsigned long long aug_seed(const dad_aug& ap) {
    return ap.seed + (ap.step_ptr ? (unsigned long long)(*ap.step_ptr) * 0x9E3779B97F4A7C15ull : 0ull);
}
__device__ __forceinline__ float aug_add_noise(float x, float n, float std) { return __fadd_rn(x, __fmul_rn(n, std)); }

// Philox4x32-10 (Salmon et al. 2011), counter = (idx4 lo, idx4 hi, stream lo, stream hi), key = seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
// four unit normals for the aligned group of 4 consecutive elements starting at flat index 4*idx4
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long stream, unsigned long long idx4) {
    uint4 r = philox4x32_10(make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float k = 2.3283064365386963e-10f;  // 2^-32
    float u0 = fmaf((float)r.x, k, k * 0.5f), u1 = (float)r.y * k;
    float u2 = fmaf((float)r.z, k, k * 0.5f), u3 = (float)r.w * k;
    float r0 = sqrtf(-2.0f * __logf(u0)), r1 = sqrtf(-2.0f * __logf(u2));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u1, &s0, &c0);
    __sincosf(6.283185307179586f * u3, &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}
