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

// Power-of-two operand scales of the tcgen05 "f16x3" path (exact in fp32; see k_enc_tc.cu): features x 2^3, encoder weight x 2^8.
constexpr float TC_XS = 8.0f;
constexpr float TC_WS = 256.0f;
constexpr float TC_INV_XW = 1.0f / (TC_XS * TC_WS);

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

// Philox4x32 (Salmon et al. 2011), counter = (idx4 lo, idx4 hi, stream lo, stream hi), key = seed.  7 rounds: the smallest count
// that passes BigCrush in the paper (their Table 2; cuRAND's default of 10 adds safety margin); this generator only feeds the
// augmentation noise, and the converters of the noisy forward views are instruction-issue bound (ncu: 61 % issue active, 52 %
// tensor active against 74 % for the clean view), so the 24 integer instructions per 8 normals matter.
constexpr int PHILOX_ROUNDS = 7;
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < PHILOX_ROUNDS; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
// ---------------------------------------------------------------------------------------------
// Canonical in-kernel noise field (throughput mode; the parity mode reads supplied noise instead).
// One Philox4x32 call -- counter (n, c4, stream >> 1, 0), key = seed -- yields the four 32-bit words of the float4 column
// group c4 (columns 4*c4 .. 4*c4+3, c4 < 192) of frame n.  The LOW BIT of the stream id selects the word pair: even streams
// take words (x, y), odd streams (z, w) -- so the weak (stream 2) and strong (stream 3) views of the same noisy batch share
// every Philox block, and the fused weak+strong forward (k_enc_fwd_tc, two views from one read of x) pays one Philox call per
// float4 for both.  Each word feeds one Box-Muller pair with a 16-bit radius uniform u = (k + 0.5) / 65536 and a 16-bit
// angle (|z| <= 4.86; the truncated tail mass is 1.2e-6): first word -> columns 4*c4, 4*c4+1 (r cos, r sin), second word ->
// columns 4*c4+2, 4*c4+3.  Every kernel that needs the noise (materialising augment, CUDA-core and tcgen05 encoder forward
// AND backward) regenerates it from these functions, so the field is identical wherever it is consumed.
// radius_k: -2 ln2 * sigma^2 folds the noise scale into the radius (r = sigma * sqrt(-2 ln u)); pass -2 ln2 for unit normals.
// ---------------------------------------------------------------------------------------------
constexpr float NOISE_UNIT_K = -1.3862943611198906f;  // -2 ln 2
// flush-to-zero MUFU forms: one instruction each (the non-ftz forms add a denormal-rescue FSETP / FMUL pair per call, and the
// arguments here are normal by construction)
__device__ __forceinline__ float fast_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void fast_sincos(float x, float& s, float& c) {
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(x));
}
// n = GLOBAL frame index (ap.frame_offset + local frame: under data parallelism every rank draws its own part of one field)
__device__ __forceinline__ uint4 noise_words(unsigned long long seed, uint32_t stream_block, unsigned long long n, uint32_t c4) {
    return philox4x32_10(make_uint4((uint32_t)n, c4 | ((uint32_t)(n >> 32) << 8), stream_block, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
// one word -> (radius, cos, sin): the two normals are r*c and r*s
__device__ __forceinline__ void bm_word(uint32_t w, float radius_k, float& r, float& c, float& s) {
    const float u = __uint_as_float(0x3f800040u | ((w & 0xffffu) << 7)) - 1.0f;  // (k + 0.5) / 65536 in (0, 1)
    r = fast_sqrt(fast_lg2(u) * radius_k);
    const float turns = __uint_as_float(0x3f800000u | ((w >> 16) << 7));         // 1 + a / 65536
    fast_sincos(fmaf(turns, 6.283185307179586f, -9.42477796076938f), s, c);      // angle = 2 pi a / 65536 - pi in [-pi, pi)
}
struct NoiseBM4 { float r[2], c[2], s[2]; };
__device__ __forceinline__ void bm4_from_words(uint32_t wa, uint32_t wb, float radius_k, NoiseBM4& o) {
    bm_word(wa, radius_k, o.r[0], o.c[0], o.s[0]);
    bm_word(wb, radius_k, o.r[1], o.c[1], o.s[1]);
}
// x + noise for one float4 (single rounding per element: the radius already carries sigma)
__device__ __forceinline__ float4 bm4_add(const NoiseBM4& z, const float4& v) {
    return make_float4(fmaf(z.r[0], z.c[0], v.x), fmaf(z.r[0], z.s[0], v.y), fmaf(z.r[1], z.c[1], v.z), fmaf(z.r[1], z.s[1], v.w));
}
__device__ __forceinline__ void philox_bm4(unsigned long long seed, uint32_t stream, unsigned long long n, uint32_t c4, float radius_k, NoiseBM4& o) {
    const uint4 w = noise_words(seed, stream >> 1, n, c4);
    if (stream & 1u) bm4_from_words(w.z, w.w, radius_k, o);
    else bm4_from_words(w.x, w.y, radius_k, o);
}
// four unit normals for the aligned group of 4 consecutive elements starting at flat index 4*idx4 of a [N, 768] tensor whose
// first frame is frame `frame_offset` of the global field
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long stream, unsigned long long idx4,
                                                 unsigned long long frame_offset) {
    const unsigned long long n = idx4 / (DAD_D / 4);
    const uint32_t c4 = (uint32_t)(idx4 - n * (DAD_D / 4));
    NoiseBM4 z;
    philox_bm4(seed, (uint32_t)stream, n + frame_offset, c4, NOISE_UNIT_K, z);
    return make_float4(z.r[0] * z.c[0], z.r[0] * z.s[0], z.r[1] * z.c[1], z.r[1] * z.s[1]);
}
