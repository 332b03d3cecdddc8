// Materialising augmentation kernel: weak / strong views of a [B,T,D] feature batch in one coalesced pass.
// Reference semantics: I/utils.py:328-375 (DataAugmentation.weak_augment / strong_augment / _apply_temporal_masking).
// HBM-bound: reads x (+ supplied noise), writes out; float4 vectorised, grid-stride, no shared memory (no reuse).
#include "dad_common.cuh"

template <bool kPhilox>
__global__ void __launch_bounds__(256) k_augment(const float* __restrict__ x, float* __restrict__ out, AugParams ap,
                                                 int64_t n_vec, int T, int D) {
    const int dv = D >> 2;  // float4 per frame
    const unsigned long long seed = kPhilox ? aug_seed(ap) : 0ull;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / dv;  // b*T + t
        const int d4 = (int)(i - row * dv);
        float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        float4 n;
        if (kPhilox) n = philox_normal4(seed, ap.stream, (unsigned long long)i, (unsigned long long)ap.frame_offset);
        else n = __ldg(reinterpret_cast<const float4*>(ap.noise) + i);
        v.x = aug_add_noise(v.x, n.x, ap.sigma);
        v.y = aug_add_noise(v.y, n.y, ap.sigma);
        v.z = aug_add_noise(v.z, n.z, ap.sigma);
        v.w = aug_add_noise(v.w, n.w, ap.sigma);
        if (ap.mode == 2) {
            if (ap.feat_keep) {
                const float4 k = __ldg(reinterpret_cast<const float4*>(ap.feat_keep) + d4);
                v.x = __fmul_rn(v.x, k.x); v.y = __fmul_rn(v.y, k.y); v.z = __fmul_rn(v.z, k.z); v.w = __fmul_rn(v.w, k.w);
            }
            if (ap.t_start && ap.mask_len > 0) {
                const int b = (int)(row / T), t = (int)(row - (int64_t)b * T);
                const int s = __ldg(ap.t_start + b);
                if (t >= s && t < s + ap.mask_len) v = make_float4(0.f, 0.f, 0.f, 0.f);  // assignment of +0 (utils.py:373)
            }
        }
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

int launch_augment(const float* x, float* out, const AugParams& ap, int64_t B, int64_t T, int64_t D, cudaStream_t st) {
    const int64_t n_vec = B * T * D / 4;
    if (n_vec == 0) return DAD_OK;
    int sms = 148;
    int64_t blocks = ceil_div64(n_vec, 256);
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
    if (ap.noise_src == 1) k_augment<true><<<(unsigned)blocks, 256, 0, st>>>(x, out, ap, n_vec, (int)T, (int)D);
    else k_augment<false><<<(unsigned)blocks, 256, 0, st>>>(x, out, ap, n_vec, (int)T, (int)D);
    DAD_CHECK_LAUNCH("k_augment");
    return DAD_OK;
}

// Device-side batch assembly (the reference's collator, I/dataload_clean.py:177-193, as one HBM-bound pass):
// utterance b of the batch = rows [offsets[b], offsets[b] + sizes[b]) of the split's resident [sum T, D] feature array; the output is
// the zero-padded [B, T, D] batch and its padding mask (1 = pad).  One float4 per thread, grid-stride, fully coalesced.
__global__ void __launch_bounds__(256) k_gather_pad(const float* __restrict__ feats, const int64_t* __restrict__ offsets, const int32_t* __restrict__ sizes,
                                                    float* __restrict__ x, uint8_t* __restrict__ pad, int64_t n_vec, int T, int D) {
    const int dv = D >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / dv;  // b*T + t
        const int d4 = (int)(i - row * dv);
        const int b = (int)(row / T), t = (int)(row - (int64_t)b * T);
        const bool valid = t < __ldg(sizes + b);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) v = __ldg(reinterpret_cast<const float4*>(feats + (__ldg(offsets + b) + t) * (int64_t)D) + d4);
        reinterpret_cast<float4*>(x)[i] = v;
        if (d4 == 0 && pad) pad[row] = valid ? 0 : 1;
    }
}

int launch_gather_pad(const float* feats, const int64_t* offsets, const int32_t* sizes, float* x, uint8_t* pad, int64_t B, int64_t T, int64_t D,
                      cudaStream_t st) {
    const int64_t n_vec = B * T * D / 4;
    if (n_vec == 0) return DAD_OK;
    int64_t blocks = ceil_div64(n_vec, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_gather_pad<<<(unsigned)blocks, 256, 0, st>>>(feats, offsets, sizes, x, pad, n_vec, (int)T, (int)D);
    DAD_CHECK_LAUNCH("k_gather_pad");
    return DAD_OK;
}

// Per-step draws of the throughput mode (dropout keep masks, feature keep mask, temporal span starts).
__global__ void __launch_bounds__(256) k_step_rng(unsigned long long seed, const int64_t* step_ptr, float* keep_clean, float* keep_strong,
                                                  int64_t n_keep, float p_drop, float* feat_keep, float feat_rate, int32_t* t_start,
                                                  int64_t B, int64_t T, int mask_len, int64_t sample_offset) {
    const unsigned long long s = seed + (step_ptr ? (unsigned long long)(*step_ptr) * 0x9E3779B97F4A7C15ull : 0ull);
    const float k = 2.3283064365386963e-10f;
    const int64_t n4 = (n_keep + 3) / 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        // per-sample draws are indexed by the GLOBAL sample (sample_offset = rank * B_local under data parallelism): ranks draw
        // disjoint parts of one field, as the reference's single nn.Dropout over the whole batch does; n_keep = B * 256 is a multiple of 4
        const int64_t gi = i + sample_offset * (DAD_H / 4);
        const uint4 r = philox4x32_10(make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), 101u, 0u), make_uint2((uint32_t)s, (uint32_t)(s >> 32)));
        const uint4 q = philox4x32_10(make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), 102u, 0u), make_uint2((uint32_t)s, (uint32_t)(s >> 32)));
        const uint32_t rv[4] = {r.x, r.y, r.z, r.w}, qv[4] = {q.x, q.y, q.z, q.w};
        for (int j = 0; j < 4; ++j) {
            const int64_t e = i * 4 + j;
            if (e >= n_keep) break;
            if (keep_clean) keep_clean[e] = ((float)rv[j] * k >= p_drop) ? 1.f : 0.f;
            if (keep_strong) keep_strong[e] = ((float)qv[j] * k >= p_drop) ? 1.f : 0.f;
        }
    }
    if (blockIdx.x == 0) {
        for (int64_t i = threadIdx.x; i < (DAD_D + 3) / 4; i += blockDim.x) {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)i, 0u, 103u, 0u), make_uint2((uint32_t)s, (uint32_t)(s >> 32)));
            const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
            for (int j = 0; j < 4; ++j)
                if (feat_keep && i * 4 + j < DAD_D) feat_keep[i * 4 + j] = ((float)rv[j] * k > feat_rate) ? 1.f : 0.f;  // rand > rate (utils.py:343)
        }
        const int64_t span = (T - mask_len + 1) > 1 ? (T - mask_len + 1) : 1;
        for (int64_t b = threadIdx.x; t_start && b < B; b += blockDim.x) {
            const int64_t gb = b + sample_offset;
            const uint4 r = philox4x32_10(make_uint4((uint32_t)gb, (uint32_t)(gb >> 32), 104u, 0u), make_uint2((uint32_t)s, (uint32_t)(s >> 32)));
            t_start[b] = (int32_t)(((unsigned long long)r.x * (unsigned long long)span) >> 32);
        }
    }
}

int launch_step_rng(unsigned long long seed, const int64_t* step_ptr, float* keep_clean, float* keep_strong, int64_t B, float p_drop,
                    float* feat_keep, float feat_rate, int32_t* t_start, int64_t T, int mask_len, int64_t sample_offset, cudaStream_t st) {
    const int64_t n_keep = B * DAD_H;
    int64_t blocks = ceil_div64((n_keep + 3) / 4, 256);
    if (blocks < 1) blocks = 1;
    if (blocks > 592) blocks = 592;
    k_step_rng<<<(unsigned)blocks, 256, 0, st>>>(seed, step_ptr, keep_clean, keep_strong, n_keep, p_drop, feat_keep, feat_rate, t_start, B, T, mask_len, sample_offset);
    DAD_CHECK_LAUNCH("k_step_rng");
    return DAD_OK;
}
