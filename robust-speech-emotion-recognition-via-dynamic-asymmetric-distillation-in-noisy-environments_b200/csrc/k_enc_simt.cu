// Exact-fp32 (CUDA-core FMA) encoder kernels: the generic / precise path.
//   forward : emb[b,:] = sum_t valid(b,t) * relu(W x[b,t,:] + bias) / max(len_b,1)     (I/model.py:18-41)
//   backward: dW = sum_n dpre[n,:] (x) x[n,:],  db = sum_n dpre[n,:],
//             dpre[n,h] = [pre>0] * valid(n) * g[b(n),h]/len_b                           (autograd of the above)
// Both fuse the weak/strong augmentation of I/utils.py:328-375 into the tile loader, never materialise the
// [N,256] activations (only a 1-bit/element ReLU mask survives the forward) and are deterministic
// (fixed-order partial sums, no fp32 atomics).  Handles any B, T >= 1.  The tcgen05 path (k_enc_tc.cu)
// is the fast path for T >= 32; this one is also its on-device cross-check.
#include "dad_common.cuh"

namespace {

constexpr int F_BM = 64, F_BN = 128, F_BK = 16, F_THREADS = 256;

__device__ __forceinline__ float4 load_x_aug(const float* __restrict__ x, const AugParams& ap, unsigned long long seed,
                                             int64_t e /*flat elem idx, %4==0*/, int b, int t, int k) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x + e));
    if (ap.mode == 0) return v;
    float4 n;
    if (ap.noise_src == 1) n = philox_normal4(seed, ap.stream, (unsigned long long)(e >> 2), (unsigned long long)ap.frame_offset);
    else n = __ldg(reinterpret_cast<const float4*>(ap.noise + e));
    v.x = aug_add_noise(v.x, n.x, ap.sigma);
    v.y = aug_add_noise(v.y, n.y, ap.sigma);
    v.z = aug_add_noise(v.z, n.z, ap.sigma);
    v.w = aug_add_noise(v.w, n.w, ap.sigma);
    if (ap.mode == 2) {
        if (ap.feat_keep) {
            const float4 kp = __ldg(reinterpret_cast<const float4*>(ap.feat_keep + k));
            v.x = __fmul_rn(v.x, kp.x); v.y = __fmul_rn(v.y, kp.y); v.z = __fmul_rn(v.z, kp.z); v.w = __fmul_rn(v.w, kp.w);
        }
        if (ap.t_start && ap.mask_len > 0) {
            const int s = __ldg(ap.t_start + b);
            if (t >= s && t < s + ap.mask_len) v = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    return v;
}

// grid = (ceil(T/64), B, 2); one CTA = 64 frames of one utterance x 128 hidden units.
__global__ void __launch_bounds__(F_THREADS) k_enc_fwd_simt(const float* __restrict__ x, const uint8_t* __restrict__ pad,
                                                           const float* __restrict__ W, const float* __restrict__ bias, AugParams ap,
                                                           float* __restrict__ part, uint32_t* __restrict__ bits,
                                                           float* __restrict__ emb, float* __restrict__ lens_out,
                                                           unsigned* __restrict__ counters, int B, int T) {
    __shared__ __align__(16) float As[2][F_BK][F_BM + 4];
    __shared__ __align__(16) float Bs[2][F_BK][F_BN + 4];
    __shared__ float red[16][F_BN];
    __shared__ float s_len;
    __shared__ int s_last;

    const int tid = threadIdx.x;
    const unsigned long long seed = (ap.mode && ap.noise_src == 1) ? aug_seed(ap) : 0ull;
    const int tile_t = blockIdx.x, b = blockIdx.y, ct = blockIdx.z;
    const int tiles_t = gridDim.x;
    const int t0 = tile_t * F_BM, col0 = ct * F_BN;

    // loader coordinates
    const int a_row = tid >> 2, a_kq = tid & 3;          // A: 64 rows x 4 float4
    const int b_col = tid >> 2, b_kq = tid & 3;          // B: 128 cols x 4 float4 (two passes: col, col+64)
    const int a_t = t0 + a_row;
    const bool a_in = a_t < T;
    const int64_t a_base = ((int64_t)b * T + a_t) * DAD_D;

    float4 ra, rb0, rb1;
    auto gload = [&](int kb) {
        const int k = kb * F_BK + a_kq * 4;
        ra = a_in ? load_x_aug(x, ap, seed, a_base + k, b, a_t, k) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int kw = kb * F_BK + b_kq * 4;
        rb0 = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(col0 + b_col) * DAD_D + kw));
        rb1 = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(col0 + b_col + 64) * DAD_D + kw));
    };
    auto sstore = [&](int buf) {
        As[buf][a_kq * 4 + 0][a_row] = ra.x; As[buf][a_kq * 4 + 1][a_row] = ra.y;
        As[buf][a_kq * 4 + 2][a_row] = ra.z; As[buf][a_kq * 4 + 3][a_row] = ra.w;
        Bs[buf][b_kq * 4 + 0][b_col] = rb0.x; Bs[buf][b_kq * 4 + 1][b_col] = rb0.y;
        Bs[buf][b_kq * 4 + 2][b_col] = rb0.z; Bs[buf][b_kq * 4 + 3][b_col] = rb0.w;
        Bs[buf][b_kq * 4 + 0][b_col + 64] = rb1.x; Bs[buf][b_kq * 4 + 1][b_col + 64] = rb1.y;
        Bs[buf][b_kq * 4 + 2][b_col + 64] = rb1.z; Bs[buf][b_kq * 4 + 3][b_col + 64] = rb1.w;
    };

    const int ty = tid >> 4, tx = tid & 15;  // rows ty*4..+3, cols tx*8..+7
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    constexpr int KB = DAD_D / F_BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kb = 0; kb < KB; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < KB) gload(kb + 1);
#pragma unroll
        for (int k = 0; k < F_BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8 + 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kb + 1 < KB) sstore(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue: bias + ReLU + padding mask, ReLU bit mask, pooled partial sums
    float bcol[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bcol[j] = __ldg(bias + col0 + tx * 8 + j);
    float psum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) psum[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty * 4 + i;
        const bool in = t < T;
        const bool valid = in && !(pad && pad[(int64_t)b * T + t]);
        uint32_t byte = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float pre = acc[i][j] + bcol[j];
            if (pre > 0.f && valid) {  // padded frames carry no ReLU bits: they can never contribute to the backward
                byte |= 1u << j;
                psum[j] += pre;
            }
        }
        if (bits) {  // 4 neighbouring lanes (tx%4) hold the 4 bytes of one 32-column word
            uint32_t w = byte << ((tx & 3) * 8);
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            if (in && (tx & 3) == 0) bits[((int64_t)b * T + t) * DAD_HW + ct * 4 + (tx >> 2)] = w;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = psum[j];
    __syncthreads();
    if (tid < F_BN) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) s += red[r][tid];
        part[((int64_t)b * tiles_t + tile_t) * DAD_H + col0 + tid] = s;
    }
    // ---- last CTA of (b, ct) sums the tile partials in fixed order and divides by the valid length
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&counters[b * 2 + ct], 1u);
        s_last = (prev == (unsigned)tiles_t - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float cnt = 0.f;
    if (pad) {
        for (int t = tid; t < T; t += F_THREADS) cnt += pad[(int64_t)b * T + t] ? 0.f : 1.f;
        cnt = warp_sum(cnt);
        if ((tid & 31) == 0) red[0][tid >> 5] = cnt;
        __syncthreads();
        if (tid == 0) {
            float c = 0.f;
            for (int w = 0; w < F_THREADS / 32; ++w) c += red[0][w];
            s_len = fmaxf(c, 1.0f);
        }
    } else if (tid == 0) {
        s_len = (float)T;
    }
    __syncthreads();
    if (tid < F_BN) {
        float s = 0.f;
        for (int i = 0; i < tiles_t; ++i) s += __ldcg(&part[((int64_t)b * tiles_t + i) * DAD_H + col0 + tid]);
        emb[(int64_t)b * DAD_H + col0 + tid] = s / s_len;
    }
    if (tid == 0) {
        counters[b * 2 + ct] = 0u;  // self-reset: workspace stays zeroed for the next launch / graph replay
        if (ct == 0 && lens_out) lens_out[b] = s_len;
    }
}

constexpr int G_BM = 128, G_BN = 128, G_BK = 16, G_THREADS = 256;

// grid = (D/128, H/128, S).  Split s covers rows [s*rows_per_split, ...) of the flattened [N=B*T] frame axis.
__global__ void __launch_bounds__(G_THREADS) k_enc_bwd_simt(const float* __restrict__ x, const uint8_t* __restrict__ pad, AugParams ap,
                                                           const uint32_t* __restrict__ bits, const float* __restrict__ g_emb,
                                                           const float* __restrict__ lens, float* __restrict__ dWp, float* __restrict__ dbp, int64_t N, int T,
                                                           int64_t rows_per_split) {
    __shared__ __align__(16) float As[2][G_BK][G_BM + 4];  // dpre^T tile  [n][h]
    __shared__ __align__(16) float Bs[2][G_BK][G_BN + 4];  // x tile       [n][d]
    const int tid = threadIdx.x;
    const unsigned long long seed = (ap.mode && ap.noise_src == 1) ? aug_seed(ap) : 0ull;
    const int d0 = blockIdx.x * G_BN, h0 = blockIdx.y * G_BM, split = blockIdx.z;
    const int64_t n_begin = (int64_t)split * rows_per_split;
    const int64_t n_end = (n_begin + rows_per_split < N) ? n_begin + rows_per_split : N;

    const int l_row = tid >> 4, l_seg = tid & 15;  // loader: row within k-block, 8-wide segment
    float4 ra0, ra1, rb0, rb1;
    auto gload = [&](int64_t nb) {
        const int64_t n = nb + l_row;
        const bool in = n < n_end;
        ra0 = ra1 = rb0 = rb1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in) {
            const int b = (int)(n / T), t = (int)(n - (int64_t)b * T);
            const int d = d0 + l_seg * 8;
            rb0 = load_x_aug(x, ap, seed, n * DAD_D + d, b, t, d);
            rb1 = load_x_aug(x, ap, seed, n * DAD_D + d + 4, b, t, d + 4);
            const bool valid = !(pad && pad[n]);
            if (valid) {
                const int h = h0 + l_seg * 8;
                const uint32_t word = __ldg(bits + n * DAD_HW + (h >> 5));
                const uint32_t byte = (word >> (h & 31)) & 0xffu;
                float4 g0 = __ldg(reinterpret_cast<const float4*>(g_emb + (int64_t)b * DAD_H + h));
                float4 g1 = __ldg(reinterpret_cast<const float4*>(g_emb + (int64_t)b * DAD_H + h + 4));
                const float len = __ldg(lens + b);  // autograd of sum/len: g / len (division, as the reference)
                g0.x /= len; g0.y /= len; g0.z /= len; g0.w /= len; g1.x /= len; g1.y /= len; g1.z /= len; g1.w /= len;
                ra0 = make_float4((byte & 1u) ? g0.x : 0.f, (byte & 2u) ? g0.y : 0.f, (byte & 4u) ? g0.z : 0.f, (byte & 8u) ? g0.w : 0.f);
                ra1 = make_float4((byte & 16u) ? g1.x : 0.f, (byte & 32u) ? g1.y : 0.f, (byte & 64u) ? g1.z : 0.f, (byte & 128u) ? g1.w : 0.f);
            }
        }
    };
    auto sstore = [&](int buf) {
        *reinterpret_cast<float4*>(&As[buf][l_row][l_seg * 8]) = ra0;
        *reinterpret_cast<float4*>(&As[buf][l_row][l_seg * 8 + 4]) = ra1;
        *reinterpret_cast<float4*>(&Bs[buf][l_row][l_seg * 8]) = rb0;
        *reinterpret_cast<float4*>(&Bs[buf][l_row][l_seg * 8 + 4]) = rb1;
    };

    const int ty = tid >> 4, tx = tid & 15;  // h rows ty*8..+7, d cols tx*8..+7
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float dbacc = 0.f;

    const int64_t nkb = (n_end > n_begin) ? ceil_div64(n_end - n_begin, G_BK) : 0;
    if (nkb > 0) {
        gload(n_begin);
        sstore(0);
    }
    __syncthreads();
    for (int64_t kb = 0; kb < nkb; ++kb) {
        const int buf = (int)(kb & 1);
        if (kb + 1 < nkb) gload(n_begin + (kb + 1) * G_BK);
#pragma unroll
        for (int k = 0; k < G_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8 + 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (blockIdx.x == 0 && tid < G_BM) {
#pragma unroll
            for (int k = 0; k < G_BK; ++k) dbacc += As[buf][k][tid];
        }
        if (kb + 1 < nkb) sstore(buf ^ 1);
        __syncthreads();
    }
    float* out = dWp + (int64_t)split * DAD_H * DAD_D;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float* row = out + (int64_t)(h0 + ty * 8 + i) * DAD_D + d0 + tx * 8;
        *reinterpret_cast<float4*>(row) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(row + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
    if (blockIdx.x == 0 && tid < G_BM) dbp[(int64_t)split * DAD_H + h0 + tid] = dbacc;
}

}  // namespace

size_t enc_fwd_simt_ws_bytes(int64_t B, int64_t T) {
    const int64_t tiles_t = ceil_div64(T, F_BM);
    return (size_t)(B * tiles_t * DAD_H * sizeof(float));
}

// counters: >= 2*B zero-initialised unsigned (self-resetting).
int launch_enc_fwd_simt(const float* x, const uint8_t* pad, const float* W, const float* bias, const AugParams& ap, float* emb,
                        uint32_t* bits, float* lens_out, float* part_ws, unsigned* counters, int64_t B, int64_t T, cudaStream_t st) {
    if (B == 0 || T == 0) return DAD_OK;
    dim3 grid((unsigned)ceil_div64(T, F_BM), (unsigned)B, DAD_H / F_BN);
    k_enc_fwd_simt<<<grid, F_THREADS, 0, st>>>(x, pad, W, bias, ap, part_ws, bits, emb, lens_out, counters, (int)B, (int)T);
    DAD_CHECK_LAUNCH("k_enc_fwd_simt");
    return DAD_OK;
}

int enc_bwd_simt_splits(int64_t N) {
    // 12 output tiles per split; aim at ~2 waves of 148 SMs, at least 4 k-blocks of work per split
    int64_t s = 24;
    const int64_t max_s = ceil_div64(N, 4 * G_BK);
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    return (int)s;
}

// dWp: [S][H][D], dbp: [S][H]; the caller reduces the S partials (k_optim.cu: grad reduce) in fixed order.
int launch_enc_bwd_simt(const float* x, const uint8_t* pad, const AugParams& ap, const uint32_t* bits, const float* g_emb, const float* lens,
                        float* dWp, float* dbp, int splits, int64_t B, int64_t T, cudaStream_t st) {
    const int64_t N = B * T;
    const int64_t rows_per_split = ceil_div64(ceil_div64(N, splits), G_BK) * G_BK;
    dim3 grid(DAD_D / G_BN, DAD_H / G_BM, (unsigned)splits);
    k_enc_bwd_simt<<<grid, G_THREADS, 0, st>>>(x, pad, ap, bits, g_emb, lens, dWp, dbp, N, (int)T, rows_per_split);
    DAD_CHECK_LAUNCH("k_enc_bwd_simt");
    return DAD_OK;
}
