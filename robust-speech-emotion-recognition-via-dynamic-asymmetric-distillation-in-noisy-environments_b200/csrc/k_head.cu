// Head kernels: classifier (dropout + 256->4 linear), softmax / log-softmax, label-smoothed CE, DACP certainty
// scores + per-class quantile thresholds + mask, DACP-masked KL consistency, and the analytic backward of all of
// it down to d(loss)/d(embedding), dWc, dbc.  All reductions are warp-shuffle / shared-memory, fixed order.
// Reference: I/model.py:54-64, I/train.py:397-447, I/utils.py:400-507.
#include "dad_common.cuh"
#include "dad_pack.cuh"

namespace {

constexpr int HD_THREADS = 128;  // 4 warps, one sample per warp

struct Logits4 { float z[DAD_C]; };

// z[c] = bc[c] + sum_h u[h] Wc[c,h]; lane holds u for h = lane*8..+7
__device__ __forceinline__ Logits4 head_gemv(const float (&u)[8], const float* __restrict__ Wc, const float* __restrict__ bc, int lane) {
    Logits4 r;
#pragma unroll
    for (int c = 0; c < DAD_C; ++c) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(Wc + c * DAD_H + lane * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(Wc + c * DAD_H + lane * 8 + 4));
        float s = u[0] * w0.x;
        s = fmaf(u[1], w0.y, s); s = fmaf(u[2], w0.z, s); s = fmaf(u[3], w0.w, s);
        s = fmaf(u[4], w1.x, s); s = fmaf(u[5], w1.y, s); s = fmaf(u[6], w1.z, s); s = fmaf(u[7], w1.w, s);
        r.z[c] = warp_sum(s) + __ldg(bc + c);
    }
    return r;
}

__device__ __forceinline__ void load8(const float* p, float (&u)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    u[0] = a.x; u[1] = a.y; u[2] = a.z; u[3] = a.w; u[4] = b.x; u[5] = b.y; u[6] = b.z; u[7] = b.w;
}

__device__ __forceinline__ void log_softmax4(const float (&z)[DAD_C], float (&lp)[DAD_C]) {
    const float m = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
    const float s = expf(z[0] - m) + expf(z[1] - m) + expf(z[2] - m) + expf(z[3] - m);
    const float l = logf(s);
#pragma unroll
    for (int c = 0; c < DAD_C; ++c) lp[c] = z[c] - m - l;
}
__device__ __forceinline__ void softmax4(const float (&z)[DAD_C], float (&p)[DAD_C]) {
    const float m = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
    float e[DAD_C], s = 0.f;
#pragma unroll
    for (int c = 0; c < DAD_C; ++c) { e[c] = expf(z[c] - m); s += e[c]; }
#pragma unroll
    for (int c = 0; c < DAD_C; ++c) p[c] = e[c] / s;
}

// DACPManager.calculate_certainty_scores (I/utils.py:400-428): max prob, argmax, optional entropy weighting
__device__ __forceinline__ void certainty4(const float (&p)[DAD_C], int use_entropy, float& score, int& pred) {
    float maxp = p[0];
    pred = 0;
#pragma unroll
    for (int c = 1; c < DAD_C; ++c)
        if (p[c] > maxp) { maxp = p[c]; pred = c; }
    score = maxp;
    if (use_entropy) {
        float ent = 0.f;
#pragma unroll
        for (int c = 0; c < DAD_C; ++c) ent -= p[c] * log2f(p[c] + 1e-8f);
        score = maxp * (1.f - ent / 2.0f);  // log2(4) = 2
    }
}

}  // namespace

// stand-alone certainty scores for the op-level API (DACPManager.calculate_certainty_scores): one thread per sample
__global__ void __launch_bounds__(256) k_certainty(const float* __restrict__ probs, int B, int use_entropy, float* __restrict__ scores,
                                                   int* __restrict__ preds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float4 q = *reinterpret_cast<const float4*>(probs + (int64_t)i * DAD_C);
    const float p[DAD_C] = {q.x, q.y, q.z, q.w};
    float score;
    int pred;
    certainty4(p, use_entropy, score, pred);
    scores[i] = score;
    preds[i] = pred;
}

// ---------------------------------------------------------------------------------------------------------
// k_head_fwd: per local sample.  mode bit0: clean head (CE), bit1: noisy heads (teacher probs/score + student log q)
// Outputs (local arrays, length B_l): dz_clean[B,4] (already /B_global, CE grad), probs[B,4], logq[B,4],
// and into the pack (global layout): score, pred.  ce_part[grid] = per-CTA sum of per-sample CE.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HD_THREADS) k_head_fwd(HeadFwdArgs a) {
    __shared__ float s_ce[HD_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * (HD_THREADS / 32) + warp;  // local sample
    float ce_i = 0.f;
    if (i < a.B_local) {
        const int gi = a.rank * a.B_local + i;
        float u[8];
        // ---- clean student head + label-smoothed CE (I/train.py:399-400)
        {
            load8(pack_emb_clean(a.pack, gi) + lane * 8, u);
            if (a.keep_clean) {
                float k[8];
                load8(a.keep_clean + (int64_t)i * DAD_H + lane * 8, k);
#pragma unroll
                for (int j = 0; j < 8; ++j) u[j] = u[j] * (k[j] * a.drop_scale);
            }
            const Logits4 z = head_gemv(u, a.Wc, a.bc, lane);
            float lp[DAD_C];
            log_softmax4(z.z, lp);
            // samples i >= n_clean are padding ("null" samples of a ragged last batch): label := -1, no CE, zero gradient
            const bool real = i < a.n_clean;
            const int y = real ? pack_label(a.pack, gi) : -1;
            if (!real && lane == 0) pack_label(a.pack, gi) = -1;
            const float eps = a.label_smoothing;
            float nll = 0.f, smooth = 0.f;
#pragma unroll
            for (int c = 0; c < DAD_C; ++c) {
                smooth -= lp[c];
                if (c == y) nll = -lp[c];
            }
            ce_i = real ? (1.f - eps) * nll + eps * (smooth / DAD_C) : 0.f;
            if (lane < DAD_C) {
                const float p = expf(lp[lane]);
                a.dz_clean[i * DAD_C + lane] = real ? (p - (1.f - eps) * (lane == y ? 1.f : 0.f) - eps / DAD_C) * a.inv_B_global : 0.f;
            }
        }
        if (a.noisy) {
            // ---- teacher head on the weak view: probs, certainty score, pseudo label (I/train.py:408-410, I/utils.py:400-428)
            load8(a.emb_teacher + (int64_t)i * DAD_H + lane * 8, u);
            const Logits4 zt = head_gemv(u, a.tWc, a.tbc, lane);
            float p[DAD_C];
            softmax4(zt.z, p);
            float score;
            int pred;
            certainty4(p, a.use_entropy, score, pred);
            if (i >= a.n_noisy) { score = 0.f; pred = -1; }  // null sample: no pseudo-label, never masked in, not in the quantiles
            // ---- strong student head (I/train.py:439-440)
            load8(pack_emb_strong(a.pack, gi) + lane * 8, u);
            if (a.keep_strong) {
                float k[8];
                load8(a.keep_strong + (int64_t)i * DAD_H + lane * 8, k);
#pragma unroll
                for (int j = 0; j < 8; ++j) u[j] = u[j] * (k[j] * a.drop_scale);
            }
            const Logits4 zs = head_gemv(u, a.Wc, a.bc, lane);
            float lq[DAD_C];
            log_softmax4(zs.z, lq);
            if (lane < DAD_C) {
                a.probs[i * DAD_C + lane] = p[lane];
                a.logq[i * DAD_C + lane] = lq[lane];
            }
            if (lane == 0) {
                pack_score(a.pack, gi) = score;
                pack_pred(a.pack, gi) = pred;
            }
        }
    }
    if (lane == 0) s_ce[warp] = ce_i;
    __syncthreads();
    if (threadIdx.x == 0) a.ce_part[blockIdx.x] = s_ce[0] + s_ce[1] + s_ce[2] + s_ce[3];
}

// ---------------------------------------------------------------------------------------------------------
// k_dacp: one CTA over the GLOBAL batch (after the all-gather under data parallelism).
// MSB-first radix select over sortable score keys in shared memory -> per-class linear-interpolated quantiles
// (torch.quantile semantics), EMA thresholds (state, in place), mask, per-class epoch accumulators.
// Reference: I/utils.py:449-507; fixed-threshold branch I/train.py:417-420.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2sortable(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float sortable2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(1024) k_dacp(DacpArgs a) {
    extern __shared__ uint32_t keys[];  // [Bg] sortable score keys (class in s_cls)
    __shared__ int s_cnt[DAD_C];
    __shared__ float s_thr[DAD_C];
    __shared__ int s_nmask;
    __shared__ unsigned long long s_sum[DAD_C];  // per-class score sums in 2^-40 fixed point (integer adds: order-free)
    __shared__ int s_hist[2 * DAD_C][256];   // radix-select histograms: target = class * 2 + {lo, hi}
    __shared__ uint32_t s_prefix[2 * DAD_C]; // key bits fixed so far
    __shared__ int s_rank[2 * DAD_C];        // remaining 0-based rank inside the prefix bucket (-1: class empty)
    const int tid = threadIdx.x, nt = blockDim.x;
    const int Bg = a.pack.B_local * a.pack.world;
    if (tid < DAD_C) { s_cnt[tid] = 0; s_sum[tid] = 0ull; }
    if (tid == 0) s_nmask = 0;
    __syncthreads();

    if (a.use_dacp) {
        // The per-class quantile needs two order statistics per class (torch.quantile, linear interpolation): an MSB-first
        // radix SELECT over the 32-bit sortable keys -- 4 passes of 8 bits, 8 targets at once -- instead of a full sort
        // (the bitonic sort of the global batch was ~95 us of a single CTA per step, a fixed cost on every data-parallel rank).
        uint8_t* s_cls = reinterpret_cast<uint8_t*>(keys + Bg);
        for (int i0 = 0; i0 < Bg; i0 += nt) {  // warp-uniform trip count: class counts are merged per warp (ballot) before
            const int i = i0 + tid;             // touching the 4 shared counters (4096 same-address atomics took ~20 us)
            int c = -1;
            if (i < Bg) {
                c = pack_pred(a.pack, i);
                keys[i] = f2sortable(pack_score(a.pack, i));
                s_cls[i] = (uint8_t)(c < 0 ? 255 : c);  // 255: null sample
            }
#pragma unroll
            for (int cc = 0; cc < DAD_C; ++cc) {
                const unsigned b = __ballot_sync(0xffffffffu, c == cc);
                if ((tid & 31) == 0 && b) atomicAdd(&s_cnt[cc], __popc(b));
            }
        }
        __syncthreads();
        if (tid < 2 * DAD_C) {
            const int c = tid >> 1, n = s_cnt[c];
            int rank = -1;
            if (n > 0) {  // pos = q*(n-1); lo = floor, hi = ceil
                const float pos = a.gamma * (float)(n - 1);
                rank = (tid & 1) ? (int)ceilf(pos) : (int)floorf(pos);
            }
            s_rank[tid] = rank;
            s_prefix[tid] = 0u;
        }
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            for (int e = tid; e < 2 * DAD_C * 256; e += nt) (&s_hist[0][0])[e] = 0;
            __syncthreads();
            // scores cluster in a few exponent buckets, so the first passes would hammer two or three bins with 32-way
            // conflicting shared-memory atomics: lanes with the same (target, digit) are merged first (match.any)
            for (int i0 = 0; i0 < Bg; i0 += nt) {  // warp-uniform trip count
                const int i = i0 + tid;
                const int c = i < Bg ? s_cls[i] : 255;
                const uint32_t k = i < Bg ? keys[i] : 0u;
                const uint32_t hi_bits = pass == 0 ? 0u : (k >> (shift + 8));
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const int t = (c == 255 ? 0 : c) * 2 + w;
                    const bool on = c != 255 && s_rank[t] >= 0 && (pass == 0 || hi_bits == (s_prefix[t] >> (shift + 8)));
                    const uint32_t id = on ? ((uint32_t)t << 8) | ((k >> shift) & 0xFFu) : 0xFFFFFFFFu;
                    const unsigned peers = __match_any_sync(0xffffffffu, id);
                    if (on && (tid & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[t][(k >> shift) & 0xFFu], __popc(peers));
                }
            }
            __syncthreads();
            // one warp per target: find the bin holding the rank (lane = 8 consecutive bins, exclusive warp scan)
            const int warp = tid >> 5, lane = tid & 31;
            if (warp < 2 * DAD_C && s_rank[warp] >= 0) {
                int h[8], tot = 0;
#pragma unroll
                for (int u = 0; u < 8; ++u) { h[u] = s_hist[warp][lane * 8 + u]; tot += h[u]; }
                int incl = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const int excl = incl - tot, rank = s_rank[warp];
                if (rank >= excl && rank < incl) {  // exactly one lane
                    int r = rank - excl, bin = 0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (r >= 0 && r < h[u]) { bin = u; s_rank[warp] = r; r = -1; }
                        else if (r >= 0) r -= h[u];
                    }
                    s_prefix[warp] |= (uint32_t)(lane * 8 + bin) << shift;
                }
            }
            __syncthreads();
        }
        if (tid < DAD_C) {
            const int c = tid;
            const int n = s_cnt[c];
            // class weights W_c = sigmoid(K (Q_c - mean Q))  (utils.py:466-470)
            const float q0 = a.quality[0], q1 = a.quality[1], q2 = a.quality[2], q3 = a.quality[3];
            const float meanq = (q0 + q1 + q2 + q3) / 4.0f;
            const float w = 1.0f / (1.0f + expf(-(a.sens_k * (a.quality[c] - meanq))));
            a.out_w[c] = w;
            float tau;
            if (n > 0) {  // torch.quantile(linear): pos = q*(n-1); lerp(sorted[floor], sorted[ceil], frac)
                const float pos = a.gamma * (float)(n - 1);
                const float wgt = pos - floorf(pos);
                const float vlo = sortable2f(s_prefix[2 * c]);
                const float vhi = sortable2f(s_prefix[2 * c + 1]);
                const float diff = vhi - vlo;
                tau = (wgt < 0.5f) ? vlo + wgt * diff : vhi - diff * (1.0f - wgt);
            } else {
                tau = a.ema_thr[c];
            }
            tau = tau + a.calib_lambda * (w - 0.5f);
            tau = fmaxf(tau, a.anchors[c]);
            const float thr = a.alpha * a.ema_thr[c] + a.one_minus_alpha * tau;
            a.ema_thr[c] = thr;  // state advances every post-warm-up step (utils.py:496-499)
            s_thr[c] = thr;
        }
    } else if (tid < DAD_C) {
        s_thr[tid] = a.fixed_thr;
        a.out_w[tid] = 1.0f;  // class_weights_wce = ones (I/train.py:420) -> class attention exp(0) = 1
    }
    __syncthreads();
    int local = 0;
    // per-class epoch accumulators (utils.py:503-505) in 2^-40 fixed point: scores in [2^-16, 1] convert exactly, sums of up
    // to 2^13 of them stay below 2^53, and integer adds commute -- bit-reproducible whatever the order, without fp64 (which issues
    // at 1/16 rate: the four fp64 warp reductions per 1024 samples were ~5 us of this single CTA)
    long long acc[DAD_C];
#pragma unroll
    for (int cc = 0; cc < DAD_C; ++cc) acc[cc] = 0;
    for (int i0 = 0; i0 < Bg; i0 += nt) {
        const int i = i0 + tid;
        if (i < Bg) {
            const int c = pack_pred(a.pack, i);
            const float s = pack_score(a.pack, i);
            const int m = c >= 0 && s >= s_thr[c < 0 ? 0 : c];
            a.mask[i] = (uint8_t)m;
            local += m;
            const long long q = __float2ll_rn(s * 1099511627776.0f);
#pragma unroll
            for (int cc = 0; cc < DAD_C; ++cc) acc[cc] += c == cc ? q : 0ll;
        }
    }
    if (a.use_dacp) {
#pragma unroll
        for (int cc = 0; cc < DAD_C; ++cc) {
            long long v = acc[cc];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0 && v != 0) atomicAdd(&s_sum[cc], (unsigned long long)v);
        }
    }
    local = (int)warp_sum((float)local);
    if ((tid & 31) == 0) atomicAdd(&s_nmask, local);
    __syncthreads();
    if (tid < DAD_C) {
        a.out_thr[tid] = s_thr[tid];
        if (a.use_dacp) {
            a.epoch_sum[tid] += (double)(long long)s_sum[tid] * 9.094947017729282e-13;  // 2^-40
            a.epoch_cnt[tid] += (int64_t)s_cnt[tid];
        }
    }
    if (tid == 0) a.out_nmask[0] = s_nmask;
}

// ---------------------------------------------------------------------------------------------------------
// k_head_bwd: per local sample: KD loss term + d/dz of CE and KD, back through the classifier and dropout to
// d/d(embedding) (written, not accumulated: ECDA adds on top later), and per-CTA partials of dWc / dbc.
// The last CTA reduces the partials in fixed order and writes the CE / KD loss scalars.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HD_THREADS) k_head_bwd(HeadBwdArgs a) {
    __shared__ float s_dWc[HD_THREADS / 32][DAD_C][DAD_H];  // 16 KB
    __shared__ float s_dbc[HD_THREADS / 32][DAD_C];
    __shared__ float s_kd[HD_THREADS / 32];
    __shared__ float s_red[2][HD_THREADS];
    __shared__ int s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nmask = a.noisy ? a.nmask[0] : 0;
    const bool kd_on = a.noisy && nmask > 1;  // I/train.py:444
    const float kd_den = (float)nmask + 1e-8f;
    float dWc[DAD_C][8];
#pragma unroll
    for (int c = 0; c < DAD_C; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) dWc[c][j] = 0.f;
    float dbc = 0.f, kd_i = 0.f;
    // persistent: warp w of CTA g owns samples g*4 + w, + 4*gridDim.x, ... (fixed order -> deterministic partials)
    for (int i = blockIdx.x * (HD_THREADS / 32) + warp; i < a.B_local; i += gridDim.x * (HD_THREADS / 32)) {
        const int gi = a.rank * a.B_local + i;
        // ---- clean path
        {
            float u[8], ks[8];
            load8(pack_emb_clean(a.pack, gi) + lane * 8, u);
#pragma unroll
            for (int j = 0; j < 8; ++j) ks[j] = 1.f;
            if (a.keep_clean) {
                load8(a.keep_clean + (int64_t)i * DAD_H + lane * 8, ks);
#pragma unroll
                for (int j = 0; j < 8; ++j) { ks[j] *= a.drop_scale; u[j] *= ks[j]; }
            }
            float dz[DAD_C];
#pragma unroll
            for (int c = 0; c < DAD_C; ++c) dz[c] = a.dz_clean[i * DAD_C + c];
            float g[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = 0.f;
#pragma unroll
            for (int c = 0; c < DAD_C; ++c) {
                float w[8];
                load8(a.Wc + c * DAD_H + lane * 8, w);
#pragma unroll
                for (int j = 0; j < 8; ++j) { g[j] = fmaf(dz[c], w[j], g[j]); dWc[c][j] = fmaf(dz[c], u[j], dWc[c][j]); }
            }
            if (lane < DAD_C) dbc += dz[lane];
            float* go = a.g_clean + (int64_t)i * DAD_H + lane * 8;
            *reinterpret_cast<float4*>(go) = make_float4(g[0] * ks[0], g[1] * ks[1], g[2] * ks[2], g[3] * ks[3]);
            *reinterpret_cast<float4*>(go + 4) = make_float4(g[4] * ks[4], g[5] * ks[5], g[6] * ks[6], g[7] * ks[7]);
        }
        if (a.noisy) {
            float g[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = 0.f;
            float ks[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) ks[j] = 1.f;
            const int m = a.mask[gi];
            if (kd_on && m) {
                float u[8];
                load8(pack_emb_strong(a.pack, gi) + lane * 8, u);
                if (a.keep_strong) {
                    load8(a.keep_strong + (int64_t)i * DAD_H + lane * 8, ks);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { ks[j] *= a.drop_scale; u[j] *= ks[j]; }
                }
                float dz[DAD_C];
#pragma unroll
                for (int c = 0; c < DAD_C; ++c) {
                    const float p = a.probs[i * DAD_C + c], lq = a.logq[i * DAD_C + c];
                    kd_i += (p > 0.f) ? p * (logf(p) - lq) : 0.f;  // kl_div(log q, p) = xlogy(p,p) - p log q
                    dz[c] = a.w_kd * (expf(lq) - p) / kd_den;       // d/dz_s of sum_c p(log p - log softmax(z)) = q - p
                }
#pragma unroll
                for (int c = 0; c < DAD_C; ++c) {
                    float w[8];
                    load8(a.Wc + c * DAD_H + lane * 8, w);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { g[j] = fmaf(dz[c], w[j], g[j]); dWc[c][j] = fmaf(dz[c], u[j], dWc[c][j]); }
                }
                if (lane < DAD_C) dbc += dz[lane];
            }
            float* go = a.g_strong + (int64_t)i * DAD_H + lane * 8;
            *reinterpret_cast<float4*>(go) = make_float4(g[0] * ks[0], g[1] * ks[1], g[2] * ks[2], g[3] * ks[3]);
            *reinterpret_cast<float4*>(go + 4) = make_float4(g[4] * ks[4], g[5] * ks[5], g[6] * ks[6], g[7] * ks[7]);
        }
    }
    // ---- CTA partials
#pragma unroll
    for (int c = 0; c < DAD_C; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) s_dWc[warp][c][lane * 8 + j] = dWc[c][j];
    if (lane < DAD_C) s_dbc[warp][lane] = dbc;
    if (lane == 0) s_kd[warp] = kd_i;
    __syncthreads();
    const int G = gridDim.x;
    float* pw = a.part + (int64_t)blockIdx.x * (DAD_C * DAD_H + DAD_C + 1);
    for (int e = threadIdx.x; e < DAD_C * DAD_H; e += HD_THREADS) {
        const int c = e / DAD_H, h = e - c * DAD_H;
        pw[e] = s_dWc[0][c][h] + s_dWc[1][c][h] + s_dWc[2][c][h] + s_dWc[3][c][h];
    }
    if (threadIdx.x < DAD_C) pw[DAD_C * DAD_H + threadIdx.x] = s_dbc[0][threadIdx.x] + s_dbc[1][threadIdx.x] + s_dbc[2][threadIdx.x] + s_dbc[3][threadIdx.x];
    if (threadIdx.x == 0) pw[DAD_C * DAD_H + DAD_C] = s_kd[0] + s_kd[1] + s_kd[2] + s_kd[3];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(a.counter, 1u);
        s_last = (prev == (unsigned)G - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- last CTA: fixed-order reduction of the G (<= SM count) CTA partials, 4 independent chains per element
    const int stride = DAD_C * DAD_H + DAD_C + 1;
    for (int e = threadIdx.x; e < DAD_C * DAD_H + DAD_C; e += HD_THREADS) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int g = 0;
        for (; g + 3 < G; g += 4) {
            s0 += __ldcg(a.part + (int64_t)g * stride + e);
            s1 += __ldcg(a.part + (int64_t)(g + 1) * stride + e);
            s2 += __ldcg(a.part + (int64_t)(g + 2) * stride + e);
            s3 += __ldcg(a.part + (int64_t)(g + 3) * stride + e);
        }
        for (; g < G; ++g) s0 += __ldcg(a.part + (int64_t)g * stride + e);
        const float s = (s0 + s1) + (s2 + s3);
        if (e < DAD_C * DAD_H) a.dWc[e] = s;
        else a.dbc[e - DAD_C * DAD_H] = s;
    }
    // loss scalars: thread-strided partial sums, then a fixed-order sum over the 128 threads
    float kd = 0.f, ce = 0.f;
    for (int g = threadIdx.x; g < G; g += HD_THREADS) kd += __ldcg(a.part + (int64_t)g * stride + DAD_C * DAD_H + DAD_C);
    for (int g = threadIdx.x; g < a.n_ce_part; g += HD_THREADS) ce += __ldcg(a.ce_part + g);
    s_red[0][threadIdx.x] = kd;
    s_red[1][threadIdx.x] = ce;
    __syncthreads();
    if (threadIdx.x == 0) {
        kd = 0.f; ce = 0.f;
        for (int t = 0; t < HD_THREADS; ++t) { kd += s_red[0][t]; ce += s_red[1][t]; }
        a.loss_out[0] = ce * a.inv_B_global;         // supervised CE (local share under DP; summed by the all-reduce)
        a.loss_out[1] = kd_on ? kd / kd_den : 0.f;   // consistency (local share under DP)
        *a.counter = 0u;
    }
}

int launch_head_fwd(const HeadFwdArgs& a, cudaStream_t st) {
    if (a.B_local == 0) return DAD_OK;
    k_head_fwd<<<ceil_div(a.B_local, HD_THREADS / 32), HD_THREADS, 0, st>>>(a);
    DAD_CHECK_LAUNCH("k_head_fwd");
    return DAD_OK;
}
int head_grid(int B_local) { return ceil_div(B_local, HD_THREADS / 32); }

int launch_certainty(const float* probs, int B, int use_entropy, float* scores, int* preds, cudaStream_t st) {
    if (B == 0) return DAD_OK;
    k_certainty<<<ceil_div(B, 256), 256, 0, st>>>(probs, B, use_entropy, scores, preds);
    DAD_CHECK_LAUNCH("k_certainty");
    return DAD_OK;
}

int launch_dacp(DacpArgs a, cudaStream_t st) {
    const int Bg = a.pack.B_local * a.pack.world;
    a.npow2_ = 0;
    const size_t smem = (size_t)Bg * (sizeof(uint32_t) + 1) + 16;  // sortable keys + class bytes
    if (smem > 180 * 1024) return dad_set_error(DAD_ERR_UNSUPPORTED, "dacp: global batch %d too large for the single-CTA select", Bg);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_dacp, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
        attr_set = true;
    }
    int threads = 256;
    while (threads < 1024 && threads * 4 < Bg) threads <<= 1;
    k_dacp<<<1, threads, smem, st>>>(a);
    DAD_CHECK_LAUNCH("k_dacp");
    return DAD_OK;
}

int launch_head_bwd(const HeadBwdArgs& a, cudaStream_t st) {
    if (a.B_local == 0) return DAD_OK;
    int grid = ceil_div(a.B_local, HD_THREADS / 32);  // <= dad_head_grid(B_local): the `part` workspace is sized for that
    if (grid > 148) grid = 148;
    k_head_bwd<<<grid, HD_THREADS, 0, st>>>(a);
    DAD_CHECK_LAUNCH("k_head_bwd");
    return DAD_OK;
}
