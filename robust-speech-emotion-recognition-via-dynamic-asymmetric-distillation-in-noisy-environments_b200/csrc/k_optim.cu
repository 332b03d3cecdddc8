// Optimiser tail as multi-tensor, bandwidth-bound kernels over ONE flat parameter buffer
//   flat layout (fp32): [ W 256x768 | b 256 | Wc 4x256 | bc 4 ]  = 197 892 elements (student, teacher, m, v, grad)
//   k_grad_reduce : fixed-order sum of the split-K dW/db partials of the two encoder backward launches + head grads
//   k_grad_norm   : global L2 norm (fp64 partials, last CTA) -> clip coefficient, Adam step scalars (step counter on device)
//   k_adam_ema    : clip * grad -> coupled-L2 Adam -> EMA teacher update, plus the fp16 hi/lo split of W and teacher W
//                   that feeds the next step's tcgen05 encoder kernels (fp16, scaled by 2^8).
// Reference: I/train.py:487-492 (clip_grad_norm_, Adam.step), I/train.py:362 (Adam, wd 1e-5), I/model.py:211-223 (EMA).
#include "dad_common.cuh"
#include "dad_optim.cuh"
#include "tc_ptx.cuh"

constexpr int NP_W = DAD_H * DAD_D;          // 196608
constexpr int NP_B = DAD_H;                  // 256
constexpr int NP_WC = DAD_C * DAD_H;         // 1024
constexpr int NP_BC = DAD_C;                 // 4
constexpr int NP = NP_W + NP_B + NP_WC + NP_BC;  // 197892
constexpr int OPT_THREADS = 256;


constexpr int NPX = NP + 8;  // gradient + the 8 loss-share words that travel with it

__global__ void __launch_bounds__(OPT_THREADS) k_grad_reduce(GradReduceArgs a) {
    const int e = blockIdx.x * OPT_THREADS + threadIdx.x;
    const bool xchg = a.x.world > 1;
    float s = 0.f;
    if (e >= NP) {
        if (xchg && e < NPX) s = a.g[e];  // loss shares: already in place, they only travel
    } else if (e < NP_W) {
#pragma unroll
        for (int l = 0; l < 2; ++l)
            if (a.dWp[l])
                for (int k = 0; k < a.splits[l]; ++k) s += __ldg(a.dWp[l] + (int64_t)k * NP_W + e);
    } else if (e < NP_W + NP_B) {
        const int h = e - NP_W;
#pragma unroll
        for (int l = 0; l < 2; ++l)
            if (a.dbp[l])
                for (int k = 0; k < a.splits[l]; ++k) s += __ldg(a.dbp[l] + (int64_t)k * NP_B + h);
    } else if (e < NP_W + NP_B + NP_WC) {
        s = a.dWc[e - NP_W - NP_B];
    } else {
        s = a.dbc[e - NP_W - NP_B - NP_WC];
    }
    if (!xchg) {
        if (e < NP) a.g[e] = s;
        return;
    }
    // ---- fused all-to-all half of the gradient all-reduce: every rank's stage slot [rank] receives this rank's folded gradient.
    // Four consecutive elements per 16-byte store: lanes 4k..4k+3 hand their values to lane 4k (NPX and the slot offsets are
    // multiples of 4 floats, so a quad never straddles the end).
    const float s1 = __shfl_down_sync(0xffffffffu, s, 1), s2 = __shfl_down_sync(0xffffffffu, s, 2), s3 = __shfl_down_sync(0xffffffffu, s, 3);
    if ((threadIdx.x & 3) == 0 && e < NPX) {
        const float4 v = make_float4(s, s1, s2, s3);
        const size_t off = a.x.stage_off + ((size_t)a.x.rank * NPX + (size_t)e) * sizeof(float);
        for (int r = 0; r < a.x.world; ++r) *reinterpret_cast<float4*>(reinterpret_cast<char*>(a.x.peers.p[r]) + off) = v;
    }
    __syncthreads();
    __shared__ unsigned s_last, s_seq;
    if (threadIdx.x == 0) {
        __threadfence_system();  // one system-scope fence per CTA (cumulative over the CTA's stores ordered by the barrier above)
        const unsigned prev = atomicAdd(a.x.counter, 1u);
        s_last = prev == gridDim.x - 1u;
        if (s_last) { *a.x.counter = 0u; s_seq = ++(*a.x.local_seq); }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    peer_flag_barrier(a.x.peers, a.x.world, a.x.rank, a.x.flag_off, s_seq);  // last CTA: every rank's gradient has landed everywhere
    __syncthreads();
    __threadfence_system();
}


// x^t for an integer t >= 0 by binary powering in fp64 (<= 2 log2 t dependent multiplies, relative error < 1e-14): the generic
// pow(double, double) is a few hundred fp64 instructions, and fp64 issues at 1/16 rate -- two of them on one thread were most of
// this kernel's tail.
__device__ __forceinline__ double pow_int(double x, long long t) {
    double r = 1.0;
    while (t > 0) {
        if (t & 1) r *= x;
        x *= x;
        t >>= 1;
    }
    return r;
}

__global__ void __launch_bounds__(OPT_THREADS) k_grad_norm(NormArgs a) {
    __shared__ double s_red[OPT_THREADS / 32];
    __shared__ double s_bc[2];
    __shared__ int s_last;
    double acc = 0.0;
    if (a.stage) {
        // fused second half of the gradient all-reduce: g = sum over ranks of the staged gradients, in RANK ORDER (the same order on
        // every rank: the replicated optimiser state stays bit-identical), written back for k_adam_ema and squared for the norm
        for (int e = blockIdx.x * OPT_THREADS + threadIdx.x; e < NPX; e += gridDim.x * OPT_THREADS) {
            float s = __ldcg(a.stage + e);  // written by peers: bypass L1
            for (int r = 1; r < a.world; ++r) s += __ldcg(a.stage + (size_t)r * NPX + e);
            a.g[e] = s;
            if (e < NP) acc += (double)s * (double)s;
        }
    } else {
        for (int e = blockIdx.x * OPT_THREADS + threadIdx.x; e < NP; e += gridDim.x * OPT_THREADS) {
            const double v = (double)a.g[e];
            acc += v * v;
        }
    }
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < OPT_THREADS / 32; ++w) t += s_red[w];
        a.part[blockIdx.x] = t;
        __threadfence();
        const unsigned prev = atomicAdd(a.counter, 1u);
        s_last = (prev == gridDim.x - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    // last CTA: warp 0 folds the per-CTA partials (lane-strided, fixed order), warps 1 and 2 the two Adam bias corrections
    __threadfence();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t = *a.step + 1;
    double tot = 0.0;
    if (warp == 0) {
        for (unsigned gI = lane; gI < gridDim.x; gI += 32) tot += __ldcg(a.part + gI);
        tot = warp_sum_d(tot);
    } else if (warp <= 2 && lane == 0) {
        s_bc[warp - 1] = 1.0 - pow_int(warp == 1 ? 0.9 : 0.999, t);
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const float norm = (float)sqrt(tot);
    float coef = 1.0f;
    if (a.max_norm > 0.f) coef = fminf(a.max_norm / (norm + 1e-6f), 1.0f);  // clip_grad_norm_: clamp(max/(norm+1e-6), max=1)
    *a.step = t;
    const double bc1 = s_bc[0], bc2 = s_bc[1];
    a.scal[0] = norm;
    a.scal[1] = coef;
    a.scal[2] = (float)(*a.lr / bc1);
    a.scal[3] = (float)sqrt(bc2);
    *a.counter = 0u;
}


// fp16 hi/lo operand split of an encoder weight, pre-scaled by TC_WS (dad_common.cuh)
__device__ __forceinline__ void split_f16(float x, unsigned short& hi, unsigned short& lo) { tc::split1_f16(x * TC_WS, hi, lo); }

__global__ void __launch_bounds__(OPT_THREADS) k_adam_ema(AdamArgs a) {
    const int e = blockIdx.x * OPT_THREADS + threadIdx.x;
    if (e >= NP) return;
    if (e == 0 && a.loss_src) {  // device-side loss bookkeeping: no per-step host sync (replaces the .item() calls of I/train.py:494-495)
        const float ce = a.loss_src[0], kd = a.loss_src[1], ec = a.loss_src[2], sc = a.loss_src[3];
        float tot = ce + a.w_kd * kd + a.w_ecda * ec;  // I/train.py:462-466
        if (a.w_scl != 0.f) tot += a.w_scl * sc;       // optional SCL term (absent upstream; weight 0 by default)
        a.loss_last[0] = tot; a.loss_last[1] = ce; a.loss_last[2] = kd; a.loss_last[3] = ec; a.loss_last[4] = sc;
        a.loss_acc[0] += tot; a.loss_acc[1] += ce; a.loss_acc[2] += kd; a.loss_acc[3] += ec; a.loss_acc[4] += 1.0; a.loss_acc[5] += sc;
    }
    const float coef = a.scal[1], step_size = a.scal[2], bc2s = a.scal[3];
    float p = a.p[e];
    float g = a.g[e] * coef;                 // clip: g.mul_(coef)
    g = fmaf(a.wd, p, g);                    // grad.add(param, alpha=wd)
    float m = a.m[e], v = a.v[e];
    m = m + 0.1f * (g - m);                  // exp_avg.lerp_(grad, 1-beta1)
    v = v * 0.999f + 0.001f * g * g;         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
    const float denom = sqrtf(v) / bc2s + 1e-8f;
    p = p - step_size * (m / denom);         // param.addcdiv_(exp_avg, denom, value=-step_size)
    a.m[e] = m; a.v[e] = v; a.p[e] = p;
    float t = a.t[e];
    if (a.do_ema) {
        t = __fadd_rn(__fmul_rn(t, a.ema_m), __fmul_rn(p, a.ema_1m));  // t*m + s*(1-m): two muls, one add (model.py:218-223)
        a.t[e] = t;
    }
    if (e < NP_W && a.s_hi) {
        split_f16(p, reinterpret_cast<unsigned short*>(a.s_hi)[e], reinterpret_cast<unsigned short*>(a.s_lo)[e]);
        split_f16(t, reinterpret_cast<unsigned short*>(a.t_hi)[e], reinterpret_cast<unsigned short*>(a.t_lo)[e]);
    }
}

// standalone EMA (SSRLModel.update_teacher_ema) and weight split (after loading weights)
__global__ void __launch_bounds__(OPT_THREADS) k_ema(float* t, const float* s, float ema_m, float ema_1m, int n) {
    const int e = blockIdx.x * OPT_THREADS + threadIdx.x;
    if (e < n) t[e] = __fadd_rn(__fmul_rn(t[e], ema_m), __fmul_rn(s[e], ema_1m));
}
__global__ void __launch_bounds__(OPT_THREADS) k_split_f16(const float* w, unsigned short* hi, unsigned short* lo, int n) {
    const int e = blockIdx.x * OPT_THREADS + threadIdx.x;
    if (e < n) split_f16(w[e], hi[e], lo[e]);
}

int launch_grad_reduce(const GradReduceArgs& a, cudaStream_t st) {
    k_grad_reduce<<<ceil_div(a.x.world > 1 ? NPX : NP, OPT_THREADS), OPT_THREADS, 0, st>>>(a);
    DAD_CHECK_LAUNCH("k_grad_reduce");
    return DAD_OK;
}
constexpr int NORM_GRID = 96;
int launch_grad_norm(const NormArgs& a, cudaStream_t st) {
    k_grad_norm<<<NORM_GRID, OPT_THREADS, 0, st>>>(a);
    DAD_CHECK_LAUNCH("k_grad_norm");
    return DAD_OK;
}
int launch_adam_ema(const AdamArgs& a, cudaStream_t st) {
    k_adam_ema<<<ceil_div(NP, OPT_THREADS), OPT_THREADS, 0, st>>>(a);
    DAD_CHECK_LAUNCH("k_adam_ema");
    return DAD_OK;
}
int launch_ema(float* t, const float* s, float ema_m, float ema_1m, int n, cudaStream_t st) {
    if (n == 0) return DAD_OK;
    k_ema<<<ceil_div(n, OPT_THREADS), OPT_THREADS, 0, st>>>(t, s, ema_m, ema_1m, n);
    DAD_CHECK_LAUNCH("k_ema");
    return DAD_OK;
}
int launch_split_f16(const float* w, void* hi, void* lo, int n, cudaStream_t st) {
    if (n == 0) return DAD_OK;
    k_split_f16<<<ceil_div(n, OPT_THREADS), OPT_THREADS, 0, st>>>(w, (unsigned short*)hi, (unsigned short*)lo, n);
    DAD_CHECK_LAUNCH("k_split_f16");
    return DAD_OK;
}
