// Supervised contrastive loss (SupCon, Khosla et al. 2020), forward + analytic backward to d(loss)/d(embedding).
//
// PARITY UNPINNED: the reference deleted its SCL (C/utils.py:377-378; the old constructor call C/train_CASIA.py:396 shows
// SupervisedContrastiveLoss(temperature=0.1); its weight is forced to 0, C/train_CASIA.py:424-425, I/config.py:64,104).  This op
// follows SURVEY.md Appendix A and is checked against oracle.dad_oracle.scl_loss (an fp64 restatement of the published loss):
//   members = clean samples (label >= 0) U masked-in noisy samples (pseudo-label); f = z / max(|z|, 1e-12); s_ij = f_i.f_j / tau
//   L_i = -(1/|P(i)|) sum_{p in P(i)} [ s_ip - log sum_{a != i} exp(s_ia) ],  P(i) = members of i's class other than i
//   L   = mean of L_i over anchors with |P(i)| >= 1
// Off by default (cfg.TARGET_SCL_WEIGHT = 0, cfg.SCL_START_EPOCH = 5001), so it is written for clarity, not speed: one warp per
// row, dot products by warp shuffle.  Under data parallelism every rank computes the row statistics of all global rows
// (redundantly, identically) and the gradient rows it owns; loss shares sum to the global loss.
#include "dad_common.cuh"
#include "dad_pack.cuh"


namespace {

constexpr int SCL_WARPS = 8;

__device__ __forceinline__ const float* scl_row_ptr(const SclArgs& a, int r, int Bg) {
    return r < Bg ? pack_emb_clean(a.pack, r) : pack_emb_strong(a.pack, r - Bg);
}

__global__ void __launch_bounds__(SCL_WARPS * 32) k_scl_prep(SclArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Bg = a.pack.B_local * a.pack.world, R = 2 * Bg;
    const int r = blockIdx.x * SCL_WARPS + warp;
    if (r >= R) return;
    int c = -1;
    if (r < Bg) c = pack_label(a.pack, r);
    else if (a.noisy && a.mask[r - Bg]) c = pack_pred(a.pack, r - Bg);
    if (c >= DAD_C) c = -1;
    const float* z = scl_row_ptr(a, r, Bg);
    float v[8], ss = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = z[lane + 32 * k]; ss = fmaf(v[k], v[k], ss); }
    ss = warp_sum(ss);
    const float n = fmaxf(sqrtf(ss), 1e-12f);  // F.normalize(eps=1e-12)
#pragma unroll
    for (int k = 0; k < 8; ++k) a.F[(int64_t)r * DAD_H + lane + 32 * k] = c >= 0 ? v[k] / n : 0.f;
    if (lane == 0) {
        a.nrm[r] = n;
        a.cls[r] = c;
        if (c >= 0) atomicAdd(&a.cnt[c], 1);  // integer: order-independent
    }
}

// per global row i: lse_i over all members a != i, 1/|P(i)|, L_i
__global__ void __launch_bounds__(SCL_WARPS * 32) k_scl_rows(SclArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Bg = a.pack.B_local * a.pack.world, R = 2 * Bg;
    const int i = blockIdx.x * SCL_WARPS + warp;
    if (i >= R) return;
    const int ci = a.cls[i];
    const bool anchor = ci >= 0 && a.cnt[ci] >= 2;
    float lse = 0.f, invp = 0.f, li = 0.f;
    if (anchor) {
        float f[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = a.F[(int64_t)i * DAD_H + lane + 32 * k];
        float m = -3.0e38f, sum = 0.f, possum = 0.f;
        int npos = 0;
        for (int j = 0; j < R; ++j) {
            const int cj = a.cls[j];
            if (cj < 0 || j == i) continue;  // warp-uniform
            float d = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) d = fmaf(f[k], a.F[(int64_t)j * DAD_H + lane + 32 * k], d);
            const float s = warp_sum(d) * a.inv_tau;
            if (s > m) { sum = sum * expf(m - s) + 1.f; m = s; }
            else sum += expf(s - m);
            if (cj == ci) { possum += s; ++npos; }
        }
        lse = m + logf(sum);
        invp = 1.f / (float)npos;
        li = lse - possum * invp;
    }
    if (lane == 0) { a.lse[i] = lse; a.invp[i] = invp; a.row_loss[i] = li; }
}

__device__ __forceinline__ int scl_n_valid(const int* cnt) {
    int nv = 0;
#pragma unroll
    for (int c = 0; c < DAD_C; ++c) nv += cnt[c] >= 2 ? cnt[c] : 0;
    return nv;
}

// gradient rows owned by this rank: local clean rows then local noisy rows
__global__ void __launch_bounds__(SCL_WARPS * 32) k_scl_grad(SclArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B = a.pack.B_local, Bg = B * a.pack.world, R = 2 * Bg;
    const int l = blockIdx.x * SCL_WARPS + warp;  // 0 .. 2B-1
    if (l >= 2 * B) return;
    const bool noisy_row = l >= B;
    const int li = noisy_row ? l - B : l;
    const int i = (noisy_row ? Bg : 0) + a.rank * B + li;
    const int ci = a.cls[i];
    const int nv = scl_n_valid(a.cnt);
    if (ci < 0 || nv == 0) return;
    const float inv_nv = 1.f / (float)nv;
    float f[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { f[k] = a.F[(int64_t)i * DAD_H + lane + 32 * k]; acc[k] = 0.f; }
    const float lse_i = a.lse[i], invp_i = a.invp[i];
    for (int j = 0; j < R; ++j) {
        const int cj = a.cls[j];
        if (cj < 0 || j == i) continue;
        float fj[8], d = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { fj[k] = a.F[(int64_t)j * DAD_H + lane + 32 * k]; d = fmaf(f[k], fj[k], d); }
        const float s = warp_sum(d) * a.inv_tau;
        const float same = cj == ci ? 1.f : 0.f;
        const float invp_j = a.invp[j];
        float c = 0.f;
        if (invp_i > 0.f) c += expf(s - lse_i) - invp_i * same;       // i as anchor, j as contrast
        if (invp_j > 0.f) c += expf(s - a.lse[j]) - invp_j * same;    // j as anchor, i as contrast
        c *= inv_nv;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(c, fj[k], acc[k]);
    }
    // d/dz = (I - f f^T) (dL/df) / |z|,  dL/df = acc / tau
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc[k] *= a.inv_tau; dot = fmaf(acc[k], f[k], dot); }
    dot = warp_sum(dot);
    const float scale = a.w_scl / a.nrm[i];
    float* g = (noisy_row ? a.g_strong : a.g_clean) + (int64_t)li * DAD_H;
#pragma unroll
    for (int k = 0; k < 8; ++k) g[lane + 32 * k] += scale * (acc[k] - f[k] * dot);
}

// this rank's loss share: fixed-order sum of L_i over the rows it owns, / n_valid
__global__ void __launch_bounds__(256) k_scl_loss(SclArgs a) {
    __shared__ float s_red[256];
    const int B = a.pack.B_local, Bg = B * a.pack.world;
    float s = 0.f;
    for (int l = threadIdx.x; l < 2 * B; l += 256) {
        const int i = (l >= B ? Bg + a.rank * B + (l - B) : a.rank * B + l);
        s += a.row_loss[i];
    }
    s_red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int nv = scl_n_valid(a.cnt);
        a.loss_out[3] = nv > 0 ? s_red[0] / (float)nv : 0.f;
    }
}

}  // namespace

int launch_scl(const SclArgs& a, cudaStream_t st) {
    const int B = a.pack.B_local, R = 2 * B * a.pack.world;
    cudaError_t e = cudaMemsetAsync(a.cnt, 0, DAD_C * sizeof(int), st);
    if (e != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "scl: memset: %s", cudaGetErrorString(e));
    k_scl_prep<<<ceil_div(R, SCL_WARPS), SCL_WARPS * 32, 0, st>>>(a);
    DAD_CHECK_LAUNCH("k_scl_prep");
    k_scl_rows<<<ceil_div(R, SCL_WARPS), SCL_WARPS * 32, 0, st>>>(a);
    DAD_CHECK_LAUNCH("k_scl_rows");
    k_scl_grad<<<ceil_div(2 * B, SCL_WARPS), SCL_WARPS * 32, 0, st>>>(a);
    DAD_CHECK_LAUNCH("k_scl_grad");
    k_scl_loss<<<1, 256, 0, st>>>(a);
    DAD_CHECK_LAUNCH("k_scl_loss");
    return DAD_OK;
}
