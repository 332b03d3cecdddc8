// ECDA: class-aware, attention-weighted 5-kernel Gaussian MMD + compactness + centroid repulsion, forward and
// analytic backward to d(loss)/d(embedding), in fixed-shape masked form.
// Reference: I/utils.py:510-652 (ECDALoss._gaussian_kernel / forward), gate I/train.py:444,450.
//
//  k_ecda_prep : grid over the 2*B_global rows (clean rows then noisy rows of the pack) -> per-(domain,class)
//                partial sums; the last CTA reduces them in fixed order (fp64) into class statistics (counts,
//                bandwidths via the centred closed form  sum_ij|zi-zj|^2 = 2n sum_i|zi-m|^2, centroids, compactness,
//                repulsion + its gradient, attention), builds the per-class member lists and the row-tile table.
//  k_ecda_pairs: one CTA per 32-row tile of one class; loops over all members of the class (32 columns at a
//                time, staged in shared memory): L2 by direct differences (as the reference), sum of 5 Gaussians,
//                weighted block sums (loss) and P = coef*K' (gradient); each row is owned by exactly one CTA, so
//                d/dz rows are accumulated in registers and added to g without atomics.  Under data parallelism a
//                rank launches tiles only for the rows it owns (its local clean / noisy samples).
#include "dad_common.cuh"
#include "dad_pack.cuh"

namespace {

constexpr int EP_THREADS = 256;
constexpr int ZS = DAD_H + 4;  // padded smem row stride (floats): conflict-free float4 access (see k_ecda_pairs)

__device__ __forceinline__ double block_sum_d(double v, double* s_red) {
    v = warp_sum_d(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < EP_THREADS / 32; ++w) t += s_red[w];
    return t;
}

// row r in [0, 2*Bg): r < Bg clean sample r, else noisy sample r-Bg.  Returns slot = dom*4+cls or -1.
__device__ __forceinline__ int row_slot(const EcdaArgs& a, int r, int Bg, float& omega, const float*& z) {
    if (r < Bg) {
        omega = 1.f;
        z = pack_emb_clean(a.pack, r);
        const int y = pack_label(a.pack, r);
        if (y < 0) return -1;  // null sample (padding of a ragged batch)
        return a.class_aware ? y : 0;
    }
    const int gi = r - Bg;
    z = pack_emb_strong(a.pack, gi);
    if (!a.mask[gi]) return -1;
    omega = a.class_aware ? pack_score(a.pack, gi) : 1.f;
    return 4 + (a.class_aware ? pack_pred(a.pack, gi) : 0);
}

__device__ __forceinline__ int list_lower_bound(const int* lists, int lo, int hi, int key) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldcg(lists + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

}  // namespace

__global__ void __launch_bounds__(EP_THREADS) k_ecda_prep(EcdaArgs a) {
    __shared__ double s_red[EP_THREADS / 32];
    __shared__ int s_last;
    __shared__ int s_cnt[8], s_off[8], s_wc[EP_THREADS / 32][8];
    __shared__ double s_sq[8], s_ws[8];
    __shared__ float s_dist[ECDA_MAXC][ECDA_MAXC];
    __shared__ int s_slot[EP_THREADS];
    __shared__ float s_om[EP_THREADS];
    const int tid = threadIdx.x, G = gridDim.x;
    const int Bg = a.pack.B_local * a.pack.world;
    const int R = 2 * Bg;
    const bool active = a.nmask[0] > 1;

    // ---------------- phase 1: partial sums of this CTA's row chunk (thread = feature h)
    float vec[8];
    double sq[8], wsum[8];
    int cnt[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) { vec[s] = 0.f; sq[s] = 0.0; wsum[s] = 0.0; cnt[s] = 0; }
    if (active) {
        const int chunk = ceil_div(R, G);  // <= EP_THREADS (host guarantees)
        const int r0 = blockIdx.x * chunk, r1 = min(R, r0 + chunk);
        if (r0 + tid < r1) {  // classify the chunk's rows in parallel first (no dependent-load chain in the loop below)
            float om = 0.f;
            const float* z;
            s_slot[tid] = row_slot(a, r0 + tid, Bg, om, z);
            s_om[tid] = om;
        }
        __syncthreads();
        for (int r = r0; r < r1; ++r) {
            const int slot = s_slot[r - r0];
            if (slot < 0) continue;
            const float omega = s_om[r - r0];
            const float* z = r < Bg ? pack_emb_clean(a.pack, r) : pack_emb_strong(a.pack, r - Bg);
            const float v = z[tid];
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if (s == slot) { vec[s] += v; sq[s] += (double)v * (double)v; wsum[s] += (double)omega; cnt[s] += 1; }
        }
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        a.part_vec[((int64_t)blockIdx.x * 8 + s) * DAD_H + tid] = vec[s];
        const double t = block_sum_d(sq[s], s_red);
        if (tid == 0) {
            a.part_sc[((int64_t)blockIdx.x * 8 + s) * 2 + 0] = t;
            a.part_sc[((int64_t)blockIdx.x * 8 + s) * 2 + 1] = wsum[s];
            a.part_cnt[blockIdx.x * 8 + s] = cnt[s];
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&a.counter[0], 1u);
        s_last = (prev == (unsigned)G - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // ---------------- phase 2 (last CTA): class statistics
    EcdaStats* st = a.stats;
    if (tid == 0) { a.counter[0] = 0u; st->active = active ? 1 : 0; }
    if (!active) {
        if (tid == 0) { st->n_tiles = 0; st->any_feasible = 0; st->rep = 0.f; st->rho = 0.f; }
        return;
    }
    double tv[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        double t = 0.0;
        for (int g = 0; g < G; ++g) t += (double)__ldcg(&a.part_vec[((int64_t)g * 8 + s) * DAD_H + tid]);
        tv[s] = t;
    }
    if (tid < 8) {
        double q = 0.0, w = 0.0;
        int c = 0;
        for (int g = 0; g < G; ++g) {
            q += __ldcg(&a.part_sc[((int64_t)g * 8 + tid) * 2 + 0]);
            w += __ldcg(&a.part_sc[((int64_t)g * 8 + tid) * 2 + 1]);
            c += __ldcg(&a.part_cnt[g * 8 + tid]);
        }
        s_sq[tid] = q; s_ws[tid] = w; s_cnt[tid] = c;
    }
    __syncthreads();
    const int NC = a.class_aware ? ECDA_MAXC : 1;
    float wmean = 0.f;
    for (int c = 0; c < ECDA_MAXC; ++c) wmean += a.w_class[c];
    wmean *= 0.25f;
    double mu[ECDA_MAXC];  // this thread's component of the noisy centroid of each class
    double rho = 0.0;
    for (int c = 0; c < ECDA_MAXC; ++c) {
        const int ns = (c < NC) ? s_cnt[c] : 0, nt = (c < NC) ? s_cnt[4 + c] : 0;
        const int n = ns + nt;
        const double m = n > 0 ? (tv[c] + tv[4 + c]) / (double)n : 0.0;
        mu[c] = nt > 0 ? tv[4 + c] / (double)nt : 0.0;
        const double m2 = block_sum_d(m * m, s_red);
        const double mu2 = block_sum_d(mu[c] * mu[c], s_red);
        if (tid == 0) {
            const bool feas = ns >= 2 && nt >= 2;
            st->n_s[c] = ns; st->n_t[c] = nt; st->feasible[c] = feas ? 1 : 0;
            double centred = s_sq[c] + s_sq[4 + c] - (double)n * m2;
            if (centred < 0.0) centred = 0.0;
            double bw = n > 1 ? 2.0 * (double)n * centred / ((double)n * n - (double)n) : 1.0;  // utils.py:540
            bw /= 4.0;                                                                          // kernel_mul^(kernel_num//2)
            for (int r = 0; r < 5; ++r) st->inv_beta[c][r] = (float)(1.0 / (bw * (double)(1 << r) + 1e-8));
            const double wt = a.class_aware ? s_ws[4 + c] : (double)nt;
            st->c_ss[c] = (float)(1.0 / ((double)ns * ns + 1e-8));
            st->c_tt[c] = (float)(1.0 / (wt * wt + 1e-8));
            st->c_st[c] = (float)(1.0 / ((double)ns * wt + 1e-8));
            const float att = a.class_aware ? expf(a.lambda_cls * (wmean - a.w_class[c])) : 1.f;  // utils.py:598-599
            st->att[c] = att;
            double comp = nt > 0 ? s_sq[4 + c] / (double)nt - mu2 : 0.0;  // mean |z - mu|^2
            st->comp[c] = (float)(comp > 0.0 ? comp : 0.0);
        }
        a.mu_t[c * DAD_H + tid] = (float)mu[c];
    }
    __syncthreads();
    for (int c = 0; c < NC; ++c)
        if (st->feasible[c]) rho += (double)st->att[c];
    rho *= (double)a.delta_rep;
    // repulsion: -mean pairwise distance between noisy centroids of classes with >= 1 masked member (utils.py:582-595)
    int nvalid = 0;
    for (int c = 0; c < NC; ++c) nvalid += (s_cnt[4 + c] > 0);
    double rep = 0.0;
    for (int c = 0; c < ECDA_MAXC; ++c)
        for (int d = c + 1; d < ECDA_MAXC; ++d) {
            const bool on = a.class_aware && c < NC && d < NC && s_cnt[4 + c] > 0 && s_cnt[4 + d] > 0;
            const double df = on ? mu[c] - mu[d] : 0.0;
            const double d2 = block_sum_d(df * df, s_red);
            if (tid == 0) { s_dist[c][d] = s_dist[d][c] = (float)sqrt(d2); }
            if (on) rep += sqrt(d2);
        }
    __syncthreads();
    const int npairs = nvalid * (nvalid - 1) / 2;
    rep = (a.class_aware && nvalid > 1) ? -rep / (double)npairs : 0.0;
    for (int c = 0; c < ECDA_MAXC; ++c) {
        double g = 0.0;
        if (a.class_aware && nvalid > 1 && c < NC && s_cnt[4 + c] > 0) {
            for (int d = 0; d < NC; ++d)
                if (d != c && s_cnt[4 + d] > 0 && s_dist[c][d] > 0.f) g += (mu[c] - mu[d]) / (double)s_dist[c][d];
            g = -g / (double)npairs / (double)s_cnt[4 + c];
        }
        a.rep_grad[c * DAD_H + tid] = (float)(rho * g);
    }
    if (tid == 0) {
        st->rep = (float)rep;
        st->rho = (float)rho;
        st->n_classes = NC;
        int any = 0, off = 0;
        for (int c = 0; c < ECDA_MAXC; ++c) {
            st->list_off[c] = off;
            off += s_cnt[c] + s_cnt[4 + c];
            any |= st->feasible[c];
        }
        st->any_feasible = any;
    }
    // ---------------- phase 3: ordered member lists (per class: clean ascending, then noisy ascending)
    if (tid < 8) {
        const int c = tid & 3, dom = tid >> 2;
        int off = 0;
        for (int cc = 0; cc < c; ++cc) off += s_cnt[cc] + s_cnt[4 + cc];
        s_off[tid] = off + (dom ? s_cnt[c] : 0);
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int base = 0; base < R; base += EP_THREADS) {
        const int r = base + tid;
        int slot = -1;
        if (r < R) {
            float om;
            const float* z;
            slot = row_slot(a, r, Bg, om, z);
        }
        unsigned myballot = 0;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const unsigned bal = __ballot_sync(0xffffffffu, slot == s);
            if (lane == 0) s_wc[warp][s] = __popc(bal);
            if (slot == s) myballot = bal;
        }
        __syncthreads();
        if (slot >= 0) {
            int pos = s_off[slot] + __popc(myballot & ((1u << lane) - 1u));
            for (int w = 0; w < warp; ++w) pos += s_wc[w][slot];
            a.lists[pos] = r;
        }
        __syncthreads();
        if (tid < 8) {
            int add = 0;
            for (int w = 0; w < EP_THREADS / 32; ++w) add += s_wc[w][tid];
            s_off[tid] += add;
        }
        __syncthreads();
    }
    // ---------------- phase 4: row tiles for the rows this rank owns
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int lo = a.rank * a.pack.B_local, hi = lo + a.pack.B_local;
        int nt_ = 0;
        for (int c = 0; c < NC; ++c) {
            const int off = st->list_off[c], ns = s_cnt[c], ntc = s_cnt[4 + c];
            const bool feas = st->feasible[c] != 0;
            for (int dom = 0; dom < 2; ++dom) {
                if (dom == 0 && !feas) continue;                              // clean rows only matter through the MMD
                if (dom == 1 && !feas && !(a.class_aware && nvalid > 1)) continue;  // noisy rows: MMD/comp or repulsion
                const int b0 = off + (dom ? ns : 0), cntd = dom ? ntc : ns;
                const int shift = dom ? Bg : 0;
                // [first,last): members owned by this rank (lists are ascending) -> two lower bounds
                const int first = list_lower_bound(a.lists, b0, b0 + cntd, lo + shift);
                const int last = list_lower_bound(a.lists, b0, b0 + cntd, hi + shift);
                for (int p = first; p < last && nt_ < a.max_tiles; p += ECDA_TR) {
                    EcdaTile t;
                    t.cls = c; t.row_begin = p; t.row_count = min(ECDA_TR, last - p); t.feasible = feas ? 1 : 0;
                    a.tiles[nt_++] = t;
                }
            }
        }
        st->n_tiles = nt_;
    }
}

__global__ void __launch_bounds__(EP_THREADS) k_ecda_pairs(EcdaArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* zi = smem;                      // [32][ZS]
    float* zj = zi + ECDA_TR * ZS;         // [32][ZS]
    float* Pm = zj + ECDA_TR * ZS;         // [32][33]
    float* wi = Pm + ECDA_TR * 33;         // [32] omega_i
    float* wj = wi + ECDA_TR;              // [32] omega_j
    int* ri = reinterpret_cast<int*>(wj + ECDA_TR);  // [32] row ids
    int* dj = ri + ECDA_TR;                // [32] column domain (0 clean, 1 noisy, -1 invalid)
    __shared__ float s_red[3][EP_THREADS / 32];
    __shared__ int s_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const EcdaStats* st = a.stats;
    const int Bg = a.pack.B_local * a.pack.world;
    const int n_tiles = st->n_tiles;
    float sums[3] = {0.f, 0.f, 0.f};

    if ((int)blockIdx.x < n_tiles) {
        const EcdaTile t = a.tiles[blockIdx.x];
        const int c = t.cls;
        const int off = st->list_off[c], ns = st->n_s[c], n = ns + st->n_t[c];
        const int dom_i = (t.row_begin - off) >= ns ? 1 : 0;
        // ---- stage the row tile
        if (tid < ECDA_TR) {
            int r = -1;
            float om = 0.f;
            if (tid < t.row_count) {
                r = a.lists[t.row_begin + tid];
                om = (dom_i && a.class_aware) ? pack_score(a.pack, r - Bg) : 1.f;
            }
            ri[tid] = r;
            wi[tid] = om;
        }
        __syncthreads();
        for (int e = tid; e < ECDA_TR * (DAD_H / 4); e += EP_THREADS) {
            const int row = e / (DAD_H / 4), q = e - row * (DAD_H / 4);
            const int r = ri[row];
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r >= 0) v = *reinterpret_cast<const float4*>((r < Bg ? pack_emb_clean(a.pack, r) : pack_emb_strong(a.pack, r - Bg)) + q * 4);
            *reinterpret_cast<float4*>(zi + row * ZS + q * 4) = v;
        }
        __syncthreads();
        // phase-B ownership: rows warp*4..+3, features lane+32m
        float ziv[4][8], acc[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int m = 0; m < 8; ++m) { ziv[r][m] = zi[(warp * 4 + r) * ZS + lane + 32 * m]; acc[r][m] = 0.f; }

        if (t.feasible) {
            const float ib0 = st->inv_beta[c][0], ib1 = st->inv_beta[c][1], ib2 = st->inv_beta[c][2], ib3 = st->inv_beta[c][3], ib4 = st->inv_beta[c][4];
            const float css = st->c_ss[c], ctt = st->c_tt[c], cst = st->c_st[c];
            const float gscale = -4.f * st->att[c] * a.w_ecda;
            const int ti = tid >> 3, tj = tid & 7;  // phase A: row ti, columns tj + 8m
            for (int j0 = 0; j0 < n; j0 += ECDA_TR) {
                __syncthreads();  // previous iteration's readers of zj / Pm are done
                if (tid < ECDA_TR) {
                    int d = -1;
                    float om = 0.f;
                    if (j0 + tid < n) {
                        const int r = a.lists[off + j0 + tid];
                        d = (r >= Bg) ? 1 : 0;
                        om = (d && a.class_aware) ? pack_score(a.pack, r - Bg) : 1.f;
                    }
                    dj[tid] = d;
                    wj[tid] = om;
                }
                for (int e = tid; e < ECDA_TR * (DAD_H / 4); e += EP_THREADS) {
                    const int row = e / (DAD_H / 4), q = e - row * (DAD_H / 4);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j0 + row < n) {
                        const int r = a.lists[off + j0 + row];
                        v = *reinterpret_cast<const float4*>((r < Bg ? pack_emb_clean(a.pack, r) : pack_emb_strong(a.pack, r - Bg)) + q * 4);
                    }
                    *reinterpret_cast<float4*>(zj + row * ZS + q * 4) = v;
                }
                __syncthreads();
                // ---- phase A: squared distances by direct differences (I/utils.py:537), kernels, P
                float l2[4] = {0.f, 0.f, 0.f, 0.f};
                const float* zrow = zi + ti * ZS;
#pragma unroll 4
                for (int h = 0; h < DAD_H; h += 4) {
                    const float4 x = *reinterpret_cast<const float4*>(zrow + h);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const float4 y = *reinterpret_cast<const float4*>(zj + (tj + 8 * m) * ZS + h);
                        float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
                        l2[m] = fmaf(d0, d0, l2[m]); l2[m] = fmaf(d1, d1, l2[m]); l2[m] = fmaf(d2, d2, l2[m]); l2[m] = fmaf(d3, d3, l2[m]);
                    }
                }
                const bool i_ok = ti < t.row_count;
                const float om_i = wi[ti];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int j = tj + 8 * m;
                    const int d = dj[j];
                    float p = 0.f;
                    if (i_ok && d >= 0) {
                        const float e0 = expf(-l2[m] * ib0), e1 = expf(-l2[m] * ib1), e2 = expf(-l2[m] * ib2), e3 = expf(-l2[m] * ib3), e4 = expf(-l2[m] * ib4);
                        const float K = e0 + e1 + e2 + e3 + e4;
                        const float Kp = e0 * ib0 + e1 * ib1 + e2 * ib2 + e3 * ib3 + e4 * ib4;
                        const float ww = om_i * wj[j];
                        float s;
                        if (dom_i == 0 && d == 0) { sums[0] += ww * K; s = css; }
                        else if (dom_i == 1 && d == 1) { sums[1] += ww * K; s = ctt; }
                        else { if (dom_i == 0) sums[2] += ww * K; s = -cst; }
                        p = gscale * s * ww * Kp;
                    }
                    Pm[ti * 33 + j] = p;
                }
                __syncthreads();
                // ---- phase B: g_i += sum_j P_ij (z_i - z_j)
#pragma unroll 2
                for (int j = 0; j < ECDA_TR; ++j) {
                    float zjv[8];
#pragma unroll
                    for (int m = 0; m < 8; ++m) zjv[m] = zj[j * ZS + lane + 32 * m];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float p = Pm[(warp * 4 + r) * 33 + j];
#pragma unroll
                        for (int m = 0; m < 8; ++m) acc[r][m] = fmaf(p, ziv[r][m] - zjv[m], acc[r][m]);
                    }
                }
            }
        }
        // ---- compactness + repulsion gradients (noisy rows), then accumulate into g
        const float comp_k = (t.feasible && dom_i) ? a.w_ecda * st->att[c] * a.gamma_comp * 2.f / (float)st->n_t[c] : 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = warp * 4 + r;
            const int rid = ri[row];
            if (rid < 0) continue;
            const int gi = rid < Bg ? rid : rid - Bg;
            float* g = (rid < Bg ? a.g_clean : a.g_strong) + (int64_t)(gi - a.rank * a.pack.B_local) * DAD_H;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int h = lane + 32 * m;
                float v = acc[r][m];
                if (dom_i && a.class_aware) v += comp_k * (ziv[r][m] - a.mu_t[c * DAD_H + h]) + a.w_ecda * a.rep_grad[c * DAD_H + h];
                g[h] += v;
            }
        }
    }
    // ---- tile partial sums -> fixed-order loss assembly by the last CTA
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float v = warp_sum(sums[k]);
        if (lane == 0) s_red[k][warp] = v;
    }
    __syncthreads();
    if (tid < 3 && (int)blockIdx.x < a.max_tiles) {
        float v = 0.f;
        for (int w = 0; w < EP_THREADS / 32; ++w) v += s_red[tid][w];
        a.tile_sums[blockIdx.x * 3 + tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&a.counter[1], 1u);
        s_last = (prev == gridDim.x - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // per-class block sums over this rank's tiles: thread-strided partials, then a fixed-order cross-thread sum
    float cs[ECDA_MAXC][3];
#pragma unroll
    for (int c = 0; c < ECDA_MAXC; ++c) cs[c][0] = cs[c][1] = cs[c][2] = 0.f;
    for (int tl = tid; tl < n_tiles; tl += EP_THREADS) {
        const int c = a.tiles[tl].cls;
        if (!a.tiles[tl].feasible) continue;
        const float v0 = __ldcg(&a.tile_sums[tl * 3 + 0]), v1 = __ldcg(&a.tile_sums[tl * 3 + 1]), v2 = __ldcg(&a.tile_sums[tl * 3 + 2]);
#pragma unroll
        for (int cc = 0; cc < ECDA_MAXC; ++cc)
            if (cc == c) { cs[cc][0] += v0; cs[cc][1] += v1; cs[cc][2] += v2; }
    }
    __shared__ float s_cs[ECDA_MAXC * 3][EP_THREADS / 32];
#pragma unroll
    for (int c = 0; c < ECDA_MAXC; ++c)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = warp_sum(cs[c][k]);
            if (lane == 0) s_cs[c * 3 + k][warp] = v;
        }
    __syncthreads();
    if (tid != 0) return;
    a.counter[1] = 0u;
    float loss = 0.f;
    if (st->active) {
        for (int c = 0; c < st->n_classes; ++c) {
            if (!st->feasible[c]) continue;
            float t3[3];
            for (int k = 0; k < 3; ++k) {
                float v = 0.f;
                for (int w = 0; w < EP_THREADS / 32; ++w) v += s_cs[c * 3 + k][w];
                t3[k] = v;
            }
            float lc = st->c_ss[c] * t3[0] + st->c_tt[c] * t3[1] - 2.f * st->c_st[c] * t3[2];
            if (a.rank == 0 && a.class_aware) lc += a.gamma_comp * st->comp[c] + a.delta_rep * st->rep;
            loss += st->att[c] * lc;
        }
    }
    a.loss_out[2] = loss;  // this rank's share; shares sum to the global-batch ECDA loss
}

size_t ecda_pairs_smem_bytes() { return (size_t)(2 * ECDA_TR * ZS + ECDA_TR * 33 + 2 * ECDA_TR) * sizeof(float) + 2 * ECDA_TR * sizeof(int); }
int ecda_prep_grid(int B_global) {
    // rows per CTA <= 256 (k_ecda_prep classifies its chunk with one thread per row); aim at ~16 rows per CTA
    int g = ceil_div(2 * B_global, 16);
    if (g > 296) g = 296;
    const int gmin = ceil_div(2 * B_global, EP_THREADS);
    if (g < gmin) g = gmin;
    return g < 1 ? 1 : g;
}
int ecda_max_tiles(int B_local) { return ceil_div(2 * B_local, ECDA_TR) + 2 * ECDA_MAXC; }

int launch_ecda(const EcdaArgs& a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_ecda_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ecda_pairs_smem_bytes());
        attr_set = true;
    }
    k_ecda_prep<<<a.prep_grid, EP_THREADS, 0, st>>>(a);
    DAD_CHECK_LAUNCH("k_ecda_prep");
    k_ecda_pairs<<<a.max_tiles, EP_THREADS, ecda_pairs_smem_bytes(), st>>>(a);
    DAD_CHECK_LAUNCH("k_ecda_pairs");
    return DAD_OK;
}
