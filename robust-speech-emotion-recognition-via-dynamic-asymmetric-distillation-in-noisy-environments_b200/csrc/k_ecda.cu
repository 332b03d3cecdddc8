// ECDA: class-aware, attention-weighted 5-kernel Gaussian MMD + compactness + centroid repulsion, forward and
// analytic backward to d(loss)/d(embedding), in fixed-shape masked form.
// Reference: I/utils.py:510-652 (ECDALoss._gaussian_kernel / forward), gate I/train.py:444,450.
//
//  k_ecda_prep : grid over the 2*B_global rows (clean rows then noisy rows of the pack) -> per-(domain,class)
//                partial sums;  k_ecda_reduce (8 CTAs, one per slot) folds them in fixed order (fp64);  k_ecda_stats
//                (one CTA) turns them into class statistics (counts, bandwidths via the centred closed form
//                sum_ij|zi-zj|^2 = 2n sum_i|zi-m|^2, centroids, compactness, repulsion + its gradient, attention) and
//                builds the per-class member lists and the row-tile table.
//  k_ecda_pack + k_gram_tc (k_gram_tc.cu): class-centred embeddings -> fp16 hi/lo, Gram G = Z' Z'^T on tcgen05 (rows this rank owns)
//  k_ecda_p    : elementwise over G: distances, 5 Gaussians, weighted block sums (loss) and the gradient weights P (fp16 hi/lo)
//  k_pz_tc     (k_gram_tc.cu): P Z' on tcgen05;   k_ecda_fold: d/dz rows;   k_ecda_loss: fixed-order loss assembly.
//                Under data parallelism a rank works only on the member rows it owns (its local clean / noisy samples).
#include "dad_common.cuh"
#include "dad_pack.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int EP_THREADS = 256;

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

// 1/n to fp64 accuracy without the fp64 divide (a ~35-instruction dependent chain on a pipe that issues one warp instruction every
// 16 cycles -- the divides of the class statistics used to be most of k_ecda_stats): fp32 reciprocal seed + two Newton steps.
__device__ __forceinline__ double recip_int(int n) {
    const double dn = (double)n;
    double r = (double)__frcp_rn((float)n);
    r = r * (2.0 - dn * r);
    r = r * (2.0 - dn * r);
    return r;
}

}  // namespace

__global__ void __launch_bounds__(EP_THREADS) k_ecda_prep(EcdaArgs a) {
    __shared__ double s_red[EP_THREADS / 32], s_redw[EP_THREADS / 32];
    __shared__ int s_redc[EP_THREADS / 32];
    __shared__ int s_slot[EP_THREADS];
    __shared__ float s_om[EP_THREADS];
    __shared__ double s_vec[8][EP_THREADS], s_sq[8][EP_THREADS];  // per-slot accumulators of thread tid (= feature h)
    __shared__ const float* s_ptr[EP_THREADS];                    // embedding row of each chunk row
    const int tid = threadIdx.x, G = gridDim.x;
    const int Bg = a.pack.B_local * a.pack.world;
    const int R = 2 * Bg;
    const bool active = a.nmask[0] > 1;

    // ---------------- partial sums of this CTA's row chunk (thread = feature h)
    // The chunk's rows are classified in parallel first (slot, weight, row pointer -> shared memory); the accumulation then walks
    // them in ascending order in batches of eight with all eight embedding loads in flight -- UNCONDITIONAL loads (a non-member
    // row is read and ignored): guarded loads end up in one branch region each together with their use, which serialises one
    // L2 round trip per row.  The sums are fp64 (the class bandwidth is a difference of two nearly equal moments) and live in
    // shared memory indexed by the row's slot: one add and one fma per row and thread.
#pragma unroll
    for (int s = 0; s < 8; ++s) { s_vec[s][tid] = 0.0; s_sq[s][tid] = 0.0; }
    s_slot[tid] = -1;  // threads beyond the chunk: no row
    s_om[tid] = 0.f;
    if (active) {
        const int chunk = ceil_div(R, G);  // <= EP_THREADS (host guarantees)
        const int r0 = blockIdx.x * chunk, r1 = min(R, r0 + chunk);
        if (r0 + tid < r1) {
            float om = 0.f;
            const float* z;
            s_slot[tid] = row_slot(a, r0 + tid, Bg, om, z);
            s_om[tid] = om;
            s_ptr[tid] = z;
            a.slot_of[r0 + tid] = (uint8_t)(s_slot[tid] < 0 ? 255 : s_slot[tid]);  // for the ordered scatter of k_ecda_stats
        }
        __syncthreads();
        const int nrow = r1 - r0;
        for (int rb = 0; rb < nrow; rb += 8) {
            float v8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v8[u] = __ldcg(s_ptr[min(rb + u, nrow - 1)] + tid);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int slot = rb + u < nrow ? s_slot[rb + u] : -1;
                if (slot < 0) continue;  // CTA-uniform
                const double v = (double)v8[u];
                s_vec[slot][tid] += v;
                s_sq[slot][tid] = fma(v, v, s_sq[slot][tid]);
            }
        }
    }
    // per-slot scalars: sum of squares over the 256 features (thread tid holds feature tid) and, over the chunk's rows (thread tid
    // classified row r0 + tid), the weight sum and the member count -- three reductions behind one barrier pair per slot
    const int my_slot = s_slot[tid];
    const double my_om = my_slot >= 0 ? (double)s_om[tid] : 0.0;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        a.part_vec[((int64_t)blockIdx.x * 8 + s) * DAD_H + tid] = (float)s_vec[s][tid];
        const double q = warp_sum_d(s_sq[s][tid]);
        const double w = warp_sum_d(my_slot == s ? my_om : 0.0);
        const int c = __popc(__ballot_sync(0xffffffffu, my_slot == s));
        __syncthreads();
        if ((tid & 31) == 0) { s_red[tid >> 5] = q; s_redw[tid >> 5] = w; s_redc[tid >> 5] = c; }
        __syncthreads();
        if (tid == 0) {
            double tq = 0.0, tw = 0.0;
            int tc = 0;
            for (int ww = 0; ww < EP_THREADS / 32; ++ww) { tq += s_red[ww]; tw += s_redw[ww]; tc += s_redc[ww]; }
            a.part_sc[((int64_t)blockIdx.x * 8 + s) * 2 + 0] = tq;
            a.part_sc[((int64_t)blockIdx.x * 8 + s) * 2 + 1] = tw;
            a.part_cnt[blockIdx.x * 8 + s] = tc;
        }
    }
}

// k_ecda_reduce: folds the G per-CTA partials of every slot in a fixed order.  Eight independent accumulators per thread keep
// eight L2 loads in flight; 32 CTAs (slot x feature quarter) work in parallel (as the tail of one "last CTA" this
// reduction alone was >0.1 ms at global batch 4096 -- a fixed cost on every data-parallel rank).
__global__ void __launch_bounds__(EP_THREADS) k_ecda_reduce(EcdaArgs a, int G) {
    // grid = 8 slots x 4 feature quarters; thread (f = tid & 63, part = tid >> 6) sums the part-th quarter of the g range
    // (8 loads in flight), the four quarters are then combined in fixed order through shared memory
    __shared__ double s_part[4][64];
    const int tid = threadIdx.x, s = blockIdx.x >> 2, f = (blockIdx.x & 3) * 64 + (tid & 63), part = tid >> 6;
    const int per = ceil_div(G, 4), g0 = part * per, g1 = min(G, g0 + per);
    double t8[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int g = g0;
    for (; g + 8 <= g1; g += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) t8[u] += (double)a.part_vec[((int64_t)(g + u) * 8 + s) * DAD_H + f];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)  // tail (static register indices: a runtime index would put t8 in local memory)
        if (g + u < g1) t8[u] += (double)a.part_vec[((int64_t)(g + u) * 8 + s) * DAD_H + f];
    s_part[part][tid & 63] = ((t8[0] + t8[1]) + (t8[2] + t8[3])) + ((t8[4] + t8[5]) + (t8[6] + t8[7]));
    __syncthreads();
    if (tid < 64) a.red_vec[s * DAD_H + f] = (s_part[0][tid] + s_part[1][tid]) + (s_part[2][tid] + s_part[3][tid]);
    if ((blockIdx.x & 3) == 0 && tid >= 64 && tid < 96) {  // scalar partials (one warp of the slot's first CTA)
        const int l = tid - 64;
        double q = 0.0, w = 0.0;
        int c = 0;
        for (int gg = l; gg < G; gg += 32) {
            q += a.part_sc[((int64_t)gg * 8 + s) * 2 + 0];
            w += a.part_sc[((int64_t)gg * 8 + s) * 2 + 1];
            c += a.part_cnt[gg * 8 + s];
        }
        q = warp_sum_d(q); w = warp_sum_d(w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (l == 0) { a.red_sc[s * 2 + 0] = q; a.red_sc[s * 2 + 1] = w; a.red_cnt[s] = c; }
    }
}

// k_ecda_stats (one CTA): class statistics, ordered member lists, row-tile table
__global__ void __launch_bounds__(EP_THREADS) k_ecda_stats(EcdaArgs a) {
    static_assert(ECDA_MAXC == 4, "six centroid pairs below");
    constexpr int NW = EP_THREADS / 32;
    constexpr int NBS = 2 * ECDA_MAXC + 6;  // block sums: |m_c|^2, |mu_c|^2 per class, |mu_c - mu_d|^2 per centroid pair
    __shared__ double s_bs[NBS][NW], s_bt[NBS];
    __shared__ int s_cnt[8], s_off[8], s_wc[NW][8];
    __shared__ double s_sq[8], s_ws[8], s_rcp[8];
    __shared__ double s_invd[ECDA_MAXC][ECDA_MAXC], s_pd[6], s_gs[ECDA_MAXC];
    __shared__ float s_att[ECDA_MAXC];
    __shared__ int s_feas[ECDA_MAXC], s_loff[ECDA_MAXC];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Bg = a.pack.B_local * a.pack.world;
    const int R = 2 * Bg;
    const bool active = a.nmask[0] > 1;

    // ---------------- phase 2: class statistics
    // fp64 issues at 1/16 rate and an fp64 divide is a ~35-instruction dependent chain, so the arithmetic is laid out for
    // latency: reciprocals of the counts once (recip_int), all 14 block sums behind ONE barrier pair, the per-class scalars on
    // four threads of four different warps, the centroid-pair scalars on six more; divides that the reference itself does in
    // fp32 (1/beta_r, the block normalisers) stay fp32.  (One thread doing ~40 fp64 divides in sequence was ~25 us.)
    EcdaStats* st = a.stats;
    if (tid == 0) st->active = active ? 1 : 0;
    if (!active) {
        if (tid == 0) { st->n_tiles = 0; st->n_gram_items = 0; st->n_pz_items = 0; st->p_split = 1; st->any_feasible = 0; st->rep = 0.f; st->rho = 0.f; }
        return;
    }
    const int NC = a.class_aware ? ECDA_MAXC : 1;
    double tv[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) tv[s] = a.red_vec[s * DAD_H + tid];
    if (tid < 8) {
        s_sq[tid] = a.red_sc[tid * 2 + 0]; s_ws[tid] = a.red_sc[tid * 2 + 1]; s_cnt[tid] = a.red_cnt[tid];
        const int c = tid & 3;
        const int ns = c < NC ? a.red_cnt[c] : 0, nt = c < NC ? a.red_cnt[4 + c] : 0;
        const int n = tid < 4 ? ns + nt : nt;  // s_rcp[c] = 1/(n_s+n_t), s_rcp[4+c] = 1/n_t  (0 when empty)
        s_rcp[tid] = n > 0 ? recip_int(n) : 0.0;
    }
    __syncthreads();
    int nvalid = 0;  // classes with >= 1 masked noisy member
    for (int c = 0; c < NC; ++c) nvalid += (s_cnt[4 + c] > 0);
    const int npairs = nvalid * (nvalid - 1) / 2;
    double m[ECDA_MAXC], mu[ECDA_MAXC], bsv[NBS];  // this thread's component of the class mean / the noisy centroid
#pragma unroll
    for (int c = 0; c < ECDA_MAXC; ++c) {
        m[c] = c < NC ? (tv[c] + tv[4 + c]) * s_rcp[c] : 0.0;
        mu[c] = c < NC ? tv[4 + c] * s_rcp[4 + c] : 0.0;
        bsv[c] = m[c] * m[c];
        bsv[4 + c] = mu[c] * mu[c];
    }
    {
        int p = 0;
#pragma unroll
        for (int c = 0; c < ECDA_MAXC; ++c)
#pragma unroll
            for (int d = c + 1; d < ECDA_MAXC; ++d, ++p) {
                const bool on = a.class_aware && c < NC && d < NC && s_cnt[4 + c] > 0 && s_cnt[4 + d] > 0;
                const double df = on ? mu[c] - mu[d] : 0.0;
                bsv[8 + p] = df * df;
            }
    }
#pragma unroll
    for (int k = 0; k < NBS; ++k) {
        const double v = warp_sum_d(bsv[k]);
        if (lane == 0) s_bs[k][warp] = v;
    }
    __syncthreads();
    if (tid < NBS) {
        double t = 0.0;
        for (int w = 0; w < NW; ++w) t += s_bs[tid][w];
        s_bt[tid] = t;
    }
    __syncthreads();
    float wmean = 0.f;
    for (int c = 0; c < ECDA_MAXC; ++c) wmean += a.w_class[c];
    wmean *= 0.25f;
    if (lane == 0 && warp < ECDA_MAXC) {  // per-class scalars: class c on warp c
        const int c = warp;
        const int ns = (c < NC) ? s_cnt[c] : 0, nt = (c < NC) ? s_cnt[4 + c] : 0;
        const int n = ns + nt;
        const double m2 = s_bt[c], mu2 = s_bt[4 + c];
        const bool feas = ns >= 2 && nt >= 2;
        st->n_s[c] = ns; st->n_t[c] = nt; st->feasible[c] = feas ? 1 : 0;
        s_feas[c] = feas ? 1 : 0;
        double centred = s_sq[c] + s_sq[4 + c] - (double)n * m2;
        if (centred < 0.0) centred = 0.0;
        double bw = n > 1 ? 2.0 * (double)n * centred / ((double)n * n - (double)n) : 1.0;  // utils.py:540
        bw *= 0.25;                                                                          // / kernel_mul^(kernel_num//2)
        const float bwf = (float)bw;
        float sb = 0.f;
        for (int r = 0; r < 5; ++r) {
            const float ib = 1.0f / (bwf * (float)(1 << r) + 1e-8f);
            st->inv_beta[c][r] = ib;
            sb += ib;
        }
        const double wt = a.class_aware ? s_ws[4 + c] : (double)nt;
        const float css = 1.0f / ((float)ns * (float)ns + 1e-8f);
        const float ctt = 1.0f / (float)(wt * wt + 1e-8);
        const float cst = 1.0f / (float)((double)ns * wt + 1e-8);
        st->c_ss[c] = css; st->c_tt[c] = ctt; st->c_st[c] = cst;
        const float att = a.class_aware ? expf(a.lambda_cls * (wmean - a.w_class[c])) : 1.f;  // utils.py:598-599
        st->att[c] = att;
        s_att[c] = att;
        // power-of-two operand scale of this class's gradient weights P = coef K' before their fp16 hi/lo split (k_ecda_p):
        // |P| <= 4 att w_ecda max(c_ss, c_tt, c_st) sum_r 1/beta_r (scores <= 1, exp <= 1); the largest value lands below 2^13
        {
            const float cmax = fmaxf(css, fmaxf(ctt, cst));
            const float pmax = 4.f * att * fabsf(a.w_ecda) * cmax * sb;
            float sc = 1.0f;
            if (pmax > 0.f && pmax < 3.0e38f) {
                int e = (int)((__float_as_uint(pmax) >> 23) & 0xFFu) - 127;  // floor(log2 pmax)
                int se = 12 - e;
                se = se > 100 ? 100 : (se < -100 ? -100 : se);
                sc = __uint_as_float((uint32_t)(127 + se) << 23);
            }
            st->p_scale[c] = sc;
        }
        const double comp = nt > 0 ? s_sq[4 + c] * s_rcp[4 + c] - mu2 : 0.0;  // mean |z - mu|^2
        st->comp[c] = (float)(comp > 0.0 ? comp : 0.0);
        // d rep / d mu_c scale: -1 / (npairs * n_t)   (utils.py:582-595 differentiated)
        s_gs[c] = (a.class_aware && nvalid > 1 && c < NC && nt > 0) ? -1.0 / ((double)npairs * (double)nt) : 0.0;
    }
    if (warp >= 4 && lane < 2) {  // centroid pairs: pair p on warp 4 + (p & 3), lane p >> 2
        const int p = (warp - 4) + 4 * lane;
        if (p < 6) {
            const int c = p < 3 ? 0 : (p < 5 ? 1 : 2), d = p < 3 ? p + 1 : (p < 5 ? p - 1 : 3);
            const bool on = a.class_aware && c < NC && d < NC && s_cnt[4 + c] > 0 && s_cnt[4 + d] > 0;
            const double dist = sqrt(s_bt[8 + p]);
            const float distf = (float)dist;
            s_pd[p] = on ? dist : 0.0;
            const double inv = (on && distf > 0.f) ? 1.0 / (double)distf : 0.0;
            s_invd[c][d] = inv;
            s_invd[d][c] = inv;
        }
    }
    __syncthreads();
    // repulsion: -mean pairwise distance between noisy centroids of classes with >= 1 masked member (utils.py:582-595)
    double rho = 0.0;
    for (int c = 0; c < NC; ++c)
        if (s_feas[c]) rho += (double)s_att[c];
    rho *= (double)a.delta_rep;
#pragma unroll
    for (int c = 0; c < ECDA_MAXC; ++c) {
        double g = 0.0;
        if (a.class_aware && nvalid > 1 && c < NC && s_cnt[4 + c] > 0) {
#pragma unroll
            for (int d = 0; d < ECDA_MAXC; ++d)
                if (d != c && d < NC && s_cnt[4 + d] > 0) g += (mu[c] - mu[d]) * s_invd[c][d];
            g *= s_gs[c];
        }
        a.rep_grad[c * DAD_H + tid] = (float)(rho * g);
        a.mu_t[c * DAD_H + tid] = (float)mu[c];
        a.mu_all[c * DAD_H + tid] = (float)m[c];
    }
    if (tid == 0) {
        double rep = 0.0;
        for (int p = 0; p < 6; ++p) rep += s_pd[p];
        rep = (a.class_aware && nvalid > 1) ? -rep / (double)npairs : 0.0;
        st->rep = (float)rep;
        st->rho = (float)rho;
        st->n_classes = NC;
        int any = 0, off = 0;
        for (int c = 0; c < ECDA_MAXC; ++c) {
            st->list_off[c] = off;
            s_loff[c] = off;
            off += s_cnt[c] + s_cnt[4 + c];
            any |= s_feas[c];
        }
        st->any_feasible = any;
    }
    // ---------------- phase 3: ordered member lists (per class: clean ascending, then noisy ascending)
    // All row slots were published by phase 1; one coalesced pass brings them into shared memory (every load in flight at
    // once), so the ordered scatter below touches no global memory on its critical path.  (The first version re-derived each
    // slot from the pack inside the loop -- 32 dependent global round trips -- and then binary-searched the lists it had just
    // written with L2-latency probes from one thread: together ~0.15 ms of a single CTA at global batch 4096.)
    extern __shared__ __align__(16) uint8_t s_sl[];  // [R], padded to 16 bytes
    __shared__ int s_lt_lo[8], s_lt_hi[8];
    if (tid < 8) {
        const int c = tid & 3, dom = tid >> 2;
        int off = 0;
        for (int cc = 0; cc < c; ++cc) off += s_cnt[cc] + s_cnt[4 + cc];
        s_off[tid] = off + (dom ? s_cnt[c] : 0);
        s_lt_lo[tid] = 0; s_lt_hi[tid] = 0;
    }
    {   // 16 rows per load (the table is 256-byte aligned and padded to a multiple of 16 bytes)
        const uint4* src = reinterpret_cast<const uint4*>(a.slot_of);
        uint4* dst = reinterpret_cast<uint4*>(s_sl);
        for (int q = tid; q < (R + 15) / 16; q += EP_THREADS) dst[q] = __ldcg(src + q);
    }
    __syncthreads();
    const int own_lo = a.rank * a.pack.B_local, own_hi = own_lo + a.pack.B_local;
    for (int base = 0; base < R; base += EP_THREADS) {
        const int r = base + tid;
        int slot = -1;
        if (r < R) { const int v = s_sl[r]; slot = v == 255 ? -1 : v; }
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
            // members of this slot that precede the rows this rank owns / precede the end of them (lists are ascending)
            const int id = slot >= 4 ? r - Bg : r;
            if (id < own_lo) atomicAdd(&s_lt_lo[slot], 1);
            if (id < own_hi) atomicAdd(&s_lt_hi[slot], 1);
        }
        __syncthreads();
        if (tid < 8) {
            int add = 0;
            for (int w = 0; w < EP_THREADS / 32; ++w) add += s_wc[w][tid];
            s_off[tid] += add;
        }
        __syncthreads();
    }
    // ---------------- phase 4: work tables for the rows this rank owns (fold tiles of 64 rows, Gram items, (P Z') items)
    // Thread 0 lays out the <= 8 (class, domain) segments (first / last owned position, bases into G and into the tables); then
    // all threads emit the entries in parallel (a serial loop over ~700 entries cost 30 us of this single CTA).
    __threadfence();
    __syncthreads();
    __shared__ int sg_first[8], sg_last[8], sg_grow[8], sg_tile0[8], sg_item0[8], sg_pz0[8], sg_feas[8], sg_on[8], s_tot[4];
    if (tid == 0) {
        int nt_ = 0, ng_ = 0, np_ = 0, g_rows = 0;
        for (int sgi = 0; sgi < 8; ++sgi) {
            const int c = sgi >> 1, dom = sgi & 1;
            sg_on[sgi] = 0;
            if (c >= NC) continue;
            const int off = s_loff[c], ns = s_cnt[c], n_c = s_cnt[c] + s_cnt[4 + c];
            const bool feas = s_feas[c] != 0;
            if (dom == 0 && !feas) continue;                              // clean rows only matter through the MMD
            if (dom == 1 && !feas && !(a.class_aware && nvalid > 1)) continue;  // noisy rows: MMD/comp or repulsion
            const int b0 = off + (dom ? ns : 0);
            const int first = b0 + s_lt_lo[dom * 4 + c], last = b0 + s_lt_hi[dom * 4 + c];
            int ntl = (last - first + ECDA_TR - 1) / ECDA_TR;
            if (nt_ + ntl > a.max_tiles) ntl = a.max_tiles - nt_;
            sg_on[sgi] = 1; sg_first[sgi] = first; sg_last[sgi] = last; sg_grow[sgi] = g_rows; sg_feas[sgi] = feas ? 1 : 0;
            sg_tile0[sgi] = nt_; sg_item0[sgi] = ng_; sg_pz0[sgi] = np_;
            nt_ += ntl;
            if (feas) {
                const int rb = (last - first + 127) / 128, cbs = (n_c + 255) / 256;
                ng_ += rb * cbs; np_ += rb;
                g_rows += rb * 128;
            }
        }
        if (ng_ > a.max_gram_items || np_ * ECDA_PZ_SPLITS > a.max_pz_items) { ng_ = 0; np_ = 0; }  // cannot happen (tables are sized for the worst case)
        int spl = np_ > 0 ? 148 / np_ : 1;  // split-K slices of (P Z'): one wave of items over the SMs
        spl = spl < 1 ? 1 : (spl > ECDA_PZ_SPLITS ? ECDA_PZ_SPLITS : spl);
        s_tot[0] = nt_; s_tot[1] = ng_; s_tot[2] = np_; s_tot[3] = spl;
        // k_ecda_p splits the 128 rows of a Gram item over 1, 2 or 4 CTAs: with few items (a data-parallel rank owns 1/world of the
        // rows) a single CTA per item leaves most SMs idle while each warp walks 16 rows back to back
        const int psp = ng_ <= 148 ? 4 : (ng_ <= 296 ? 2 : 1);
        st->n_tiles = nt_; st->n_gram_items = ng_; st->n_pz_items = np_ * spl; st->pz_splits = spl; st->p_split = psp;
    }
    __syncthreads();
    for (int sgi = 0; sgi < 8; ++sgi) {
        if (!sg_on[sgi]) continue;
        const int c = sgi >> 1;
        const int off = s_loff[c], n_c = s_cnt[c] + s_cnt[4 + c];
        const int first = sg_first[sgi], last = sg_last[sgi], grow = sg_grow[sgi], feas = sg_feas[sgi];
        const int ntl = min((last - first + ECDA_TR - 1) / ECDA_TR, a.max_tiles - sg_tile0[sgi]);
        for (int i = tid; i < ntl; i += EP_THREADS) {
            const int p = first + i * ECDA_TR;
            EcdaTile t;
            t.cls = c; t.row_begin = p; t.row_count = min(ECDA_TR, last - p); t.feasible = feas; t.g_row = grow + (p - first);
            t.pad_[0] = t.pad_[1] = t.pad_[2] = 0;
            a.tiles[sg_tile0[sgi] + i] = t;
        }
        if (feas && s_tot[1] > 0) {
            const int rb = (last - first + 127) / 128, cbs = (n_c + 255) / 256;
            for (int i = tid; i < rb * cbs; i += EP_THREADS) {
                const int r = i / cbs, cb = i - r * cbs;
                const int p = first + r * 128;
                GramItem gi;
                gi.a_row = p; gi.b_row = off + cb * 256; gi.g_row = grow + r * 128; gi.g_col = cb * 256;
                gi.n_rows = min(128, last - p); gi.n_cols = n_c; gi.cls = c; gi.pad_ = 0;
                a.gram_items[sg_item0[sgi] + i] = gi;
            }
            const int spl = s_tot[3], nkb = (n_c + 63) / 64, per = (nkb + spl - 1) / spl;
            for (int i = tid; i < rb * spl; i += EP_THREADS) {
                const int r = i / spl, sl = i - r * spl;
                PzItem pi;
                pi.g_row = grow + r * 128; pi.b_row = off; pi.kb0 = sl * per; pi.n_kb = max(0, min(per, nkb - sl * per));
                pi.slice = sl; pi.cls = c; pi.pad_[0] = pi.pad_[1] = 0;
                a.pz_items[(sg_pz0[sgi] + r) * spl + sl] = pi;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_ecda_p: elementwise pass over the tensor-core Gram G -- one CTA per Gram work item (128 owned rows x 256 members of the class),
// or per half / quarter of its rows when the rank has few items (EcdaStats::p_split, chosen by k_ecda_stats):
//   L2_ij = |z'_i|^2 + |z'_j|^2 - 2 G_ij (exactly 0 on the diagonal),  K = sum_r exp(-L2/beta_r),  K' = sum_r exp(-L2/beta_r)/beta_r
//   -> weighted block sums of K (the three MMD terms, I/utils.py:546-563) and P_ij = coef * K' (the gradient weights), written as a
//      per-class power-of-two scaled fp16 hi/lo split for the second tensor-core GEMM (k_pz_tc: sum_j P_ij z'_j), plus the row sums
//      of P per column block.  Warp = 16 (8, 4) of the rows, lane = 8 consecutive columns: 32-byte G loads, 16-byte P stores.
// Deterministic: per-item sums and per-(column block, row) partial row sums are folded in fixed order by k_ecda_fold / k_ecda_loss.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EP_THREADS) k_ecda_p(EcdaArgs a) {
    __shared__ float s_red[3][EP_THREADS / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const EcdaStats* st = a.stats;
    float sums[3] = {0.f, 0.f, 0.f};
    const int split = st->p_split, rpw = 16 / split;  // rows per warp
    const int n_units = st->n_gram_items * split;
    for (int w = blockIdx.x; w < n_units; w += gridDim.x) {
        const GramItem gi = a.gram_items[w / split];
        const int row0 = (w % split) * (128 / split) + warp * rpw;
        const int c = gi.cls;
        const float ib0 = st->inv_beta[c][0], ib1 = st->inv_beta[c][1], ib2 = st->inv_beta[c][2], ib3 = st->inv_beta[c][3], ib4 = st->inv_beta[c][4];
        const float css = st->c_ss[c], ctt = st->c_tt[c], cst = st->c_st[c];
        const bool chain = ib0 < 100.f;
        const float pcoef = -4.f * st->att[c] * a.w_ecda * st->p_scale[c];  // gradient weight scale x the fp16 operand scale of this class
        // this lane's 8 columns
        float njv[8], wjv[8];
        int dj[8];
        bool jok[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int col = lane * 8 + u;
            jok[u] = gi.g_col + col < gi.n_cols;
            const int pos = gi.b_row + col;
            njv[u] = jok[u] ? a.pos_nsq[pos] : 0.f;
            wjv[u] = jok[u] ? a.pos_om[pos] : 0.f;
            dj[u] = jok[u] ? a.pos_dom[pos] : 0;
        }
        sums[0] = sums[1] = sums[2] = 0.f;
#pragma unroll 4  // four rows' G loads in flight per warp (the pass is L2-latency bound)
        for (int rr = 0; rr < rpw; ++rr) {
            const int r = row0 + rr;
            const bool iok = r < gi.n_rows;
            const int pos_i = gi.a_row + r;
            const float nii = iok ? a.pos_nsq[pos_i] : 0.f, om_i = iok ? a.pos_om[pos_i] : 0.f;
            const int dom_i = iok ? a.pos_dom[pos_i] : 0;
            const int64_t base = (int64_t)(gi.g_row + r) * a.ldg + gi.g_col + lane * 8;
            float gv[8];
            if (iok) {
                const float4 g0 = __ldcg(reinterpret_cast<const float4*>(a.G + base)), g1 = __ldcg(reinterpret_cast<const float4*>(a.G + base + 4));
                gv[0] = g0.x; gv[1] = g0.y; gv[2] = g0.z; gv[3] = g0.w; gv[4] = g1.x; gv[5] = g1.y; gv[6] = g1.z; gv[7] = g1.w;
            }
            float pv[8], rowp = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float p = 0.f;
                if (iok && jok[u]) {
                    const bool diag = pos_i == gi.b_row + lane * 8 + u;
                    const float l2 = diag ? 0.f : fmaxf(nii + njv[u] - 2.f * gv[u], 0.f);
                    // beta_r = bw * 2^r (kernel_mul = 2, I/utils.py:541-543): exp(-L2/beta_r) = exp(-L2/beta_{r+1})^2 -- one exponential and
                    // four squarings instead of five exponentials (this pass is MUFU / issue bound).  In fp32 the squared chain is as
                    // accurate as the five direct evaluations, whose own error is dominated by the rounding of the argument
                    // (max relative error of e0 over L2/beta_0 <= 128: 6.0e-6 against 4.0e-6; of K: 8e-7 against 4e-7).  The reference
                    // adds 1e-8 to every beta_r, which breaks the exact ratio of two when the class bandwidth is tiny: the chain is used
                    // only where that term is below 1e-6 of beta_0 (bandwidth > 0.01, warp-uniform per work item), else five exponentials
                    float e0, e1, e2, e3;
                    const float e4 = expf(-l2 * ib4);
                    if (chain) { e3 = e4 * e4; e2 = e3 * e3; e1 = e2 * e2; e0 = e1 * e1; }
                    else { e0 = expf(-l2 * ib0); e1 = expf(-l2 * ib1); e2 = expf(-l2 * ib2); e3 = expf(-l2 * ib3); }
                    const float K = e0 + e1 + e2 + e3 + e4;
                    const float Kp = e0 * ib0 + e1 * ib1 + e2 * ib2 + e3 * ib3 + e4 * ib4;
                    const float ww = om_i * wjv[u];
                    float sblk;
                    if (dom_i == 0 && dj[u] == 0) { sums[0] += ww * K; sblk = css; }
                    else if (dom_i == 1 && dj[u] == 1) { sums[1] += ww * K; sblk = ctt; }
                    else { if (dom_i == 0) sums[2] += ww * K; sblk = -cst; }
                    p = diag ? 0.f : pcoef * sblk * ww * Kp;  // (z_i - z_i) = 0: the diagonal carries no gradient
                }
                pv[u] = p;
                rowp += p;
            }
            uint32_t h[4], l[4];
            tc::split2_f16(pv[0], pv[1], h[0], l[0]); tc::split2_f16(pv[2], pv[3], h[1], l[1]);
            tc::split2_f16(pv[4], pv[5], h[2], l[2]); tc::split2_f16(pv[6], pv[7], h[3], l[3]);
            *reinterpret_cast<uint4*>(a.p_hi + base) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(a.p_lo + base) = make_uint4(l[0], l[1], l[2], l[3]);
            rowp = warp_sum(rowp);
            if (lane == 0) a.rowp_part[(int64_t)(gi.g_col >> 8) * a.g_rows_alloc + gi.g_row + r] = rowp;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = warp_sum(sums[k]);
            if (lane == 0) s_red[k][warp] = v;
        }
        __syncthreads();
        if (tid < 3) {
            float v = 0.f;
            for (int ww = 0; ww < EP_THREADS / 32; ++ww) v += s_red[tid][ww];
            a.item_sums[w * 3 + tid] = v;
        }
        __syncthreads();
    }
}

// k_ecda_loss (one CTA): per-class block sums over this rank's Gram items in fixed order -> this rank's share of the ECDA loss
__global__ void __launch_bounds__(EP_THREADS) k_ecda_loss(EcdaArgs a) {
    __shared__ float s_cs[ECDA_MAXC * 3][EP_THREADS / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const EcdaStats* st = a.stats;
    const int split = st->p_split, n_items = st->n_gram_items * split;  // one entry per k_ecda_p work unit
    float cs[ECDA_MAXC][3];
#pragma unroll
    for (int c = 0; c < ECDA_MAXC; ++c) cs[c][0] = cs[c][1] = cs[c][2] = 0.f;
    for (int e = tid; e < n_items; e += EP_THREADS) {
        const int c = a.gram_items[e / split].cls;
        const float v0 = a.item_sums[e * 3 + 0], v1 = a.item_sums[e * 3 + 1], v2 = a.item_sums[e * 3 + 2];
#pragma unroll
        for (int cc = 0; cc < ECDA_MAXC; ++cc)
            if (cc == c) { cs[cc][0] += v0; cs[cc][1] += v1; cs[cc][2] += v2; }
    }
#pragma unroll
    for (int c = 0; c < ECDA_MAXC; ++c)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = warp_sum(cs[c][k]);
            if (lane == 0) s_cs[c * 3 + k][warp] = v;
        }
    __syncthreads();
    if (tid != 0) return;
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

// k_ecda_fold: one warp per owned member row: d/dz_i = rowP_i z'_i - (P Z')_i  (= sum_j P_ij (z_i - z_j), I/utils.py:537 differentiated;
// (P Z') from k_pz_tc, rowP folded over the column blocks in fixed order) + the compactness and repulsion gradients of the noisy
// rows, accumulated into g.  Every gradient row has exactly one writer: no float atomics, bit-reproducible.
__global__ void __launch_bounds__(EP_THREADS) k_ecda_fold(EcdaArgs a) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const EcdaStats* st = a.stats;
    const int tile_id = blockIdx.x;
    if (tile_id >= st->n_tiles) return;
    const int Bg = a.pack.B_local * a.pack.world;
    const EcdaTile t = a.tiles[tile_id];
    const int c = t.cls;
    const int dom_i = (t.row_begin - st->list_off[c]) >= st->n_s[c] ? 1 : 0;
    const float comp_k = (t.feasible && dom_i) ? a.w_ecda * st->att[c] * a.gamma_comp * 2.f / (float)st->n_t[c] : 0.f;
    const int row = blockIdx.y * 8 + warp;
    if (row >= t.row_count) return;
    const int rid = a.lists[t.row_begin + row];
    float g8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float rp = 0.f;
    if (t.feasible) {
        const float unscale = 1.0f / st->p_scale[c];
        const int n_cb = (st->n_s[c] + st->n_t[c] + 255) >> 8;
        for (int cb = 0; cb < n_cb; ++cb) rp += a.rowp_part[(int64_t)cb * a.g_rows_alloc + t.g_row + row];
        rp *= unscale;
        for (int sl = 0; sl < st->pz_splits; ++sl) {  // split-K slices in fixed order
            const float* pa = a.accz + ((int64_t)sl * a.g_rows_alloc + t.g_row + row) * DAD_H;
            const float4 u0 = *reinterpret_cast<const float4*>(pa + lane * 4), u1 = *reinterpret_cast<const float4*>(pa + 128 + lane * 4);
            g8[0] += u0.x; g8[1] += u0.y; g8[2] += u0.z; g8[3] += u0.w; g8[4] += u1.x; g8[5] += u1.y; g8[6] += u1.z; g8[7] += u1.w;
        }
#pragma unroll
        for (int f = 0; f < 8; ++f) g8[f] *= unscale;
    }
    const int gi = rid < Bg ? rid : rid - Bg;
    const float* z = rid < Bg ? pack_emb_clean(a.pack, rid) : pack_emb_strong(a.pack, rid - Bg);
    float* g = (rid < Bg ? a.g_clean : a.g_strong) + (int64_t)(gi - a.rank * a.pack.B_local) * DAD_H;
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        const int h = lane * 4 + (f & 3) + 128 * (f >> 2);
        const float zr = z[h];
        float v = rp * (zr - a.mu_all[c * DAD_H + h]) - g8[f];
        if (dom_i && a.class_aware) v += comp_k * (zr - a.mu_t[c * DAD_H + h]) + a.w_ecda * a.rep_grad[c * DAD_H + h];
        g[h] += v;
    }
}

int launch_ecda_gram(const EcdaArgs& a, cudaStream_t st);  // k_gram_tc.cu: pack + Gram GEMM
int launch_ecda_pz(const EcdaArgs& a, cudaStream_t st);    // k_gram_tc.cu: (P Z') GEMM
int ecda_max_gram_items(int B_local, int world) { return (ceil_div(2 * B_local, 128) + 8) * (ceil_div(2 * B_local * world, 256) + 1); }
int ecda_max_pz_items(int B_local) { return (ceil_div(2 * B_local, 128) + 8) * ECDA_PZ_SPLITS; }
int ecda_prep_grid(int B_global) {
    // rows per CTA <= 256 (k_ecda_prep classifies its chunk with one thread per row); aim at ~16 rows per CTA, at most one CTA
    // per SM (every partial is one more term of the fixed-order fold in k_ecda_reduce)
    int g = ceil_div(2 * B_global, 16);
    if (g > 148) g = 148;
    const int gmin = ceil_div(2 * B_global, EP_THREADS);
    if (g < gmin) g = gmin;
    return g < 1 ? 1 : g;
}
int ecda_max_tiles(int B_local) { return ceil_div(2 * B_local, ECDA_TR) + 2 * ECDA_MAXC; }

int launch_ecda(const EcdaArgs& a, cudaStream_t st) {
    const size_t stats_smem = ((size_t)2 * a.pack.B_local * a.pack.world + 15) & ~(size_t)15;  // row slots, one byte each
    if (stats_smem > 160 * 1024) return dad_set_error(DAD_ERR_UNSUPPORTED, "ecda: global batch too large for the single-CTA list builder");
    if (stats_smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_ecda_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "ecda: cannot opt in to shared memory: %s", cudaGetErrorString(e));
    }
    if (a.phases & 1) {
        k_ecda_prep<<<a.prep_grid, EP_THREADS, 0, st>>>(a);
        DAD_CHECK_LAUNCH("k_ecda_prep");
        k_ecda_reduce<<<32, EP_THREADS, 0, st>>>(a, a.prep_grid);
        DAD_CHECK_LAUNCH("k_ecda_reduce");
        k_ecda_stats<<<1, EP_THREADS, stats_smem, st>>>(a);
        DAD_CHECK_LAUNCH("k_ecda_stats");
        int rc = launch_ecda_gram(a, st);  // z' -> fp16 hi/lo, G = Z' Z'^T on tcgen05
        if (rc) return rc;
        const int pg = a.max_gram_items * 4 < 4 * 148 ? a.max_gram_items * 4 : 4 * 148;  // work units: up to 4 per Gram item
        k_ecda_p<<<pg < 1 ? 1 : pg, EP_THREADS, 0, st>>>(a);
        DAD_CHECK_LAUNCH("k_ecda_p");
        rc = launch_ecda_pz(a, st);        // (P Z') on tcgen05
        if (rc) return rc;
    }
    if (a.phases & 2) {
        k_ecda_fold<<<dim3(a.max_tiles, ECDA_TR / 8), EP_THREADS, 0, st>>>(a);
        DAD_CHECK_LAUNCH("k_ecda_fold");
        k_ecda_loss<<<1, EP_THREADS, 0, st>>>(a);
        DAD_CHECK_LAUNCH("k_ecda_loss");
    }
    return DAD_OK;
}
