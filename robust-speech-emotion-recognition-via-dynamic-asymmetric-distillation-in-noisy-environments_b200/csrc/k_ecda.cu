// ECDA: class-aware, attention-weighted 5-kernel Gaussian MMD + compactness + centroid repulsion, forward and
// analytic backward to d(loss)/d(embedding), in fixed-shape masked form.
// Reference: I/utils.py:510-652 (ECDALoss._gaussian_kernel / forward), gate I/train.py:444,450.
//
//  k_ecda_prep : grid over the 2*B_global rows (clean rows then noisy rows of the pack) -> per-(domain,class)
//                partial sums;  k_ecda_reduce (8 CTAs, one per slot) folds them in fixed order (fp64);  k_ecda_stats
//                (one CTA) turns them into class statistics (counts, bandwidths via the centred closed form
//                sum_ij|zi-zj|^2 = 2n sum_i|zi-m|^2, centroids, compactness, repulsion + its gradient, attention) and
//                builds the per-class member lists and the row-tile table.
//  k_ecda_pairs: one CTA per (64-row tile of one class, column split): two register-tiled SIMT GEMMs on class-centred
//                embeddings (Gram form of the squared distances, then P z) -- see the comment above the kernel.  Under data
//                parallelism a rank launches tiles only for the rows it owns (its local clean / noisy samples).
#include "dad_common.cuh"
#include "dad_pack.cuh"

namespace {

constexpr int EP_THREADS = 256;

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

}  // namespace

__global__ void __launch_bounds__(EP_THREADS) k_ecda_prep(EcdaArgs a) {
    __shared__ double s_red[EP_THREADS / 32];
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
            a.slot_of[r0 + tid] = (uint8_t)(s_slot[tid] < 0 ? 255 : s_slot[tid]);  // for the ordered scatter of phase 3
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
    for (int u = 0; g < g1; ++g, ++u) t8[u] += (double)a.part_vec[((int64_t)g * 8 + s) * DAD_H + f];
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
    __shared__ double s_red[EP_THREADS / 32];
    __shared__ int s_cnt[8], s_off[8], s_wc[EP_THREADS / 32][8];
    __shared__ double s_sq[8], s_ws[8];
    __shared__ float s_dist[ECDA_MAXC][ECDA_MAXC];
    const int tid = threadIdx.x;
    const int Bg = a.pack.B_local * a.pack.world;
    const int R = 2 * Bg;
    const bool active = a.nmask[0] > 1;

    // ---------------- phase 2: class statistics
    EcdaStats* st = a.stats;
    if (tid == 0) st->active = active ? 1 : 0;
    if (!active) {
        if (tid == 0) { st->n_tiles = 0; st->n_gram_items = 0; st->any_feasible = 0; st->rep = 0.f; st->rho = 0.f; }
        return;
    }
    double tv[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) tv[s] = a.red_vec[s * DAD_H + tid];
    if (tid < 8) { s_sq[tid] = a.red_sc[tid * 2 + 0]; s_ws[tid] = a.red_sc[tid * 2 + 1]; s_cnt[tid] = a.red_cnt[tid]; }
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
        a.mu_all[c * DAD_H + tid] = (float)m;
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
    // All row slots were published by phase 1; one coalesced pass brings them into shared memory (every load in flight at
    // once), so the ordered scatter below touches no global memory on its critical path.  (The first version re-derived each
    // slot from the pack inside the loop -- 32 dependent global round trips -- and then binary-searched the lists it had just
    // written with L2-latency probes from one thread: together ~0.15 ms of a single CTA at global batch 4096.)
    extern __shared__ uint8_t s_sl[];  // [R]
    __shared__ int s_lt_lo[8], s_lt_hi[8];
    if (tid < 8) {
        const int c = tid & 3, dom = tid >> 2;
        int off = 0;
        for (int cc = 0; cc < c; ++cc) off += s_cnt[cc] + s_cnt[4 + cc];
        s_off[tid] = off + (dom ? s_cnt[c] : 0);
        s_lt_lo[tid] = 0; s_lt_hi[tid] = 0;
    }
    for (int r = tid; r < R; r += EP_THREADS) s_sl[r] = __ldcg(a.slot_of + r);
    __syncthreads();
    const int own_lo = a.rank * a.pack.B_local, own_hi = own_lo + a.pack.B_local;
    const int warp = tid >> 5, lane = tid & 31;
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
    // ---------------- phase 4: row tiles for the rows this rank owns
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        int nt_ = 0, ng_ = 0, g_rows = 0;
        for (int c = 0; c < NC; ++c) {
            const int off = st->list_off[c], ns = s_cnt[c], n_c = s_cnt[c] + s_cnt[4 + c];
            const bool feas = st->feasible[c] != 0;
            for (int dom = 0; dom < 2; ++dom) {
                if (dom == 0 && !feas) continue;                              // clean rows only matter through the MMD
                if (dom == 1 && !feas && !(a.class_aware && nvalid > 1)) continue;  // noisy rows: MMD/comp or repulsion
                const int b0 = off + (dom ? ns : 0);
                const int first = b0 + s_lt_lo[dom * 4 + c], last = b0 + s_lt_hi[dom * 4 + c];
                for (int p = first; p < last && nt_ < a.max_tiles; p += ECDA_TR) {
                    EcdaTile t;
                    t.cls = c; t.row_begin = p; t.row_count = min(ECDA_TR, last - p); t.feasible = feas ? 1 : 0;
                    t.g_row = g_rows + (p - first);
                    a.tiles[nt_++] = t;
                }
                if (feas) {  // Gram work items: 128-row blocks of this segment x 256-column blocks of the class (k_gram_tc)
                    for (int p = first; p < last; p += 128)
                        for (int cb = 0; cb < n_c && ng_ < a.max_gram_items; cb += 256) {
                            GramItem gi;
                            gi.a_row = p; gi.b_row = off + cb; gi.g_row = g_rows + (p - first); gi.g_col = cb;
                            a.gram_items[ng_++] = gi;
                        }
                    g_rows += ((last - first + 127) / 128) * 128;
                }
            }
        }
        st->n_tiles = nt_;
        st->n_gram_items = ng_;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_ecda_pairs: one CTA per (64-row tile of one class, column split q).  The class's members are processed 64 columns at a
// time as two register-tiled SIMT GEMMs on class-CENTRED embeddings z' = z - mean_c (centring removes the cancellation of the
// Gram form: |z'| is of the order of the pair distances themselves):
//   phase A  dot_ij = z'_i . z'_j  read from the tcgen05 Gram G (k_gram_tc.cu; thread = 4x4 outputs),
//            L2_ij = |z'_i|^2 + |z'_j|^2 - 2 dot_ij  (exactly 0 on the diagonal),
//            K = sum_r exp(-L2/beta_r), K' = sum_r exp(-L2/beta_r)/beta_r -> weighted block sums (loss) and P_ij = coef K'
//   phase B  accZ_i += sum_j P_ij z'_j,  rowP_i += sum_j P_ij            (thread = 8 rows x 8 features)
// so that d/dz_i = rowP_i z'_i - accZ_i  (= sum_j P_ij (z_i - z_j), I/utils.py:537 differentiated).  Each CTA writes its partial
// (accZ, rowP, sums) to the workspace; k_ecda_fold then folds the Q partials of a row tile in fixed q order, adds the
// compactness / repulsion gradients and accumulates into g -- every gradient row has exactly one writer, no float atomics,
// bit-reproducible.  The column split gives ~2 CTAs per SM on one GPU and keeps all SMs busy when a rank owns only 1/8 of the rows.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int PT = ECDA_TR;          // 64: rows per tile = columns per chunk
constexpr int PZS = DAD_H + 4;       // padded row stride of the staged embeddings (floats): conflict-free float4 rows
constexpr int PPS = PT + 4;          // padded row stride of P
}

__global__ void __launch_bounds__(EP_THREADS, 1) k_ecda_pairs(EcdaArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* zi = smem;                       // [64][PZS]  centred row tile
    float* zj = zi + PT * PZS;              // [64][PZS]  centred column chunk
    float* Pm = zj + PT * PZS;              // [64][PPS]
    float* mu = Pm + PT * PPS;              // [256] class mean
    float* ni = mu + DAD_H;                 // [64] |z'_i|^2
    float* nj = ni + PT;                    // [64]
    float* wi = nj + PT;                    // [64] omega_i
    float* wj = wi + PT;                    // [64]
    int* ri = reinterpret_cast<int*>(wj + PT);  // [64] row ids (-1 = padding)
    int* pj = ri + PT;                      // [64] list position of column j (-1 = padding)
    int* dj = pj + PT;                      // [64] column domain (0 clean, 1 noisy)
    __shared__ float s_red[3][EP_THREADS / 32];
    __shared__ int s_flag;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const EcdaStats* st = a.stats;
    const int Bg = a.pack.B_local * a.pack.world;
    const int n_tiles = st->n_tiles;
    const int Q = a.col_splits;
    const int tile_id = blockIdx.x / Q, q = blockIdx.x - tile_id * Q;
    float sums[3] = {0.f, 0.f, 0.f};

    if (tile_id < n_tiles) {
        const EcdaTile t = a.tiles[tile_id];
        const int c = t.cls;
        const int off = st->list_off[c], ns = st->n_s[c], n = ns + st->n_t[c];
        const int dom_i = (t.row_begin - off) >= ns ? 1 : 0;
        float* part_acc = a.part_acc + (int64_t)blockIdx.x * PT * DAD_H;
        float* part_rowp = a.part_rowp + (int64_t)blockIdx.x * PT;
        // ---- stage the class mean and the (centred) row tile
        mu[tid] = a.mu_all[c * DAD_H + tid];
        if (tid < PT) {
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
        auto stage = [&](float* dst, float* nsq, const int* rows_or_null, int list_base, int count) {
            // 8 warps x 8 rows; lane owns features lane*4..+3 and 128+lane*4..+3
#pragma unroll 2
            for (int rr = 0; rr < 8; ++rr) {
                const int row = warp * 8 + rr;
                int r = -1;
                if (rows_or_null) r = rows_or_null[row];
                else if (row < count) r = a.lists[list_base + row];
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                if (r >= 0) {
                    const float* z = r < Bg ? pack_emb_clean(a.pack, r) : pack_emb_strong(a.pack, r - Bg);
                    v0 = *reinterpret_cast<const float4*>(z + lane * 4);
                    v1 = *reinterpret_cast<const float4*>(z + 128 + lane * 4);
                    const float4 m0 = *reinterpret_cast<const float4*>(mu + lane * 4), m1 = *reinterpret_cast<const float4*>(mu + 128 + lane * 4);
                    v0.x -= m0.x; v0.y -= m0.y; v0.z -= m0.z; v0.w -= m0.w;
                    v1.x -= m1.x; v1.y -= m1.y; v1.z -= m1.z; v1.w -= m1.w;
                }
                *reinterpret_cast<float4*>(dst + row * PZS + lane * 4) = v0;
                *reinterpret_cast<float4*>(dst + row * PZS + 128 + lane * 4) = v1;
                float s2 = v0.x * v0.x;
                s2 = fmaf(v0.y, v0.y, s2); s2 = fmaf(v0.z, v0.z, s2); s2 = fmaf(v0.w, v0.w, s2);
                s2 = fmaf(v1.x, v1.x, s2); s2 = fmaf(v1.y, v1.y, s2); s2 = fmaf(v1.z, v1.z, s2); s2 = fmaf(v1.w, v1.w, s2);
                s2 = warp_sum(s2);
                if (lane == 0) nsq[row] = s2;
            }
        };
        stage(zi, ni, ri, 0, 0);
        __syncthreads();

        // phase-B accumulators: rows warp*8 + r, features lane*4 + (f&3) + 128*(f>>2)
        float acc[8][8], rowp[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            rowp[r] = 0.f;
#pragma unroll
            for (int f = 0; f < 8; ++f) acc[r][f] = 0.f;
        }
        if (t.feasible) {
            const float ib0 = st->inv_beta[c][0], ib1 = st->inv_beta[c][1], ib2 = st->inv_beta[c][2], ib3 = st->inv_beta[c][3], ib4 = st->inv_beta[c][4];
            const float css = st->c_ss[c], ctt = st->c_tt[c], cst = st->c_st[c];
            const float gscale = -4.f * st->att[c] * a.w_ecda;
            const int ty = tid >> 4, tx = tid & 15;  // phase A: rows ty + 16a, columns tx + 16b
            const int n_chunks = ceil_div(n, PT), per = ceil_div(n_chunks, Q);
            const int ch0 = q * per, ch1 = min(n_chunks, ch0 + per);
            for (int ch = ch0; ch < ch1; ++ch) {
                const int j0 = ch * PT;
                __syncthreads();  // previous chunk's readers of zj / Pm are done
                if (tid < PT) {
                    int pos = -1, d = 0;
                    float om = 0.f;
                    if (j0 + tid < n) {
                        pos = off + j0 + tid;
                        const int r = a.lists[pos];
                        d = (r >= Bg) ? 1 : 0;
                        om = (d && a.class_aware) ? pack_score(a.pack, r - Bg) : 1.f;
                    }
                    pj[tid] = pos; dj[tid] = d; wj[tid] = om;
                }
                stage(zj, nj, nullptr, off + j0, n - j0);
                __syncthreads();
                // ---- phase A: the 4x4 dot products of this thread come from the tensor-core Gram (k_gram_tc; L2-resident)
                float d[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const float* grow = a.G + (int64_t)(t.g_row + ty + 16 * x) * a.ldg + j0 + tx;
#pragma unroll
                    for (int y = 0; y < 4; ++y) d[x][y] = __ldcg(grow + 16 * y);
                }
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int i = ty + 16 * x;
                    const bool i_ok = i < t.row_count;
                    const float om_i = wi[i], nii = ni[i];
                    const int pos_i = t.row_begin + i;
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const int j = tx + 16 * y;
                        const int pos = pj[j];
                        float p = 0.f;
                        if (i_ok && pos >= 0) {
                            const bool diag = pos == pos_i;
                            const float l2 = diag ? 0.f : fmaxf(nii + nj[j] - 2.f * d[x][y], 0.f);
                            const float e0 = expf(-l2 * ib0), e1 = expf(-l2 * ib1), e2 = expf(-l2 * ib2), e3 = expf(-l2 * ib3), e4 = expf(-l2 * ib4);
                            const float K = e0 + e1 + e2 + e3 + e4;
                            const float Kp = e0 * ib0 + e1 * ib1 + e2 * ib2 + e3 * ib3 + e4 * ib4;
                            const float ww = om_i * wj[j];
                            const int dd = dj[j];
                            float sblk;
                            if (dom_i == 0 && dd == 0) { sums[0] += ww * K; sblk = css; }
                            else if (dom_i == 1 && dd == 1) { sums[1] += ww * K; sblk = ctt; }
                            else { if (dom_i == 0) sums[2] += ww * K; sblk = -cst; }
                            p = diag ? 0.f : gscale * sblk * ww * Kp;  // (z_i - z_i) = 0: the diagonal carries no gradient
                        }
                        Pm[i * PPS + j] = p;
                    }
                }
                __syncthreads();
                // ---- phase B: accZ_i += sum_j P_ij z'_j, rowP_i += sum_j P_ij   (rows warp*8.., 8 features per lane)
#pragma unroll 1
                for (int j = 0; j < PT; j += 4) {
                    float4 pr[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) pr[r] = *reinterpret_cast<const float4*>(Pm + (warp * 8 + r) * PPS + j);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float4 z0 = *reinterpret_cast<const float4*>(zj + (j + jj) * PZS + lane * 4);
                        const float4 z1 = *reinterpret_cast<const float4*>(zj + (j + jj) * PZS + 128 + lane * 4);
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const float p = jj == 0 ? pr[r].x : jj == 1 ? pr[r].y : jj == 2 ? pr[r].z : pr[r].w;
                            rowp[r] += p;
                            acc[r][0] = fmaf(p, z0.x, acc[r][0]); acc[r][1] = fmaf(p, z0.y, acc[r][1]);
                            acc[r][2] = fmaf(p, z0.z, acc[r][2]); acc[r][3] = fmaf(p, z0.w, acc[r][3]);
                            acc[r][4] = fmaf(p, z1.x, acc[r][4]); acc[r][5] = fmaf(p, z1.y, acc[r][5]);
                            acc[r][6] = fmaf(p, z1.z, acc[r][6]); acc[r][7] = fmaf(p, z1.w, acc[r][7]);
                        }
                    }
                }
            }
        }
        // ---- this CTA's partial -> workspace
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int row = warp * 8 + r;
            *reinterpret_cast<float4*>(part_acc + row * DAD_H + lane * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            *reinterpret_cast<float4*>(part_acc + row * DAD_H + 128 + lane * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
            if (lane == 0) part_rowp[row] = rowp[r];
        }
    }
    // ---- per-CTA loss partial sums -> fixed-order loss assembly by the last CTA of the grid
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float v = warp_sum(sums[k]);
        if (lane == 0) s_red[k][warp] = v;
    }
    __syncthreads();
    if (tid < 3) {
        float v = 0.f;
        for (int w = 0; w < EP_THREADS / 32; ++w) v += s_red[tid][w];
        a.tile_sums[blockIdx.x * 3 + tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&a.counter[1], 1u);
        s_flag = (prev == gridDim.x - 1u);
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    float cs[ECDA_MAXC][3];
#pragma unroll
    for (int c = 0; c < ECDA_MAXC; ++c) cs[c][0] = cs[c][1] = cs[c][2] = 0.f;
    for (int e = tid; e < n_tiles * Q; e += EP_THREADS) {
        const int tl = e / Q;
        const int c = a.tiles[tl].cls;
        if (!a.tiles[tl].feasible) continue;
        const float v0 = __ldcg(&a.tile_sums[e * 3 + 0]), v1 = __ldcg(&a.tile_sums[e * 3 + 1]), v2 = __ldcg(&a.tile_sums[e * 3 + 2]);
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

// k_ecda_fold: one CTA per row tile: folds the Q column-split partials of k_ecda_pairs in fixed q order, adds the compactness
// and repulsion gradients of the noisy rows and accumulates into g (each gradient row has exactly one writer: deterministic).
__global__ void __launch_bounds__(EP_THREADS) k_ecda_fold(EcdaArgs a) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const EcdaStats* st = a.stats;
    const int tile_id = blockIdx.x;  // blockIdx.y = which 8 rows of the tile (one row per warp: Q dependent loads per row,
    if (tile_id >= st->n_tiles) return;  // so the fold is latency-bound and wants many independent rows in flight)
    const int Bg = a.pack.B_local * a.pack.world, Q = a.col_splits;
    const EcdaTile t = a.tiles[tile_id];
    const int c = t.cls;
    const int dom_i = (t.row_begin - st->list_off[c]) >= st->n_s[c] ? 1 : 0;
    const float comp_k = (t.feasible && dom_i) ? a.w_ecda * st->att[c] * a.gamma_comp * 2.f / (float)st->n_t[c] : 0.f;
    {
        const int row = blockIdx.y * 8 + warp;
        if (row >= t.row_count) return;
        const int rid = a.lists[t.row_begin + row];
        float g8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float rp = 0.f;
        if (t.feasible) {
            for (int qq = 0; qq < Q; ++qq) {
                const float* pa = a.part_acc + ((int64_t)(tile_id * Q + qq) * PT + row) * DAD_H;
                const float4 u0 = *reinterpret_cast<const float4*>(pa + lane * 4), u1 = *reinterpret_cast<const float4*>(pa + 128 + lane * 4);
                g8[0] += u0.x; g8[1] += u0.y; g8[2] += u0.z; g8[3] += u0.w; g8[4] += u1.x; g8[5] += u1.y; g8[6] += u1.z; g8[7] += u1.w;
                rp += a.part_rowp[(int64_t)(tile_id * Q + qq) * PT + row];
            }
        }
        const int gi = rid < Bg ? rid : rid - Bg;
        const float* z = rid < Bg ? pack_emb_clean(a.pack, rid) : pack_emb_strong(a.pack, rid - Bg);
        float* g = (rid < Bg ? a.g_clean : a.g_strong) + (int64_t)(gi - a.rank * a.pack.B_local) * DAD_H;
#pragma unroll
        for (int f = 0; f < 8; ++f) {
            const int h = lane * 4 + (f & 3) + 128 * (f >> 2);
            const float zr = z[h];
            float v = rp * (zr - a.mu_all[c * DAD_H + h]) - g8[f];  // sum_j P_ij (z_i - z_j) on the centred rows
            if (dom_i && a.class_aware) v += comp_k * (zr - a.mu_t[c * DAD_H + h]) + a.w_ecda * a.rep_grad[c * DAD_H + h];
            g[h] += v;
        }
    }
}

int launch_ecda_gram(const EcdaArgs& a, cudaStream_t st);  // k_gram_tc.cu
int ecda_max_gram_items(int B_local, int world) { return (ceil_div(2 * B_local, 128) + 8) * (ceil_div(2 * B_local * world, 256) + 1); }
size_t ecda_pairs_smem_bytes() { return (size_t)(2 * PT * PZS + PT * PPS + DAD_H + 4 * PT) * sizeof(float) + 3 * PT * sizeof(int); }
int ecda_prep_grid(int B_global) {
    // rows per CTA <= 256 (k_ecda_prep classifies its chunk with one thread per row); aim at ~16 rows per CTA
    int g = ceil_div(2 * B_global, 16);
    if (g > 296) g = 296;
    const int gmin = ceil_div(2 * B_global, EP_THREADS);
    if (g < gmin) g = gmin;
    return g < 1 ? 1 : g;
}
int ecda_max_tiles(int B_local) { return ceil_div(2 * B_local, ECDA_TR) + 2 * ECDA_MAXC; }
// column splits per row tile: the kernel runs one CTA per SM (152 KB of shared memory), so the grid should be ONE wave:
// Q = floor(148 / row tiles this rank can own), at most 16.  (Q = 13 over 24 tiles = 2.1 waves cost 2x at 1/8 of the rows.)
int ecda_col_splits(int B_local) {
    const int tiles = ceil_div(2 * B_local, ECDA_TR);
    int q = 148 / (tiles < 1 ? 1 : tiles);
    return q < 1 ? 1 : (q > 16 ? 16 : q);
}

int launch_ecda(const EcdaArgs& a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_ecda_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ecda_pairs_smem_bytes());
        attr_set = true;
    }
    const size_t stats_smem = (size_t)2 * a.pack.B_local * a.pack.world;  // row slots, one byte each
    if (stats_smem > 160 * 1024) return dad_set_error(DAD_ERR_UNSUPPORTED, "ecda: global batch too large for the single-CTA list builder");
    if (stats_smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_ecda_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "ecda: cannot opt in to shared memory: %s", cudaGetErrorString(e));
    }
    k_ecda_prep<<<a.prep_grid, EP_THREADS, 0, st>>>(a);
    DAD_CHECK_LAUNCH("k_ecda_prep");
    k_ecda_reduce<<<32, EP_THREADS, 0, st>>>(a, a.prep_grid);
    DAD_CHECK_LAUNCH("k_ecda_reduce");
    k_ecda_stats<<<1, EP_THREADS, stats_smem, st>>>(a);
    DAD_CHECK_LAUNCH("k_ecda_stats");
    int rc = launch_ecda_gram(a, st);
    if (rc) return rc;
    k_ecda_pairs<<<a.max_tiles * a.col_splits, EP_THREADS, ecda_pairs_smem_bytes(), st>>>(a);
    DAD_CHECK_LAUNCH("k_ecda_pairs");
    k_ecda_fold<<<dim3(a.max_tiles, ECDA_TR / 8), EP_THREADS, 0, st>>>(a);
    DAD_CHECK_LAUNCH("k_ecda_fold");
    return DAD_OK;
}
