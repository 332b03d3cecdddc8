// The per-step exchange buffer ("pack") and kernel argument blocks shared between the .cu files.
//
// Pack layout, one block per rank, ranks concatenated (this is exactly what the single NCCL all-gather of the
// data-parallel step moves; world == 1 on one GPU).  Per rank, in 4-byte words:
//   [ emb_clean : B_local*256 f32 ][ emb_strong : B_local*256 f32 ][ score : B_local f32 ][ pred : B_local i32 ][ label : B_local i32 ]
// Global sample index gi = rank*B_local + i.
#pragma once
#include "dad_common.cuh"

typedef dad_pack Pack;
__host__ __device__ __forceinline__ int64_t pack_rank_words(int B_local) { return (int64_t)B_local * (2 * DAD_H + 3); }
__device__ __forceinline__ float* pack_rank_base(const Pack& p, int gi, int& i) {
    const int r = gi / p.B_local;
    i = gi - r * p.B_local;
    return p.base + (int64_t)r * pack_rank_words(p.B_local);
}
__device__ __forceinline__ float* pack_emb_clean(const Pack& p, int gi) { int i; float* b = pack_rank_base(p, gi, i); return b + (int64_t)i * DAD_H; }
__device__ __forceinline__ float* pack_emb_strong(const Pack& p, int gi) { int i; float* b = pack_rank_base(p, gi, i); return b + (int64_t)(p.B_local + i) * DAD_H; }
__device__ __forceinline__ float& pack_score(const Pack& p, int gi) { int i; float* b = pack_rank_base(p, gi, i); return b[(int64_t)2 * p.B_local * DAD_H + i]; }
__device__ __forceinline__ int& pack_pred(const Pack& p, int gi) { int i; float* b = pack_rank_base(p, gi, i); return reinterpret_cast<int*>(b)[(int64_t)2 * p.B_local * DAD_H + p.B_local + i]; }
__device__ __forceinline__ int& pack_label(const Pack& p, int gi) { int i; float* b = pack_rank_base(p, gi, i); return reinterpret_cast<int*>(b)[(int64_t)2 * p.B_local * DAD_H + 2 * p.B_local + i]; }

typedef dad_head_fwd_args HeadFwdArgs;
typedef dad_dacp_args DacpArgs;
typedef dad_head_bwd_args HeadBwdArgs;

// ---- ECDA
constexpr int ECDA_TR = 64;  // rows (members of one class) per CTA tile
constexpr int ECDA_MAXC = 4;
constexpr float ECDA_ZS = 256.0f;  // power-of-two scale of the centred embeddings before the fp16 hi/lo split (k_gram_tc.cu)
struct EcdaTile { int cls, row_begin, row_count, feasible, g_row, pad_[3]; };  // g_row: first row of this tile in the Gram buffer
// 128 list rows from a_row (n_rows valid) x 256 list rows from b_row (columns g_col.. of the class's n_cols members) -> G[g_row.., g_col..]
struct GramItem { int a_row, b_row, g_row, g_col, n_rows, n_cols, cls, pad_; };
// (P Z') block: P rows g_row..+127, Z' rows from b_row; this item covers k-blocks [kb0, kb0 + n_kb) of 64 members (split-K slice `slice`)
struct PzItem { int g_row, b_row, kb0, n_kb, slice, cls, pad_[2]; };
constexpr int ECDA_PZ_SPLITS = 16;  // max split-K slices of the (P Z') GEMM: k_ecda_stats picks 148 / (row blocks this rank owns), so a data-parallel rank that owns 4-12 blocks still fills the SMs (r01: 16 CTAs, 46 us at 512 samples per rank)

struct EcdaStats {  // written by k_ecda_stats, read by the Gram / P / fold / loss kernels
    int n_s[ECDA_MAXC], n_t[ECDA_MAXC], feasible[ECDA_MAXC], list_off[ECDA_MAXC];
    float inv_beta[ECDA_MAXC][5];                             // 1/(bw*2^r + 1e-8)
    float c_ss[ECDA_MAXC], c_tt[ECDA_MAXC], c_st[ECDA_MAXC];  // 1/(D_ss+1e-8), 1/(D_tt+1e-8), 1/(D_st+1e-8)
    float att[ECDA_MAXC];                                     // class attention a_c
    float comp[ECDA_MAXC];                                    // compactness per class
    float rep;                                                // repulsion loss (<= 0)
    float rho;                                                // delta * sum_{feasible} a_c
    int n_tiles, any_feasible, n_classes, active;
    int n_gram_items, n_pz_items, pz_splits, p_split;       // p_split: CTAs per Gram item in k_ecda_p (1, 2 or 4)
    float p_scale[ECDA_MAXC];                                 // power-of-two scale of P before its fp16 split
};

struct EcdaArgs {
    Pack pack;
    int rank, class_aware;  // class_aware 0 -> global-MMD ablation (I/utils.py:633-650)
    const uint8_t* mask;    // [B_global]
    const int* nmask;       // [1] (gate: > 1, I/train.py:444)
    const float* w_class;   // [4] W_c from DACP (ones for the fixed-threshold path)
    float lambda_cls, gamma_comp, delta_rep, w_ecda;
    // workspace
    float* part_vec;        // [G][8][256]
    double* part_sc;        // [G][8][2]  (sq-norm sum, weight sum)
    int* part_cnt;          // [G][8]
    EcdaStats* stats;
    int* lists;             // [2*B_global] member lists; per class: clean members then noisy members (row ids: clean gi, noisy B_global+gi)
    EcdaTile* tiles;        // [max_tiles]
    float* mu_t;            // [4][256] noisy centroids
    float* rep_grad;        // [4][256] rho * d rep / d z_i for noisy members of class c
    double* red_vec;        // [8][256] slot sums over all rows (k_ecda_reduce)
    double* red_sc;         // [8][2]
    int* red_cnt;           // [8]
    uint8_t* slot_of;       // [2*B_global] slot (dom*4 + class) of every row, 255 = not a member
    unsigned short* z_hi;   // [2*B_global][256] fp16 hi / lo split of the scaled, class-centred member embeddings, in list order
    unsigned short* z_lo;
    float* G;               // [2*B_local + 8*128][ldg] Gram rows of the member rows this rank owns (k_gram_tc)
    long long ldg;
    GramItem* gram_items;   // [max_gram_items]
    int max_gram_items;
    unsigned short* p_hi;   // [g_rows_alloc][ldg] fp16 hi / lo split of the scaled gradient weights P (k_ecda_p)
    unsigned short* p_lo;
    float* rowp_part;       // [ldg/256][g_rows_alloc] row sums of (scaled) P per 256-column block
    float* accz;            // [ECDA_PZ_SPLITS][g_rows_alloc][256] (scaled) P Z' split-K slices (k_pz_tc), summed in slice order by k_ecda_fold
    float* item_sums;       // [max_gram_items * 4][3] weighted block sums of K per k_ecda_p work unit (<= 4 per Gram item)
    float* pos_nsq;         // [2*B_global] |z'|^2, omega, domain of every list position (k_ecda_pack)
    float* pos_om;
    int* pos_dom;
    PzItem* pz_items;       // [max_pz_items]
    int max_pz_items, g_rows_alloc;
    float* mu_all;          // [4][256] mean of ALL members (clean + masked noisy) of class c: the centring point of the Gram (k_ecda_pack)
    float *g_clean, *g_strong;  // [B_local,256], accumulated (+=)
    float* loss_out;        // [2] = ecda loss share, [3] = full-batch ecda loss when world == 1
    unsigned* counter;      // [2] zero-initialised, self-resetting
    int max_tiles, prep_grid, phases;
};

// supervised contrastive loss (csrc/k_scl.cu)
struct SclArgs {
    dad_pack pack;
    int rank, noisy;
    const uint8_t* mask;   // [Bg] (noisy != 0)
    float inv_tau, w_scl;
    float *g_clean, *g_strong;  // [B_local, H] accumulated
    float* loss_out;       // [3] := this rank's share of the (unweighted) loss
    float* F;              // [2Bg, H] normalised embeddings (0 for non-members)
    float* nrm;            // [2Bg]
    int* cls;              // [2Bg] class or -1
    float* lse;            // [2Bg]
    float* invp;           // [2Bg] 1/|P(i)| for valid anchors else 0
    float* row_loss;       // [2Bg]
    int* cnt;              // [DAD_C] members per class (zeroed by the launcher)
};

