/* libdad_sm100a.so -- C ABI of the B200 (sm_100a) DAD training-step kernels.
 *
 * The reference (TMZZ22331/Robust-Speech-Emotion-Recognition-via-Dynamic-Asymmetric-Distillation-in-Noisy-Environments)
 * is pure Python/PyTorch and has no FFI of its own; each entry point below replaces the ATen call
 * sequence of the cited reference lines.  I/ = IEMOCAP/DAD-train-IEMOCAP/ (the CASIA / EMODB copies are identical
 * on this path except where noted).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller (outputs and workspace
 *    included); nothing is allocated, freed or retained by the library;
 *  - every op enqueues on `stream` (a cudaStream_t passed as void*), never synchronises the host and is CUDA-graph
 *    capturable; state that advances (DACP thresholds, Adam moments, step counter, parameters) is updated in place;
 *  - return value 0 = ok, negative = error (dad_last_error() holds a thread-local message); nothing throws or aborts;
 *  - workspaces must be zero-initialised once when allocated (they hold self-resetting arrival counters at offsets that do
 *    not depend on the batch shape, so one workspace sized for the largest batch serves every smaller one);
 *    use a separate workspace per concurrently-running op;
 *  - fp32 tensors are contiguous row-major and 16-byte aligned; D = 768, H = 256, C = 4 are compile-time widths
 *    (I/config.py:54-55,48).
 */
#ifndef DAD_SM100A_H
#define DAD_SM100A_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAD_ABI_VERSION 3
#define DAD_D 768
#define DAD_H 256
#define DAD_C 4
#define DAD_MAX_WORLD 16  /* ranks of one NVLink domain served by the peer-memory exchange */
#define DAD_NPARAM 197892 /* flat parameter buffer: [W 256x768 | b 256 | Wc 4x256 | bc 4] */

/* GEMM arithmetic of the encoder kernels */
#define DAD_GEMM_FP32 0      /* CUDA-core fp32 FMA: exact path, any B,T */
#define DAD_GEMM_TC_F16X3 1  /* tcgen05 (5th-gen tensor core, TMEM accumulators): power-of-two scaled fp16 hi/lo split, 3 products, fp32 accumulate */

int dad_abi_version(void);
int dad_max_world(void); /* DAD_MAX_WORLD */
const char* dad_last_error(void);
/* 0 when the current CUDA device is compute capability 10.x, else an error */
int dad_device_check(void);

/* ---- augmentation parameters (I/utils.py:317-375 DataAugmentation) ------------------------------------------- */
typedef struct dad_aug {
    int32_t mode;       /* 0 none, 1 weak (noise only), 2 strong (noise + feature keep + temporal span) */
    int32_t noise_src;  /* 0 = supplied `noise` tensor (parity mode, bit-exact), 1 = in-kernel Philox4x32-7 + Box-Muller */
    const float* noise; /* [B,T,D] unit normal draws (noise_src 0) */
    uint64_t seed;      /* Philox key (noise_src 1) */
    uint64_t stream;    /* Philox stream id: weak and strong views use different streams.  Streams 2k and 2k+1 share their Philox
                           blocks (one call yields the four normals of a float4 for BOTH): the weak / strong views of one batch
                           use such a pair so that dad_aug_encoder_fwd draws both from one call */
    float sigma;        /* WEAK_NOISE_STD / STRONG_NOISE_STD: out = x + (n*sigma), two fp32 roundings */
    const float* feat_keep;  /* [D] 0/1 multiplier shared by the whole batch, no rescale (I/utils.py:343-344); NULL = none */
    const int32_t* t_start;  /* [B] temporal span starts (I/utils.py:371); NULL = none */
    int32_t mask_len;        /* int(T_padded * TEMPORAL_MASK_RATIO) (I/utils.py:366); rows [s, s+len) := +0 */
    const int64_t* step_ptr; /* nullable device counter mixed into the Philox key, so a replayed CUDA graph draws fresh noise
                                every step while the forward and backward kernels of one step regenerate identical noise */
    int64_t frame_offset;    /* index of this call's first frame in the GLOBAL noise field (noise_src 1): rank * B_local * T under
                                data parallelism, so that the ranks draw disjoint parts of one i.i.d. field as the reference's
                                single randn_like over the whole batch does (I/utils.py:330,338); 0 on one GPU */
} dad_aug;

/* per-step device-side draws for the throughput mode (replaces the host RNG calls of I/model.py:62 nn.Dropout x2,
 * I/utils.py:343 torch.rand(D) and I/utils.py:371 B x torch.randint): keep_* [B,H] 0/1 with P(keep)=1-p_drop,
 * feat_keep [D] 0/1 with P(keep)=1-feat_rate, t_start [B] uniform in [0, max(1, T-mask_len+1)).  Any output may be NULL.
 * sample_offset = index of this call's first sample in the global batch (rank * B_local under data parallelism, else 0): the
 * per-sample draws (keep_*, t_start) of the ranks are disjoint parts of one field; feat_keep is the same on every rank (the
 * reference shares it over the whole batch). */
int dad_step_rng(uint64_t seed, const int64_t* step_ptr, float* keep_clean, float* keep_strong, int64_t B, float p_drop,
                 float* feat_keep, float feat_rate, int32_t* t_start, int64_t T, int32_t mask_len, int64_t sample_offset,
                 void* stream);

/* replaces DataAugmentation.weak_augment / strong_augment (I/utils.py:328-375): out[B,T,D] = aug(x) */
int dad_augment(const float* x, float* out, const dad_aug* aug, int64_t B, int64_t T, void* stream);

/* device-side batch assembly: the reference collator (I/dataload_clean.py:177-193) over a split whose
 * features stay resident in HBM.  feats [sum T, D] f32; offsets [B] i64 (first row of each utterance), sizes [B] i32 (frames);
 * out: x [B,T,D] zero-padded, pad [B,T] uint8 (1 = padded frame; nullable).  T >= max(sizes). */
int dad_gather_pad(const float* feats, const int64_t* offsets, const int32_t* sizes, float* x, uint8_t* pad, int64_t B,
                   int64_t T, void* stream);

/* ---- encoder (I/model.py:18-41 Emotion2VecEncoder.forward) and its autograd backward ------------------------- */
size_t dad_encoder_fwd_workspace_bytes(int64_t B, int64_t T, int32_t gemm_mode);
/* emb[B,H] = masked-mean_t relu(W aug(x)[b,t,:] + bias).  pad: [B,T] uint8, 1 = padded frame, NULL = none.
 * relu_bits (nullable): [B*T, H/32] uint32, bit j of word w of frame n = (pre[n, 32w+j] > 0) -- all the backward needs.
 * xs_hi / xs_lo (nullable, TC mode only): OPERAND CACHE, fp16 [B*T, D] buffers.  When given, the forward also writes the
 *   hi (and, if xs_lo is given, lo) half of the split of the augmented, power-of-two-scaled input it feeds to the tensor cores;
 *   dad_encoder_bwd given the same buffers loads its B operand from them by TMA instead of re-reading x and regenerating the
 *   augmentation (same values, bit-identical dW).  The default 2-product backward (dad_tc_set_bwd_products) needs xs_hi only:
 *   N*D*2 bytes written + read instead of the N*D*4 bytes of x read again.
 * lens (nullable): [B] valid length clamped to >= 1 (I/model.py:36).  W_hi/W_lo: fp16 [H,D] splits from dad_split_f16 (TC mode only). */
int dad_encoder_fwd(const float* x, const uint8_t* pad, const float* W, const float* bias, const void* W_hi,
                    const void* W_lo, const dad_aug* aug, float* emb, uint32_t* relu_bits, void* xs_hi, void* xs_lo,
                    float* lens, void* ws, size_t ws_bytes, int64_t B, int64_t T, int32_t gemm_mode, void* stream);

/* Fused noisy forward (SURVEY.md 8b `aug_encoder_pool_fwd`; replaces I/train.py:405-409,439 = weak_augment + teacher encoder,
 * strong_augment + student encoder, I/utils.py:328-375 + I/model.py:18-41 twice): x [B,T,D] is read ONCE; the weak view
 * (aug_weak, mode 1) feeds the teacher weights, the strong view (aug_strong, mode 2) the student weights, on the tcgen05 path.
 * Outputs: emb_teacher, emb_student [B,H]; relu_bits / xs_hi / xs_lo (nullable) of the STUDENT view as in dad_encoder_fwd.
 * In Philox mode the two descriptors must share seed / step_ptr / frame_offset and use the streams (2k, 2k+1). */
size_t dad_aug_encoder_fwd_workspace_bytes(int64_t B, int64_t T);
int dad_aug_encoder_fwd(const float* x, const uint8_t* pad, const void* Wt_hi, const void* Wt_lo, const float* bias_t,
                        const dad_aug* aug_weak, float* emb_teacher, const void* Ws_hi, const void* Ws_lo, const float* bias_s,
                        const dad_aug* aug_strong, float* emb_student, uint32_t* relu_bits, void* xs_hi, void* xs_lo, float* lens,
                        void* ws, size_t ws_bytes, int64_t B, int64_t T, void* stream);

/* Fused inference (replaces SSRLModel.predict's encoder + classifier, I/model.py:225-245, as used by validate
 * and its student / teacher disagreement rate, I/train.py:522-564, and by the anchor calibration, I/train.py:317-357): x [B,T,D] is
 * read ONCE on the tcgen05 path (T >= 32); the student and -- when Wt_hi != NULL -- the teacher encoder + classifier (eval mode:
 * no dropout) produce emb_* [B,H], logits_* [B,C] and pred_* [B] (first arg-max; nullable). */
size_t dad_predict_workspace_bytes(int64_t B, int64_t T);
int dad_predict(const float* x, const uint8_t* pad, const void* Ws_hi, const void* Ws_lo, const float* bias_s, const float* Wc_s,
                const float* bc_s, const void* Wt_hi, const void* Wt_lo, const float* bias_t, const float* Wc_t, const float* bc_t,
                float* emb_s, float* emb_t, float* logits_s, float* logits_t, int32_t* pred_s, int32_t* pred_t, void* ws,
                size_t ws_bytes, int64_t B, int64_t T, void* stream);

/* Process-wide switches of the tcgen05 encoder kernels (also DAD_TC_CTA2 / DAD_TC_BWD_PRODUCTS in the environment):
 *  cta_pair      forward kernels as CTA pairs (cluster of 2, tcgen05.mma.cta_group::2): 0 = never, 1 = always ([frames x units]
 *                orientation), 2 (default) = CTA pairs for the clean / inference launches (bound by the MMAs' shared-memory operand
 *                traffic), single CTAs for the one-view noisy launches, and the TRANSPOSED CTA-pair kernel k_enc_fwd_tp for
 *                dad_aug_encoder_fwd (weights as the A operand, frames as B, two tiles of accumulators in TMEM: DESIGN.md 5.1).
 *                Experiment switches of that kernel, environment only: DAD_TC_TP_NF=128 (256-frame tiles), DAD_TC_TP_TEAMS=3
 *                (24 converter warps), DAD_TC_DBG (timing experiments with wrong results; only in builds with -DDAD_TC_EXPERIMENTS)
 *  bwd_products  3 (default) = dW GEMM as dpre_hi*x_hi + dpre_hi*x_lo + dpre_lo*x_hi (fp32-grade); 2 = x_hi only (xs_lo unused):
 *                a third less tensor work and half the operand cache, but gradient errors of ~1e-4 relative on cancelling
 *                elements -- outside the parity tolerances, opt-in */
void dad_tc_set_cta_pair(int32_t on);
int32_t dad_tc_get_cta_pair(void);
void dad_tc_set_bwd_products(int32_t n);
int32_t dad_tc_get_bwd_products(void);

int32_t dad_encoder_bwd_splits(int64_t B, int64_t T, int32_t gemm_mode);
/* split-K partials: dW_part[splits, H, D], db_part[splits, H]; dad_grad_reduce sums them in fixed order.
 * g_emb[B,H] = dL/d emb; x / pad / aug identical to the forward call (the augmented view is regenerated) -- or, in TC
 * mode, xs_hi (/ xs_lo for the 3-product backward) = the operand cache the forward call filled (then x and aug are ignored
 * and may be NULL). */
int dad_encoder_bwd(const float* x, const uint8_t* pad, const dad_aug* aug, const uint32_t* relu_bits, const void* xs_hi,
                    const void* xs_lo, const float* g_emb, const float* lens, float* dW_part, float* db_part,
                    int32_t splits, void* ws, size_t ws_bytes, int64_t B, int64_t T, int32_t gemm_mode, void* stream);
size_t dad_encoder_bwd_workspace_bytes(int64_t B, int64_t T, int32_t gemm_mode);

/* ---- the per-step exchange buffer ("pack"): what the single all-gather of the data-parallel step moves ------- *
 * per rank, 4-byte words: [emb_clean B_local*H f32][emb_strong B_local*H f32][score B_local f32][pred B_local i32][label B_local i32] */
typedef struct dad_pack {
    float* base;      /* world * dad_pack_rank_words(B_local) words */
    int32_t B_local;
    int32_t world;
} dad_pack;
int64_t dad_pack_rank_words(int32_t B_local);

/* ---- classifier heads + CE + teacher certainty (I/model.py:54-64, I/train.py:399-400,408-410,439-440, I/utils.py:400-428) */
typedef struct dad_head_fwd_args {
    dad_pack pack;
    int32_t rank, B_local, noisy, use_entropy;
    const float* emb_teacher;              /* [B_local,H] teacher embedding of the weak view (noisy != 0) */
    const float *Wc, *bc, *tWc, *tbc;      /* student / teacher classifier [C,H],[C] */
    const float *keep_clean, *keep_strong; /* [B_local,H] dropout keep masks (0/1) or NULL (eval, p = 0) */
    float drop_scale;                      /* fl32(1/(1-p)) (ATen dropout: x * (keep / (1-p))) */
    float label_smoothing, inv_B_global;
    float *dz_clean, *probs, *logq;        /* out [B_local,C]: dCE/dlogits (/B_global), teacher probs, student log-probs */
    float* ce_part;                        /* out [dad_head_grid(B_local)] per-CTA CE sums */
    int32_t n_clean, n_noisy;              /* real samples among the B_local clean / noisy slots (the rest is padding of a ragged
                                              batch: label := -1, pred := -1, no loss, no gradient, never in a quantile or class);
                                              inv_B_global must be 1 / (global number of REAL clean samples) */
} dad_head_fwd_args;
int32_t dad_head_grid(int32_t B_local);
int dad_head_fwd(const dad_head_fwd_args* a, void* stream);

/* ---- DACP (I/utils.py:449-507 DACPManager.calculate_mask; fixed-threshold branch I/train.py:417-420) ---------- */
typedef struct dad_dacp_args {
    dad_pack pack;
    int32_t use_dacp, npow2_;              /* npow2_: filled by the library */
    float gamma;                           /* fl32(q_start + (q_end-q_start)*epoch/EPOCHS) */
    float sens_k, calib_lambda, alpha, one_minus_alpha, fixed_thr;
    const float* quality;                  /* [C] class_quality_scores Q (state, read) */
    const float* anchors;                  /* [C] calibrated anchors */
    float* ema_thr;                        /* [C] ema_thresholds (state, updated in place when use_dacp) */
    uint8_t* mask;                         /* out [B_global] */
    float *out_w, *out_thr;                /* out [C] class weights W_c, thresholds applied */
    int32_t* out_nmask;                    /* out [1] number of masked-in samples */
    double* epoch_sum;                     /* [C] state: running per-class score sums of the epoch (I/utils.py:503-505) */
    int64_t* epoch_cnt;                    /* [C] */
} dad_dacp_args;
int dad_dacp_mask(const dad_dacp_args* a, void* stream);
/* DACPManager.calculate_certainty_scores (I/utils.py:400-428) as a stand-alone op: probs [B,C] -> scores [B] f32, preds [B] i32 */
int dad_certainty_scores(const float* probs, int32_t B, int32_t use_entropy, float* scores, int32_t* preds, void* stream);

/* ---- KD + head backward (I/train.py:444-447 and autograd through I/model.py:62-63) ---------------------------- */
typedef struct dad_head_bwd_args {
    dad_pack pack;
    int32_t rank, B_local, noisy, n_ce_part;
    const float* Wc;
    const float *keep_clean, *keep_strong;
    float drop_scale, w_kd, inv_B_global;
    const float *dz_clean, *probs, *logq, *ce_part;
    const uint8_t* mask;                   /* [B_global] */
    const int32_t* nmask;                  /* [1] */
    float *g_clean, *g_strong;             /* out [B_local,H] dL/d emb (head part; dad_ecda accumulates on top) */
    float *dWc, *dbc;                      /* out [C,H],[C] */
    float* part;                           /* workspace [dad_head_grid, C*H+C+1] */
    float* loss_out;                       /* out [0] CE share, [1] KD share of this rank (shares sum to the global loss) */
    uint32_t* counter;                     /* [1] zero-initialised arrival counter */
} dad_head_bwd_args;
int dad_head_bwd(const dad_head_bwd_args* a, void* stream);

/* ---- ECDA forward+backward (I/utils.py:510-652 ECDALoss; gate I/train.py:444,450) ----------------------------- */
typedef struct dad_ecda_args {
    dad_pack pack;
    int32_t rank, class_aware;             /* class_aware 0 = global-MMD ablation (I/utils.py:633-650) */
    const uint8_t* mask;                   /* [B_global] */
    const int32_t* nmask;                  /* [1]; ECDA runs only if > 1 */
    const float* w_class;                  /* [C] DACP class weights (ones for the fixed-threshold path) */
    float lambda_cls, gamma_comp, delta_rep, w_ecda;
    float *g_clean, *g_strong;             /* [B_local,H] accumulated: += w_ecda * dECDA/d emb for the rows this rank owns */
    float* loss_out;                       /* out [2] = this rank's share of the (unweighted) ECDA loss */
    void* ws;
    size_t ws_bytes;
    int32_t phases;                        /* 0 or 3 = everything; 1 = statistics + the two tensor-core GEMMs (reads the pack, the mask
                                              and nothing of g_*), 2 = fold into g_* + loss.  Splitting lets the caller run phase 1
                                              concurrently with dad_head_bwd (which WRITES g_*) on another stream. */
} dad_ecda_args;
size_t dad_ecda_workspace_bytes(int32_t B_local, int32_t world);
int dad_ecda(const dad_ecda_args* a, void* stream);

/* ---- supervised contrastive loss (NOT in the reference any more: removed upstream, C/utils.py:377-378; old ctor
 * C/train_CASIA.py:396 temperature=0.1).  SupCon over L2-normalised embeddings of {clean samples} U {masked-in noisy samples with
 * their pseudo-labels}, forward + backward; parity unpinned (checked against the oracle's own restatement).  Default weight 0. */
typedef struct dad_scl_args {
    dad_pack pack;
    int32_t rank, noisy;                   /* noisy 0: clean members only (warm-up) */
    const uint8_t* mask;                   /* [B_global] (noisy != 0) */
    float temperature, w_scl;
    float *g_clean, *g_strong;             /* [B_local,H] accumulated: += w_scl * dSCL/d emb for the rows this rank owns */
    float* loss_out;                       /* out [3] = this rank's share of the (unweighted) SCL loss */
    void* ws;
    size_t ws_bytes;
} dad_scl_args;
size_t dad_scl_workspace_bytes(int32_t B_local, int32_t world);
int dad_scl(const dad_scl_args* a, void* stream);

/* ---- data-parallel exchange over NVLink peer memory (replaces the NCCL all-gather / all-reduce of the step) ----
 * Every rank owns one symmetric, peer-mapped exchange buffer with the same layout; `peers[r]` (HOST array of `world` device
 * pointers) is rank r's buffer as mapped into this process (e.g. torch symmetric-memory buffer_ptrs).  All three ops are
 * stream-ordered, never synchronise the host and are CUDA-graph capturable.
 *   dad_peer_push     copy `bytes` (multiple of 4) from `src_local` to offset `dst_offset` of every peer's buffer
 *                     (own buffer too when include_self)
 *   dad_peer_barrier  all ranks' earlier pushes are complete and visible; `flag_offset` = byte offset of a zero-initialised
 *                     256-BYTE flag block inside the exchange buffer -- bytes [0, 4*world): the barrier flags (one u32 per rank);
 *                     [64, 80): start time / accumulated duration of dad_peer_push_barrier's copy; [128, 256): 16 x u64
 *                     accumulated nanoseconds this rank waited for rank r at the odd / even barriers (diagnostics written by
 *                     the barrier itself, read by bench.py --gpus N) -- `local_seq` = zero-initialised device counter private
 *                     to this rank.  Every rank must call its barriers in the same order.
 *   dad_grad_sum_ranks  out[e] = sum over r = 0..world-1 of stage[r][e]  (fixed order: identical result on every rank) */
int dad_peer_push(const void* src_local, void* const* peers, int32_t world, int32_t self_rank, int32_t include_self,
                  size_t dst_offset, size_t bytes, void* stream);
int dad_peer_barrier(void* const* peers, int32_t world, int32_t rank, size_t flag_offset, uint32_t* local_seq, void* stream);
int dad_grad_sum_ranks(const float* stage, int32_t world, int64_t n, float* out, void* stream);
/* FUSED forms (what the step engine uses; SURVEY.md 8e "all-gather written from the producing kernel's tail, all-reduce consumed at the
 * head of the optimiser kernels"):
 *   dad_peer_push_barrier   push + barrier in one launch: the last CTA of the copy runs the flag barrier.  `counter` = zero-initialised
 *                           device word (self-resetting); the segment must be 16-byte aligned and a multiple of 16 bytes
 *   dad_grad_reduce_xchg    dad_grad_reduce whose result (+ the 8 loss-share words behind g) is stored straight into slot [rank] of EVERY
 *                           rank's gradient stage (byte offset `stage_offset`, [world][DAD_NPARAM + 8] f32) and whose last CTA runs the
 *                           flag barrier: split-K fold + all-to-all + barrier = one kernel
 *   dad_grad_norm_xchg      dad_grad_norm that first forms g = sum over ranks of `stage` in rank order (identical bits on every rank),
 *                           writes it to g[DAD_NPARAM + 8] and takes the norm of the sum */
int dad_peer_push_barrier(const void* src_local, void* const* peers, int32_t world, int32_t rank, size_t dst_offset, size_t bytes,
                          size_t flag_offset, uint32_t* local_seq, uint32_t* counter, void* stream);
int dad_grad_reduce_xchg(const float* dW_part0, const float* db_part0, int32_t splits0, const float* dW_part1,
                         const float* db_part1, int32_t splits1, const float* dWc, const float* dbc, float* g, void* const* peers,
                         int32_t world, int32_t rank, size_t stage_offset, size_t flag_offset, uint32_t* local_seq,
                         uint32_t* counter, void* stream);
int dad_grad_norm_xchg(float* g, const float* stage, int32_t world, double* part_ws, uint32_t* counter, int64_t* step,
                       const double* lr, float* scal, float max_norm, void* stream);

/* ---- optimiser tail (I/train.py:487-492; I/model.py:211-223) --------------------------------------------------- */
/* g[DAD_NPARAM] = fixed-order sum of the encoder split-K partials (up to two backward launches) ++ head grads */
int dad_grad_reduce(const float* dW_part0, const float* db_part0, int32_t splits0, const float* dW_part1,
                    const float* db_part1, int32_t splits1, const float* dWc, const float* dbc, float* g, void* stream);
/* global L2 norm -> scal[0] norm, [1] clip coef (max_norm <= 0: no clipping), [2] lr/(1-b1^t), [3] sqrt(1-b2^t); ++*step */
int dad_grad_norm(const float* g, double* part_ws /*[128]*/, uint32_t* counter, int64_t* step, const double* lr,
                  float* scal, float max_norm, void* stream);
/* clip -> Adam (betas .9/.999, eps 1e-8, coupled L2 wd) -> EMA teacher (do_ema) -> fp16 hi/lo splits of W, teacher W (nullable).
 * loss_src (nullable) = [CE, KD, ECDA, SCL] of this step: loss_last[5] = {total, CE, KD, ECDA, SCL},
 * loss_acc[6] += {total, CE, KD, ECDA, 1, SCL}  (device-side replacement of the per-step .item() reads of I/train.py:494-495;
 * total = CE + w_kd KD + w_ecda ECDA + w_scl SCL). */
int dad_adam_ema(float* p, float* t, float* m, float* v, const float* g, const float* scal, float wd, float ema_m,
                 float ema_1m, int32_t do_ema, void* s_hi, void* s_lo, void* t_hi, void* t_lo, const float* loss_src,
                 float* loss_last, double* loss_acc, float w_kd, float w_ecda, float w_scl, void* stream);
/* SSRLModel.update_teacher_ema on a flat buffer: t = t*m + s*(1-m) (two muls + add, no FMA) */
int dad_ema(float* t, const float* s, float ema_m, float ema_1m, int64_t n, void* stream);
/* fp16 hi/lo split of an fp32 encoder weight matrix: hi = fp16(256 w), lo = fp16(256 w - hi) (saturating) */
int dad_split_f16(const float* w, void* hi, void* lo, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAD_SM100A_H */
