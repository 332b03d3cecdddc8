// extern "C" entry points of libdad_sm100a.so (see include/dad_sm100a.h): argument checks, workspace carving,
// kernel dispatch.  Never throws, never synchronises, never allocates.
#include "dad_common.cuh"
#include "dad_pack.cuh"
#include "dad_optim.cuh"
#include <string.h>

// ---- kernels (other translation units)
int launch_augment(const float* x, float* out, const AugParams& ap, int64_t B, int64_t T, int64_t D, cudaStream_t st);
int launch_step_rng(unsigned long long seed, const int64_t* step_ptr, float* keep_clean, float* keep_strong, int64_t B, float p_drop,
                    float* feat_keep, float feat_rate, int32_t* t_start, int64_t T, int mask_len, int64_t sample_offset, cudaStream_t st);
size_t enc_fwd_simt_ws_bytes(int64_t B, int64_t T);
int launch_enc_fwd_simt(const float* x, const uint8_t* pad, const float* W, const float* bias, const AugParams& ap, float* emb,
                        uint32_t* bits, float* lens_out, float* part_ws, unsigned* counters, int64_t B, int64_t T, cudaStream_t st);
int enc_bwd_simt_splits(int64_t N);
int launch_enc_bwd_simt(const float* x, const uint8_t* pad, const AugParams& ap, const uint32_t* bits, const float* g_emb, const float* lens,
                        float* dWp, float* dbp, int splits, int64_t B, int64_t T, cudaStream_t st);
size_t enc_fwd_tc_ws_bytes(int64_t B, int64_t T);
int launch_enc_fwd_tc(const float* x, const uint8_t* pad, const void* W0_hi, const void* W0_lo, const float* bias0, const AugParams& ap0, float* emb0,
                      uint32_t* bits0, const void* W1_hi, const void* W1_lo, const float* bias1, const AugParams* ap1, float* emb1, uint32_t* bits1,
                      void* xs_hi, void* xs_lo, float* lens_out, void* ws, size_t ws_bytes, int64_t B, int64_t T, cudaStream_t st);
int launch_predict_heads(const float* emb_s, const float* Wc_s, const float* bc_s, float* logits_s, int32_t* pred_s, const float* emb_t,
                         const float* Wc_t, const float* bc_t, float* logits_t, int32_t* pred_t, int64_t B, cudaStream_t st);
void enc_tc_set_cta_pair(int on);
int enc_tc_get_cta_pair();
void enc_tc_set_bwd_products(int n);
int enc_tc_get_bwd_products();
int enc_bwd_tc_splits(int64_t B);
size_t enc_bwd_tc_ws_bytes(int64_t B, int64_t T);
int launch_enc_bwd_tc(const float* x, const uint8_t* pad, const AugParams& ap, const uint32_t* bits, const void* xs_hi, const void* xs_lo,
                      const float* g_emb, const float* lens, float* dWp, float* dbp, int splits, void* ws, size_t ws_bytes, int64_t B, int64_t T,
                      cudaStream_t st);
int launch_head_fwd(const HeadFwdArgs& a, cudaStream_t st);
int head_grid(int B_local);
int launch_certainty(const float* probs, int B, int use_entropy, float* scores, int* preds, cudaStream_t st);
int launch_dacp(DacpArgs a, cudaStream_t st);
int launch_head_bwd(const HeadBwdArgs& a, cudaStream_t st);
int ecda_prep_grid(int B_global);
int ecda_max_tiles(int B_local);
int ecda_max_gram_items(int B_local, int world);
int ecda_max_pz_items(int B_local);
int launch_ecda(const EcdaArgs& a, cudaStream_t st);
int launch_scl(const SclArgs& a, cudaStream_t st);
int launch_gather_pad(const float* feats, const int64_t* offsets, const int32_t* sizes, float* x, uint8_t* pad, int64_t B, int64_t T, int64_t D,
                      cudaStream_t st);
int launch_peer_push(const void* src, void* const* peers, int world, int self_rank, int include_self, size_t dst_off, size_t bytes, cudaStream_t st);
int launch_peer_barrier(void* const* peers, int world, int rank, size_t flag_off, unsigned* local_seq, cudaStream_t st);
int launch_grad_sum_ranks(const float* stage, int world, int64_t n, float* out, cudaStream_t st);
int launch_peer_push_barrier(const void* src, void* const* peers, int world, int self_rank, size_t dst_off, size_t bytes, size_t flag_off,
                             unsigned* local_seq, unsigned* counter, cudaStream_t st);
int launch_grad_reduce(const GradReduceArgs& a, cudaStream_t st);
int launch_grad_norm(const NormArgs& a, cudaStream_t st);
int launch_adam_ema(const AdamArgs& a, cudaStream_t st);
int launch_ema(float* t, const float* s, float ema_m, float ema_1m, int n, cudaStream_t st);
int launch_split_f16(const float* w, void* hi, void* lo, int n, cudaStream_t st);

// ---- error plumbing
static thread_local char g_err[512] = "";
int dad_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int check_aug(const dad_aug* a, bool need) {
    if (!a) return need ? dad_set_error(DAD_ERR_ARG, "aug: null") : DAD_OK;
    DAD_REQUIRE(a->mode >= 0 && a->mode <= 2, "aug.mode must be 0,1,2 (got %d)", a->mode);
    if (a->mode == 0) return DAD_OK;
    DAD_REQUIRE(a->noise_src == 0 || a->noise_src == 1, "aug.noise_src must be 0 or 1");
    DAD_REQUIRE(a->noise_src == 1 || (a->noise && aligned16(a->noise)), "aug.noise: null or unaligned supplied-noise tensor");
    DAD_REQUIRE(!a->feat_keep || aligned16(a->feat_keep), "aug.feat_keep must be 16-byte aligned");
    DAD_REQUIRE(a->mask_len >= 0, "aug.mask_len < 0");
    return DAD_OK;
}

extern "C" {

int dad_abi_version(void) { return DAD_ABI_VERSION; }
int dad_max_world(void) { return DAD_MAX_WORLD; }
const char* dad_last_error(void) { return g_err; }

int dad_device_check(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "no CUDA device");
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return dad_set_error(DAD_ERR_UNSUPPORTED, "libdad_sm100a needs a compute-capability 10.x GPU (got %d.x)", major);
    return DAD_OK;
}

int64_t dad_pack_rank_words(int32_t B_local) { return pack_rank_words(B_local); }

int dad_augment(const float* x, float* out, const dad_aug* aug, int64_t B, int64_t T, void* stream) {
    DAD_REQUIRE(B >= 0 && T >= 0, "augment: negative size");
    if (B * T == 0) return DAD_OK;
    DAD_REQUIRE(x && out && aligned16(x) && aligned16(out), "augment: null or unaligned x/out");
    int rc = check_aug(aug, true);
    if (rc) return rc;
    DAD_REQUIRE(aug->mode != 0, "augment: mode 0 is a copy; nothing to do");
    return launch_augment(x, out, *aug, B, T, DAD_D, (cudaStream_t)stream);
}

int dad_step_rng(uint64_t seed, const int64_t* step_ptr, float* keep_clean, float* keep_strong, int64_t B, float p_drop, float* feat_keep,
                 float feat_rate, int32_t* t_start, int64_t T, int32_t mask_len, int64_t sample_offset, void* stream) {
    DAD_REQUIRE(B >= 0 && T >= 0 && mask_len >= 0 && sample_offset >= 0, "step_rng: negative size");
    return launch_step_rng(seed, step_ptr, keep_clean, keep_strong, B, p_drop, feat_keep, feat_rate, t_start, T, mask_len, sample_offset,
                           (cudaStream_t)stream);
}

// ---------------------------------------------------------------- encoder
size_t dad_encoder_fwd_workspace_bytes(int64_t B, int64_t T, int32_t gemm_mode) {
    const size_t counters = align_up((size_t)(2 * B) * sizeof(unsigned), 256);
    if (gemm_mode == DAD_GEMM_TC_F16X3) return counters + enc_fwd_tc_ws_bytes(B, T);
    return counters + enc_fwd_simt_ws_bytes(B, T);
}

int dad_encoder_fwd(const float* x, const uint8_t* pad, const float* W, const float* bias, const void* W_hi, const void* W_lo,
                    const dad_aug* aug, float* emb, uint32_t* relu_bits, void* xs_hi, void* xs_lo, float* lens, void* ws, size_t ws_bytes,
                    int64_t B, int64_t T, int32_t gemm_mode, void* stream) {
    DAD_REQUIRE(B >= 0 && T >= 0, "encoder_fwd: negative size");
    if (B == 0) return DAD_OK;
    DAD_REQUIRE(T > 0, "encoder_fwd: T must be > 0");
    DAD_REQUIRE(B <= 0x7fffffff / 2 && T <= 0x7fffffff, "encoder_fwd: size overflow");
    DAD_REQUIRE(x && aligned16(x), "encoder_fwd: x null or not 16-byte aligned");
    DAD_REQUIRE(bias && emb && aligned16(emb), "encoder_fwd: bias/emb null or unaligned");
    int rc = check_aug(aug, false);
    if (rc) return rc;
    AugParams ap;
    memset(&ap, 0, sizeof(ap));
    if (aug) ap = *aug;
    const size_t need = dad_encoder_fwd_workspace_bytes(B, T, gemm_mode);
    if (!ws || ws_bytes < need) return dad_set_error(DAD_ERR_WS, "encoder_fwd: workspace %zu < %zu bytes", ws_bytes, need);
    unsigned* counters = reinterpret_cast<unsigned*>(ws);
    char* rest = reinterpret_cast<char*>(ws) + align_up((size_t)(2 * B) * sizeof(unsigned), 256);
    if (gemm_mode == DAD_GEMM_FP32) {
        DAD_REQUIRE(W && aligned16(W), "encoder_fwd: W null or unaligned");
        // the arrival counters occupy the first 2*B words of the workspace, i.e. a region whose SIZE depends on B: a caller that
        // reuses one workspace across batch sizes (ragged last batch, cached op workspace) would hand over words that a previous,
        // smaller launch used for pooled partials.  Zero them here (a memset node under graph capture) instead of relying on the
        // caller's zero-initialisation plus the kernel's self-reset.
        cudaError_t e = cudaMemsetAsync(counters, 0, (size_t)(2 * B) * sizeof(unsigned), (cudaStream_t)stream);
        if (e != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "encoder_fwd: counter memset: %s", cudaGetErrorString(e));
        return launch_enc_fwd_simt(x, pad, W, bias, ap, emb, relu_bits, lens, reinterpret_cast<float*>(rest), counters, B, T, (cudaStream_t)stream);
    } else if (gemm_mode == DAD_GEMM_TC_F16X3) {
        DAD_REQUIRE(W_hi && W_lo, "encoder_fwd: TC mode needs the fp16 hi/lo weight splits (dad_split_f16)");
        DAD_REQUIRE(xs_hi || !xs_lo, "encoder_fwd: xs_lo without xs_hi");
        DAD_REQUIRE(!xs_hi || (aligned16(xs_hi) && aligned16(xs_lo)), "encoder_fwd: xs_hi/xs_lo not 16-byte aligned");
        return launch_enc_fwd_tc(x, pad, W_hi, W_lo, bias, ap, emb, relu_bits, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, xs_hi, xs_lo, lens,
                                 rest, ws_bytes - (size_t)(rest - (char*)ws), B, T, (cudaStream_t)stream);
    }
    return dad_set_error(DAD_ERR_ARG, "encoder_fwd: unknown gemm_mode %d", gemm_mode);
}

// fused noisy forward (I/train.py:405-409,439): weak view -> teacher, strong view -> student, x_noisy read once
size_t dad_aug_encoder_fwd_workspace_bytes(int64_t B, int64_t T) { return enc_fwd_tc_ws_bytes(B, T); }
int dad_aug_encoder_fwd(const float* x, const uint8_t* pad, const void* Wt_hi, const void* Wt_lo, const float* bias_t, const dad_aug* aug_weak,
                        float* emb_teacher, const void* Ws_hi, const void* Ws_lo, const float* bias_s, const dad_aug* aug_strong, float* emb_student,
                        uint32_t* relu_bits, void* xs_hi, void* xs_lo, float* lens, void* ws, size_t ws_bytes, int64_t B, int64_t T, void* stream) {
    DAD_REQUIRE(B >= 0 && T >= 0, "aug_encoder_fwd: negative size");
    if (B == 0) return DAD_OK;
    DAD_REQUIRE(T > 0 && B <= 0x7fffffff / 2 && T <= 0x7fffffff, "aug_encoder_fwd: bad size");
    DAD_REQUIRE(x && aligned16(x), "aug_encoder_fwd: x null or not 16-byte aligned");
    DAD_REQUIRE(Wt_hi && Wt_lo && Ws_hi && Ws_lo && bias_t && bias_s, "aug_encoder_fwd: null weight split / bias");
    DAD_REQUIRE(emb_teacher && emb_student && aligned16(emb_teacher) && aligned16(emb_student), "aug_encoder_fwd: emb null or unaligned");
    DAD_REQUIRE(aug_weak && aug_strong, "aug_encoder_fwd: both augmentation descriptors are required");
    DAD_REQUIRE(xs_hi || !xs_lo, "aug_encoder_fwd: xs_lo without xs_hi");
    int rc = check_aug(aug_weak, true);
    if (!rc) rc = check_aug(aug_strong, true);
    if (rc) return rc;
    const size_t need = dad_aug_encoder_fwd_workspace_bytes(B, T);
    if (!ws || ws_bytes < need) return dad_set_error(DAD_ERR_WS, "aug_encoder_fwd: workspace %zu < %zu bytes", ws_bytes, need);
    return launch_enc_fwd_tc(x, pad, Wt_hi, Wt_lo, bias_t, *aug_weak, emb_teacher, nullptr, Ws_hi, Ws_lo, bias_s, aug_strong, emb_student, relu_bits, xs_hi,
                             xs_lo, lens, ws, ws_bytes, B, T, (cudaStream_t)stream);
}

// fused inference: encoder + classifier + arg-max of the student and (optionally) the teacher from ONE read of x
size_t dad_predict_workspace_bytes(int64_t B, int64_t T) { return enc_fwd_tc_ws_bytes(B, T); }
int dad_predict(const float* x, const uint8_t* pad, const void* Ws_hi, const void* Ws_lo, const float* bias_s, const float* Wc_s, const float* bc_s,
                const void* Wt_hi, const void* Wt_lo, const float* bias_t, const float* Wc_t, const float* bc_t, float* emb_s, float* emb_t,
                float* logits_s, float* logits_t, int32_t* pred_s, int32_t* pred_t, void* ws, size_t ws_bytes, int64_t B, int64_t T, void* stream) {
    DAD_REQUIRE(B >= 0 && T >= 0, "predict: negative size");
    if (B == 0) return DAD_OK;
    DAD_REQUIRE(T > 0 && B <= 0x7fffffff / 2 && T <= 0x7fffffff, "predict: bad size");
    DAD_REQUIRE(x && aligned16(x), "predict: x null or not 16-byte aligned");
    DAD_REQUIRE(Ws_hi && Ws_lo && bias_s && Wc_s && bc_s && emb_s && logits_s && aligned16(emb_s) && aligned16(logits_s), "predict: null / unaligned student tensor");
    const bool two = Wt_hi != nullptr;
    DAD_REQUIRE(!two || (Wt_lo && bias_t && Wc_t && bc_t && emb_t && logits_t && aligned16(emb_t) && aligned16(logits_t)), "predict: null / unaligned teacher tensor");
    const size_t need = dad_predict_workspace_bytes(B, T);
    if (!ws || ws_bytes < need) return dad_set_error(DAD_ERR_WS, "predict: workspace %zu < %zu bytes", ws_bytes, need);
    AugParams ap;
    memset(&ap, 0, sizeof(ap));
    int rc = launch_enc_fwd_tc(x, pad, Ws_hi, Ws_lo, bias_s, ap, emb_s, nullptr, two ? Wt_hi : nullptr, two ? Wt_lo : nullptr, bias_t, two ? &ap : nullptr,
                               emb_t, nullptr, nullptr, nullptr, nullptr, ws, ws_bytes, B, T, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_predict_heads(emb_s, Wc_s, bc_s, logits_s, pred_s, two ? emb_t : nullptr, Wc_t, bc_t, logits_t, pred_t, B, (cudaStream_t)stream);
}

/* process-wide tuning switches of the tcgen05 encoder kernels */
void dad_tc_set_cta_pair(int32_t on) { enc_tc_set_cta_pair(on); }
int32_t dad_tc_get_cta_pair(void) { return enc_tc_get_cta_pair(); }
void dad_tc_set_bwd_products(int32_t n) { enc_tc_set_bwd_products(n); }
int32_t dad_tc_get_bwd_products(void) { return enc_tc_get_bwd_products(); }

int32_t dad_encoder_bwd_splits(int64_t B, int64_t T, int32_t gemm_mode) {
    return gemm_mode == DAD_GEMM_TC_F16X3 ? enc_bwd_tc_splits(B) : enc_bwd_simt_splits(B * T);
}
size_t dad_encoder_bwd_workspace_bytes(int64_t B, int64_t T, int32_t gemm_mode) {
    return gemm_mode == DAD_GEMM_TC_F16X3 ? enc_bwd_tc_ws_bytes(B, T) : 256;
}

int dad_encoder_bwd(const float* x, const uint8_t* pad, const dad_aug* aug, const uint32_t* relu_bits, const void* xs_hi, const void* xs_lo,
                    const float* g_emb, const float* lens, float* dW_part, float* db_part, int32_t splits, void* ws, size_t ws_bytes,
                    int64_t B, int64_t T, int32_t gemm_mode, void* stream) {
    DAD_REQUIRE(B > 0 && T > 0, "encoder_bwd: empty batch");
    DAD_REQUIRE(xs_hi || !xs_lo, "encoder_bwd: xs_lo without xs_hi");
    const bool cached = xs_hi != nullptr && gemm_mode == DAD_GEMM_TC_F16X3;
    DAD_REQUIRE(cached || (x && aligned16(x)), "encoder_bwd: x null/unaligned (and no operand cache given)");
    DAD_REQUIRE(relu_bits && g_emb && aligned16(g_emb) && lens, "encoder_bwd: null/unaligned input");
    DAD_REQUIRE(dW_part && db_part && aligned16(dW_part), "encoder_bwd: null/unaligned partial buffers");
    DAD_REQUIRE(splits == dad_encoder_bwd_splits(B, T, gemm_mode), "encoder_bwd: splits must equal dad_encoder_bwd_splits()");
    int rc = check_aug(aug, false);
    if (rc) return rc;
    AugParams ap;
    memset(&ap, 0, sizeof(ap));
    if (aug) ap = *aug;
    if (gemm_mode == DAD_GEMM_FP32)
        return launch_enc_bwd_simt(x, pad, ap, relu_bits, g_emb, lens, dW_part, db_part, splits, B, T, (cudaStream_t)stream);
    if (gemm_mode == DAD_GEMM_TC_F16X3) {
        const size_t need = enc_bwd_tc_ws_bytes(B, T);
        if (!ws || ws_bytes < need) return dad_set_error(DAD_ERR_WS, "encoder_bwd: workspace %zu < %zu bytes", ws_bytes, need);
        return launch_enc_bwd_tc(x, pad, ap, relu_bits, xs_hi, xs_lo, g_emb, lens, dW_part, db_part, splits, ws, ws_bytes, B, T, (cudaStream_t)stream);
    }
    return dad_set_error(DAD_ERR_ARG, "encoder_bwd: unknown gemm_mode %d", gemm_mode);
}

// ---------------------------------------------------------------- heads / DACP
int32_t dad_head_grid(int32_t B_local) { return head_grid(B_local); }

static int check_pack(const dad_pack& p, const char* who) {
    DAD_REQUIRE(p.base && aligned16(p.base), "%s: pack.base null or unaligned", who);
    DAD_REQUIRE(p.B_local > 0 && p.world > 0, "%s: pack.B_local / world must be > 0", who);
    DAD_REQUIRE((int64_t)p.B_local * p.world <= (1 << 24), "%s: global batch too large", who);
    return DAD_OK;
}

int dad_head_fwd(const dad_head_fwd_args* a, void* stream) {
    DAD_REQUIRE(a, "head_fwd: null args");
    int rc = check_pack(a->pack, "head_fwd");
    if (rc) return rc;
    DAD_REQUIRE(a->B_local == a->pack.B_local && a->rank >= 0 && a->rank < a->pack.world, "head_fwd: rank/B_local inconsistent with pack");
    DAD_REQUIRE(a->Wc && a->bc && a->dz_clean && a->ce_part, "head_fwd: null pointer");
    DAD_REQUIRE(!a->noisy || (a->emb_teacher && a->tWc && a->tbc && a->probs && a->logq), "head_fwd: noisy path needs teacher tensors");
    DAD_REQUIRE(a->n_clean >= 0 && a->n_clean <= a->B_local && a->n_noisy >= 0 && a->n_noisy <= a->B_local, "head_fwd: n_clean / n_noisy out of range");
    return launch_head_fwd(*a, (cudaStream_t)stream);
}

int dad_certainty_scores(const float* probs, int32_t B, int32_t use_entropy, float* scores, int32_t* preds, void* stream) {
    DAD_REQUIRE(B >= 0, "certainty_scores: negative size");
    if (B == 0) return DAD_OK;
    DAD_REQUIRE(probs && aligned16(probs) && scores && preds, "certainty_scores: null/unaligned pointer");
    return launch_certainty(probs, B, use_entropy, scores, preds, (cudaStream_t)stream);
}

int dad_dacp_mask(const dad_dacp_args* a, void* stream) {
    DAD_REQUIRE(a, "dacp: null args");
    int rc = check_pack(a->pack, "dacp");
    if (rc) return rc;
    DAD_REQUIRE(a->mask && a->out_w && a->out_thr && a->out_nmask, "dacp: null output");
    DAD_REQUIRE(!a->use_dacp || (a->quality && a->anchors && a->ema_thr && a->epoch_sum && a->epoch_cnt), "dacp: null state");
    return launch_dacp(*a, (cudaStream_t)stream);
}

int dad_head_bwd(const dad_head_bwd_args* a, void* stream) {
    DAD_REQUIRE(a, "head_bwd: null args");
    int rc = check_pack(a->pack, "head_bwd");
    if (rc) return rc;
    DAD_REQUIRE(a->B_local == a->pack.B_local && a->rank >= 0 && a->rank < a->pack.world, "head_bwd: rank/B_local inconsistent with pack");
    DAD_REQUIRE(a->Wc && a->dz_clean && a->ce_part && a->g_clean && a->dWc && a->dbc && a->part && a->loss_out && a->counter, "head_bwd: null pointer");
    DAD_REQUIRE(!a->noisy || (a->probs && a->logq && a->mask && a->nmask && a->g_strong), "head_bwd: noisy path needs probs/logq/mask");
    DAD_REQUIRE(a->n_ce_part == head_grid(a->B_local), "head_bwd: n_ce_part must equal dad_head_grid(B_local)");
    return launch_head_bwd(*a, (cudaStream_t)stream);
}

int dad_gather_pad(const float* feats, const int64_t* offsets, const int32_t* sizes, float* x, uint8_t* pad, int64_t B, int64_t T, void* stream) {
    DAD_REQUIRE(B >= 0 && T >= 0, "gather_pad: negative size");
    if (B == 0 || T == 0) return DAD_OK;
    DAD_REQUIRE(feats && aligned16(feats) && offsets && sizes && x && aligned16(x), "gather_pad: null/unaligned pointer");
    return launch_gather_pad(feats, offsets, sizes, x, pad, B, T, DAD_D, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- peer-memory exchange
static int check_peers(void* const* peers, int32_t world, const char* who) {
    DAD_REQUIRE(peers && world >= 1 && world <= DAD_MAX_WORLD, "%s: world must be in [1, %d]", who, DAD_MAX_WORLD);
    for (int r = 0; r < world; ++r) DAD_REQUIRE(peers[r] && aligned16(peers[r]), "%s: peers[%d] null or unaligned", who, r);
    return DAD_OK;
}
int dad_peer_push(const void* src_local, void* const* peers, int32_t world, int32_t self_rank, int32_t include_self, size_t dst_offset,
                  size_t bytes, void* stream) {
    int rc = check_peers(peers, world, "peer_push");
    if (rc) return rc;
    DAD_REQUIRE(src_local && (reinterpret_cast<uintptr_t>(src_local) & 3) == 0 && (dst_offset & 3) == 0 && (bytes & 3) == 0,
                "peer_push: src / offset / bytes must be multiples of 4 (16-byte alignment enables the vector path)");
    DAD_REQUIRE(self_rank >= 0 && self_rank < world, "peer_push: bad rank");
    if (bytes == 0) return DAD_OK;
    return launch_peer_push(src_local, peers, world, self_rank, include_self, dst_offset, bytes, (cudaStream_t)stream);
}
int dad_peer_barrier(void* const* peers, int32_t world, int32_t rank, size_t flag_offset, uint32_t* local_seq, void* stream) {
    int rc = check_peers(peers, world, "peer_barrier");
    if (rc) return rc;
    DAD_REQUIRE(rank >= 0 && rank < world && local_seq && (flag_offset & 3) == 0, "peer_barrier: bad rank / counter / offset");
    return launch_peer_barrier(peers, world, rank, flag_offset, local_seq, (cudaStream_t)stream);
}
int dad_peer_push_barrier(const void* src_local, void* const* peers, int32_t world, int32_t rank, size_t dst_offset, size_t bytes,
                          size_t flag_offset, uint32_t* local_seq, uint32_t* counter, void* stream) {
    int rc = check_peers(peers, world, "peer_push_barrier");
    if (rc) return rc;
    DAD_REQUIRE(src_local && rank >= 0 && rank < world && local_seq && counter && (flag_offset & 3) == 0, "peer_push_barrier: bad rank / counter / offset");
    return launch_peer_push_barrier(src_local, peers, world, rank, dst_offset, bytes, flag_offset, local_seq, counter, (cudaStream_t)stream);
}
int dad_grad_sum_ranks(const float* stage, int32_t world, int64_t n, float* out, void* stream) {
    DAD_REQUIRE(stage && out && world >= 1 && n >= 0, "grad_sum_ranks: bad arguments");
    if (n == 0) return DAD_OK;
    return launch_grad_sum_ranks(stage, world, n, out, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- SCL
struct SclCarve { size_t F, nrm, cls, lse, invp, row_loss, cnt, total; };
static SclCarve scl_carve(int B_local, int world) {
    const size_t R = (size_t)2 * B_local * world;
    SclCarve c;
    size_t o = 0;
    c.cnt = o; o = align_up(o + DAD_C * sizeof(int), 256);
    c.F = o; o = align_up(o + R * DAD_H * sizeof(float), 256);
    c.nrm = o; o = align_up(o + R * sizeof(float), 256);
    c.cls = o; o = align_up(o + R * sizeof(int), 256);
    c.lse = o; o = align_up(o + R * sizeof(float), 256);
    c.invp = o; o = align_up(o + R * sizeof(float), 256);
    c.row_loss = o; o = align_up(o + R * sizeof(float), 256);
    c.total = o;
    return c;
}
size_t dad_scl_workspace_bytes(int32_t B_local, int32_t world) { return scl_carve(B_local, world).total; }

int dad_scl(const dad_scl_args* a, void* stream) {
    DAD_REQUIRE(a, "scl: null args");
    int rc = check_pack(a->pack, "scl");
    if (rc) return rc;
    DAD_REQUIRE(a->rank >= 0 && a->rank < a->pack.world, "scl: bad rank");
    DAD_REQUIRE(a->g_clean && a->g_strong && a->loss_out && (!a->noisy || a->mask), "scl: null pointer");
    DAD_REQUIRE(a->temperature > 0.f, "scl: temperature must be > 0");
    const SclCarve c = scl_carve(a->pack.B_local, a->pack.world);
    if (!a->ws || a->ws_bytes < c.total) return dad_set_error(DAD_ERR_WS, "scl: workspace %zu < %zu bytes", a->ws_bytes, c.total);
    char* w = reinterpret_cast<char*>(a->ws);
    SclArgs k;
    k.pack = a->pack; k.rank = a->rank; k.noisy = a->noisy; k.mask = a->mask; k.inv_tau = 1.0f / a->temperature; k.w_scl = a->w_scl;
    k.g_clean = a->g_clean; k.g_strong = a->g_strong; k.loss_out = a->loss_out;
    k.F = (float*)(w + c.F); k.nrm = (float*)(w + c.nrm); k.cls = (int*)(w + c.cls); k.lse = (float*)(w + c.lse);
    k.invp = (float*)(w + c.invp); k.row_loss = (float*)(w + c.row_loss); k.cnt = (int*)(w + c.cnt);
    return launch_scl(k, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- ECDA
struct EcdaCarve { size_t part_vec, part_sc, part_cnt, stats, lists, tiles, mu_t, mu_all, rep_grad, slot_of, red_vec, red_sc, red_cnt, z_hi, z_lo, G, gram_items,
                   p_hi, p_lo, rowp_part, accz, item_sums, pos_nsq, pos_om, pos_dom, pz_items, counter, total; };
static EcdaCarve ecda_carve(int B_local, int world) {
    const int Bg = B_local * world, G = ecda_prep_grid(Bg), MT = ecda_max_tiles(B_local);
    const size_t g_rows = (size_t)2 * B_local + 8 * 128, ldg = align_up((size_t)2 * Bg, 256);
    EcdaCarve c;
    size_t o = 0;
    c.counter = o; o = align_up(o + 2 * sizeof(unsigned), 256);
    c.part_sc = o; o = align_up(o + (size_t)G * 8 * 2 * sizeof(double), 256);
    c.part_vec = o; o = align_up(o + (size_t)G * 8 * DAD_H * sizeof(float), 256);
    c.part_cnt = o; o = align_up(o + (size_t)G * 8 * sizeof(int), 256);
    c.stats = o; o = align_up(o + sizeof(EcdaStats), 256);
    c.lists = o; o = align_up(o + (size_t)2 * Bg * sizeof(int), 256);
    c.slot_of = o; o = align_up(o + (size_t)2 * Bg, 256);
    c.red_vec = o; o = align_up(o + (size_t)8 * DAD_H * sizeof(double), 256);
    c.red_sc = o; o = align_up(o + (size_t)16 * sizeof(double), 256);
    c.red_cnt = o; o = align_up(o + (size_t)8 * sizeof(int), 256);
    c.z_hi = o; o = align_up(o + (size_t)2 * Bg * DAD_H * sizeof(unsigned short), 1024);
    c.z_lo = o; o = align_up(o + (size_t)2 * Bg * DAD_H * sizeof(unsigned short), 1024);
    c.gram_items = o; o = align_up(o + (size_t)ecda_max_gram_items(B_local, world) * sizeof(GramItem), 256);
    c.G = o; o = align_up(o + g_rows * ldg * sizeof(float), 1024);
    c.p_hi = o; o = align_up(o + g_rows * ldg * sizeof(unsigned short), 1024);
    c.p_lo = o; o = align_up(o + g_rows * ldg * sizeof(unsigned short), 1024);
    c.rowp_part = o; o = align_up(o + (ldg / 256) * g_rows * sizeof(float), 256);
    c.accz = o; o = align_up(o + (size_t)ECDA_PZ_SPLITS * g_rows * DAD_H * sizeof(float), 256);
    c.item_sums = o; o = align_up(o + (size_t)ecda_max_gram_items(B_local, world) * 4 * 3 * sizeof(float), 256);
    c.pos_nsq = o; o = align_up(o + (size_t)2 * Bg * sizeof(float), 256);
    c.pos_om = o; o = align_up(o + (size_t)2 * Bg * sizeof(float), 256);
    c.pos_dom = o; o = align_up(o + (size_t)2 * Bg * sizeof(int), 256);
    c.pz_items = o; o = align_up(o + (size_t)ecda_max_pz_items(B_local) * sizeof(PzItem), 256);
    c.tiles = o; o = align_up(o + (size_t)MT * sizeof(EcdaTile), 256);
    c.mu_t = o; o = align_up(o + (size_t)ECDA_MAXC * DAD_H * sizeof(float), 256);
    c.rep_grad = o; o = align_up(o + (size_t)ECDA_MAXC * DAD_H * sizeof(float), 256);
    c.mu_all = o; o = align_up(o + (size_t)ECDA_MAXC * DAD_H * sizeof(float), 256);
    c.total = o;
    return c;
}
size_t dad_ecda_workspace_bytes(int32_t B_local, int32_t world) { return ecda_carve(B_local, world).total; }

int dad_ecda(const dad_ecda_args* a, void* stream) {
    DAD_REQUIRE(a, "ecda: null args");
    int rc = check_pack(a->pack, "ecda");
    if (rc) return rc;
    DAD_REQUIRE(a->rank >= 0 && a->rank < a->pack.world, "ecda: bad rank");
    DAD_REQUIRE(a->mask && a->nmask && a->w_class && a->g_clean && a->g_strong && a->loss_out, "ecda: null pointer");
    const EcdaCarve c = ecda_carve(a->pack.B_local, a->pack.world);
    if (!a->ws || a->ws_bytes < c.total) return dad_set_error(DAD_ERR_WS, "ecda: workspace %zu < %zu bytes", a->ws_bytes, c.total);
    char* w = reinterpret_cast<char*>(a->ws);
    EcdaArgs k;
    k.pack = a->pack; k.rank = a->rank; k.class_aware = a->class_aware; k.mask = a->mask; k.nmask = a->nmask; k.w_class = a->w_class;
    k.lambda_cls = a->lambda_cls; k.gamma_comp = a->gamma_comp; k.delta_rep = a->delta_rep; k.w_ecda = a->w_ecda;
    k.part_vec = (float*)(w + c.part_vec); k.part_sc = (double*)(w + c.part_sc); k.part_cnt = (int*)(w + c.part_cnt);
    k.stats = (EcdaStats*)(w + c.stats); k.lists = (int*)(w + c.lists); k.tiles = (EcdaTile*)(w + c.tiles);
    k.mu_t = (float*)(w + c.mu_t); k.rep_grad = (float*)(w + c.rep_grad);
    k.g_clean = a->g_clean; k.g_strong = a->g_strong; k.loss_out = a->loss_out; k.counter = (unsigned*)(w + c.counter);
    k.slot_of = (uint8_t*)(w + c.slot_of);
    k.z_hi = (unsigned short*)(w + c.z_hi); k.z_lo = (unsigned short*)(w + c.z_lo); k.G = (float*)(w + c.G);
    k.ldg = (long long)align_up((size_t)2 * a->pack.B_local * a->pack.world, 256);
    k.gram_items = (GramItem*)(w + c.gram_items); k.max_gram_items = ecda_max_gram_items(a->pack.B_local, a->pack.world);
    k.red_vec = (double*)(w + c.red_vec); k.red_sc = (double*)(w + c.red_sc); k.red_cnt = (int*)(w + c.red_cnt);
    k.mu_all = (float*)(w + c.mu_all);
    k.p_hi = (unsigned short*)(w + c.p_hi); k.p_lo = (unsigned short*)(w + c.p_lo); k.rowp_part = (float*)(w + c.rowp_part);
    k.accz = (float*)(w + c.accz); k.item_sums = (float*)(w + c.item_sums); k.pos_nsq = (float*)(w + c.pos_nsq);
    k.pos_om = (float*)(w + c.pos_om); k.pos_dom = (int*)(w + c.pos_dom); k.pz_items = (PzItem*)(w + c.pz_items);
    k.max_pz_items = ecda_max_pz_items(a->pack.B_local); k.g_rows_alloc = 2 * a->pack.B_local + 8 * 128;
    k.phases = (a->phases & 3) ? (a->phases & 3) : 3;
    k.max_tiles = ecda_max_tiles(a->pack.B_local);
    k.prep_grid = ecda_prep_grid(a->pack.B_local * a->pack.world);
    return launch_ecda(k, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- optimiser tail
int dad_grad_reduce(const float* dW_part0, const float* db_part0, int32_t splits0, const float* dW_part1, const float* db_part1,
                    int32_t splits1, const float* dWc, const float* dbc, float* g, void* stream) {
    DAD_REQUIRE(g && dWc && dbc, "grad_reduce: null pointer");
    DAD_REQUIRE((!dW_part0) == (!db_part0) && (!dW_part1) == (!db_part1), "grad_reduce: dW/db partials must come in pairs");
    GradReduceArgs a;
    a.dWp[0] = dW_part0; a.dbp[0] = db_part0; a.splits[0] = dW_part0 ? splits0 : 0;
    a.dWp[1] = dW_part1; a.dbp[1] = db_part1; a.splits[1] = dW_part1 ? splits1 : 0;
    a.dWc = dWc; a.dbc = dbc; a.g = g;
    memset(&a.x, 0, sizeof(a.x));
    return launch_grad_reduce(a, (cudaStream_t)stream);
}

// fused compute + exchange: split-K fold -> this rank's gradient into every rank's stage over NVLink -> flag barrier
int dad_grad_reduce_xchg(const float* dW_part0, const float* db_part0, int32_t splits0, const float* dW_part1, const float* db_part1,
                         int32_t splits1, const float* dWc, const float* dbc, float* g, void* const* peers, int32_t world, int32_t rank,
                         size_t stage_offset, size_t flag_offset, uint32_t* local_seq, uint32_t* counter, void* stream) {
    DAD_REQUIRE(g && dWc && dbc, "grad_reduce_xchg: null pointer");
    DAD_REQUIRE((!dW_part0) == (!db_part0) && (!dW_part1) == (!db_part1), "grad_reduce_xchg: dW/db partials must come in pairs");
    int rc = check_peers(peers, world, "grad_reduce_xchg");
    if (rc) return rc;
    DAD_REQUIRE(world >= 2 && rank >= 0 && rank < world && local_seq && counter && (stage_offset & 15) == 0 && (flag_offset & 3) == 0,
                "grad_reduce_xchg: bad world / rank / counters / offsets (stage offset must be 16-byte aligned)");
    GradReduceArgs a;
    a.dWp[0] = dW_part0; a.dbp[0] = db_part0; a.splits[0] = dW_part0 ? splits0 : 0;
    a.dWp[1] = dW_part1; a.dbp[1] = db_part1; a.splits[1] = dW_part1 ? splits1 : 0;
    a.dWc = dWc; a.dbc = dbc; a.g = g;
    for (int r = 0; r < DAD_MAX_WORLD; ++r) a.x.peers.p[r] = r < world ? peers[r] : nullptr;
    a.x.world = world; a.x.rank = rank; a.x.flag_off = flag_offset; a.x.stage_off = stage_offset; a.x.local_seq = local_seq; a.x.counter = counter;
    return launch_grad_reduce(a, (cudaStream_t)stream);
}

int dad_grad_norm(const float* g, double* part_ws, uint32_t* counter, int64_t* step, const double* lr, float* scal, float max_norm,
                  void* stream) {
    DAD_REQUIRE(g && part_ws && counter && step && lr && scal, "grad_norm: null pointer");
    NormArgs a;
    a.stage = nullptr; a.world = 1;
    a.g = const_cast<float*>(g); a.part = part_ws; a.counter = counter; a.step = (long long*)step; a.lr = lr; a.scal = scal; a.max_norm = max_norm; a.grad_scale = 1.f;
    return launch_grad_norm(a, (cudaStream_t)stream);
}

// rank-ordered sum of the staged gradients (second half of the all-reduce) fused into the grad-norm pass; g[DAD_NPARAM + 8] is written
int dad_grad_norm_xchg(float* g, const float* stage, int32_t world, double* part_ws, uint32_t* counter, int64_t* step, const double* lr,
                       float* scal, float max_norm, void* stream) {
    DAD_REQUIRE(g && stage && world >= 2 && world <= DAD_MAX_WORLD && part_ws && counter && step && lr && scal, "grad_norm_xchg: bad arguments");
    NormArgs a;
    a.stage = stage; a.world = world;
    a.g = g; a.part = part_ws; a.counter = counter; a.step = (long long*)step; a.lr = lr; a.scal = scal; a.max_norm = max_norm; a.grad_scale = 1.f;
    return launch_grad_norm(a, (cudaStream_t)stream);
}

int dad_adam_ema(float* p, float* t, float* m, float* v, const float* g, const float* scal, float wd, float ema_m, float ema_1m,
                 int32_t do_ema, void* s_hi, void* s_lo, void* t_hi, void* t_lo, const float* loss_src, float* loss_last,
                 double* loss_acc, float w_kd, float w_ecda, float w_scl, void* stream) {
    DAD_REQUIRE(p && t && m && v && g && scal, "adam_ema: null pointer");
    DAD_REQUIRE((!s_hi) == (!s_lo) && (!s_hi) == (!t_hi) && (!s_hi) == (!t_lo), "adam_ema: bf16 split outputs must be all set or all null");
    AdamArgs a;
    a.p = p; a.t = t; a.m = m; a.v = v; a.g = g; a.scal = scal; a.wd = wd; a.ema_m = ema_m; a.ema_1m = ema_1m; a.do_ema = do_ema;
    a.s_hi = s_hi; a.s_lo = s_lo; a.t_hi = t_hi; a.t_lo = t_lo;
    DAD_REQUIRE(!loss_src || (loss_last && loss_acc), "adam_ema: loss_src needs loss_last and loss_acc");
    a.loss_src = loss_src; a.loss_last = loss_last; a.loss_acc = loss_acc; a.w_kd = w_kd; a.w_ecda = w_ecda; a.w_scl = w_scl;
    return launch_adam_ema(a, (cudaStream_t)stream);
}

int dad_ema(float* t, const float* s, float ema_m, float ema_1m, int64_t n, void* stream) {
    DAD_REQUIRE(n >= 0 && n < (1ll << 31), "ema: bad n");
    DAD_REQUIRE(n == 0 || (t && s), "ema: null pointer");
    return launch_ema(t, s, ema_m, ema_1m, (int)n, (cudaStream_t)stream);
}

int dad_split_f16(const float* w, void* hi, void* lo, int64_t n, void* stream) {
    DAD_REQUIRE(n >= 0 && n < (1ll << 31), "split_f16: bad n");
    DAD_REQUIRE(n == 0 || (w && hi && lo), "split_f16: null pointer");
    return launch_split_f16(w, hi, lo, (int)n, (cudaStream_t)stream);
}

}  // extern "C"
