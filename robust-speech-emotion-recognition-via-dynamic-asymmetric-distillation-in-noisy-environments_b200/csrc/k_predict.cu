// Inference heads: classifier logits + arg-max of the student and, optionally, the teacher embedding of the same
// batch, for SSRLModel.predict / validate / anchor calibration (I/model.py:225-245, I/train.py:317-357,522-564).  The two embeddings
// come from ONE pass of the tcgen05 encoder kernel over x (k_enc_fwd_tc<2> with a shared A operand).  eval mode: no dropout.
// One warp per sample: lane l holds hidden units l, l+32, ... (8 values); 4 class dot products, butterfly sums, lane 0 writes.
#include "dad_common.cuh"

namespace {

struct PredictHeads {
    const float* emb[2];   // [B,256]
    const float* Wc[2];    // [4,256]
    const float* bc[2];    // [4]
    float* logits[2];      // [B,4]
    int32_t* pred[2];      // [B] (nullable)
    int views;
};

__global__ void __launch_bounds__(256) k_predict_head(PredictHeads a, int B) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + warp;
    if (b >= B) return;
    for (int v = 0; v < a.views; ++v) {
        float e[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = __ldg(a.emb[v] + (int64_t)b * DAD_H + lane + 32 * k);
        float z[DAD_C];
#pragma unroll
        for (int c = 0; c < DAD_C; ++c) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s = fmaf(e[k], __ldg(a.Wc[v] + c * DAD_H + lane + 32 * k), s);
            z[c] = warp_sum(s) + __ldg(a.bc[v] + c);
        }
        if (lane == 0) {
            *reinterpret_cast<float4*>(a.logits[v] + (int64_t)b * DAD_C) = make_float4(z[0], z[1], z[2], z[3]);
            if (a.pred[v]) {
                int best = 0;
#pragma unroll
                for (int c = 1; c < DAD_C; ++c)
                    if (z[c] > z[best]) best = c;  // first maximum, as torch.argmax
                a.pred[v][b] = best;
            }
        }
    }
}

}  // namespace

int launch_predict_heads(const float* emb_s, const float* Wc_s, const float* bc_s, float* logits_s, int32_t* pred_s, const float* emb_t,
                         const float* Wc_t, const float* bc_t, float* logits_t, int32_t* pred_t, int64_t B, cudaStream_t st) {
    PredictHeads a;
    a.emb[0] = emb_s; a.Wc[0] = Wc_s; a.bc[0] = bc_s; a.logits[0] = logits_s; a.pred[0] = pred_s;
    a.emb[1] = emb_t; a.Wc[1] = Wc_t; a.bc[1] = bc_t; a.logits[1] = logits_t; a.pred[1] = pred_t;
    a.views = emb_t ? 2 : 1;
    k_predict_head<<<(unsigned)ceil_div64(B, 8), 256, 0, st>>>(a, (int)B);
    DAD_CHECK_LAUNCH("k_predict_head");
    return DAD_OK;
}
