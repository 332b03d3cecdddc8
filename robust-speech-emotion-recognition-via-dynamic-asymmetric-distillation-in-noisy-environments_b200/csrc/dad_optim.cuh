// Argument blocks of the optimiser-tail kernels (shared by k_optim.cu and dad_abi.cu).
#pragma once
#include "dad_common.cuh"

struct GradReduceArgs {
    const float* dWp[2];   // [S][H*D] split-K partials (clean launch, strong launch); nullptr -> absent
    const float* dbp[2];   // [S][H]
    int splits[2];
    const float *dWc, *dbc;  // reduced head grads
    float* g;              // [NP (+ tail)]
};
struct NormArgs {
    const float* g;        // [NP]
    double* part;          // [grid]
    unsigned* counter;     // zero-init, self-resetting
    long long* step;       // device step counter (Adam t), incremented here
    const double* lr;      // [1] current learning rate (host updates once per epoch)
    float* scal;           // [0] grad norm, [1] clip coef, [2] step_size = lr/(1-b1^t), [3] sqrt(1-b2^t)
    float max_norm;        // <= 0: clipping disabled (coef = 1)
    float grad_scale;      // reserved (1.0)
};
struct AdamArgs {
    float *p, *t, *m, *v;  // [NP] student, teacher, Adam moments
    const float* g;        // [NP]
    const float* scal;     // from k_grad_norm
    float wd, ema_m, ema_1m;  // weight decay; fl32(momentum), fl32(1-momentum)
    int do_ema;            // 0 during warm-up (I/train.py:491)
    void *s_hi, *s_lo, *t_hi, *t_lo;  // [H*D] fp16 splits of W / teacher W (nullable)
    const float* loss_src;  // [4] CE, KD, ECDA, SCL of this step (global sums), nullable
    float* loss_last;       // [5] total, CE, KD, ECDA, SCL of this step
    double* loss_acc;       // [6] running sums (total, CE, KD, ECDA, steps, SCL) since the host last cleared them
    float w_kd, w_ecda, w_scl;
};
