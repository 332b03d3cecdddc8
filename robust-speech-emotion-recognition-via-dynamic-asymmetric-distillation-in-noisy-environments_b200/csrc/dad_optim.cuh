// Argument blocks of the optimiser-tail kernels (shared by k_optim.cu and dad_abi.cu).
#pragma once
#include "dad_common.cuh"

// One rank's view of the symmetric exchange buffers of all ranks (k_peer.cu): used by the kernels that FUSE a step of the exchange
// with their compute (gradient fold -> push over NVLink -> flag barrier; rank sum -> grad norm).
struct PeerPtrs { void* p[DAD_MAX_WORLD]; };
struct PeerXchg {
    PeerPtrs peers;         // every rank's exchange buffer as mapped into this process
    int world, rank;        // world <= 1: no exchange
    size_t flag_off;        // byte offset of the `world` u32 barrier flags inside an exchange buffer
    size_t stage_off;       // byte offset of the gradient stage [world][NP + 8] inside an exchange buffer
    unsigned* local_seq;    // this rank's barrier sequence counter (device memory, advanced by the kernel)
    unsigned* counter;      // zero-initialised arrival counter of the kernel's CTAs (self-resetting)
};
// Sequence-number barrier through the flag words (one warp; lanes < world).  Called by ONE CTA after every store of this rank that
// the peers must see has been followed by __threadfence_system().
__device__ __forceinline__ void peer_flag_barrier(const PeerPtrs& peers, int world, int rank, size_t flag_off, unsigned seq) {
    const int r = threadIdx.x;
    if (r < world) {
        unsigned* remote = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(peers.p[r]) + flag_off) + rank;
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(seq) : "memory");
        const unsigned* mine = reinterpret_cast<const unsigned*>(reinterpret_cast<char*>(peers.p[rank]) + flag_off) + r;
        unsigned v;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        } while ((int)(v - seq) < 0);
        // wait statistics (diagnostics, read by bench.py --gpus N): nanoseconds this rank spent waiting for rank r at the odd (pack)
        // and even (gradient) barriers of a step, accumulated in the unused upper half of the flag block [16 x u64 at byte 128]
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        unsigned long long* dbg = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(peers.p[rank]) + flag_off + 128) + (seq & 1u) * 8u + r;
        *dbg += t1 - t0;
    }
}

struct GradReduceArgs {
    const float* dWp[2];   // [S][H*D] split-K partials (clean launch, strong launch); nullptr -> absent
    const float* dbp[2];   // [S][H]
    int splits[2];
    const float *dWc, *dbc;  // reduced head grads
    float* g;              // [NP + 8]: gradient + the loss shares [CE, KD, ECDA, SCL, ...] written by the loss kernels before
    PeerXchg x;            // world > 1: the folded gradient (+ loss shares) goes straight into slot [rank] of EVERY rank's stage
                           //   (16-byte stores over NVLink) and the kernel's last CTA runs the flag barrier: compute + all-to-all + barrier
                           //   in one launch (k_grad_norm then sums the slots in rank order)
};
struct NormArgs {
    float* g;              // [NP + 8]: read (world <= 1) or WRITTEN as the rank-ordered sum of the staged gradients (world > 1)
    const float* stage;    // [world][NP + 8] this rank's gradient stage (filled by every rank's k_grad_reduce), nullptr when world <= 1
    int world;
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
