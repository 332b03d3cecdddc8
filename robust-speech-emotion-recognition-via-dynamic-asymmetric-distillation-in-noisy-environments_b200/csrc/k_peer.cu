// Data-parallel exchange over NVLink peer memory (no NCCL on the step's critical path).
//
// Every rank owns one SYMMETRIC exchange buffer (same layout everywhere, peer-mapped: e.g. torch symmetric memory
// `buffer_ptrs`):   [ flags : 64 x u32 | pack : world x rank_words f32 | grad stage : world x n_grad f32 ]
//   k_peer_push     writes this rank's slice of a region straight into every peer's copy (st.global over NVLink/NVSwitch,
//                   16-byte vectors) -- the all-gather of the pack and the "all-to-all" half of the gradient all-reduce;
//   k_peer_barrier  sequence-number barrier through the flag words (st.release.sys / ld.acquire.sys): rank r stores the
//                   barrier's sequence number into slot [r] of every peer's flags and spins until all slots of its own
//                   flags reached it.  The sequence counter lives in device memory and is advanced by the kernel itself, so
//                   a captured CUDA graph can be replayed;
//   k_grad_sum_ranks  g[e] = sum_r stage[r][e] in rank order: the same order on every rank, so the replicated optimiser
//                   state stays bit-identical (SURVEY.md 8e "Determinism").
// Two barriers per step (after the pack push, after the gradient push) also order the write-after-read hazards: a rank can
// only overwrite a peer's pack after that peer passed the gradient barrier of the previous step (all its pack reads are
// stream-ordered before it), and the gradient stage after the peer passed the pack barrier of the current step.
#include "dad_common.cuh"
#include "dad_optim.cuh"

namespace {

__global__ void __launch_bounds__(256) k_peer_push(const char* __restrict__ src, PeerPtrs peers, int world, int self_rank, int include_self,
                                                   size_t dst_off, size_t bytes) {
    const int peer = blockIdx.y;
    if (peer >= world || (peer == self_rank && !include_self)) return;
    char* dst = reinterpret_cast<char*>(peers.p[peer]) + dst_off;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
    const size_t n16 = vec ? bytes >> 4 : 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (size_t i = (n16 << 4) + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i + 4 <= bytes; i += stride * 4)
        *reinterpret_cast<uint32_t*>(dst + i) = *reinterpret_cast<const uint32_t*>(src + i);
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();
}

__global__ void __launch_bounds__(32) k_peer_barrier(PeerPtrs peers, int world, int rank, size_t flag_off, unsigned* local_seq) {
    __shared__ unsigned s_seq;
    if (threadIdx.x == 0) s_seq = ++(*local_seq);
    __syncthreads();
    peer_flag_barrier(peers, world, rank, flag_off, s_seq);
    __syncthreads();
    __threadfence_system();
}

// push + barrier in ONE launch (the pack all-gather of a step): every CTA copies its part of this rank's segment into every peer's
// buffer, fences at system scope and counts itself in; the LAST CTA to arrive runs the flag barrier.  (Separate push and barrier
// launches cost one more launch latency on the critical path of a step that is latency-bound at 512 samples per rank.)
__global__ void __launch_bounds__(256) k_peer_push_barrier(const char* __restrict__ src, PeerPtrs peers, int world, int self_rank, size_t dst_off,
                                                           size_t bytes, size_t flag_off, unsigned* local_seq, unsigned* counter) {
    __shared__ unsigned s_last, s_seq;
    // diagnostics (bench.py --gpus N): [flag block + 64] = start time of this launch, [+72] = accumulated ns from start to the
    // moment the last CTA has fenced its stores (= duration of the push itself, before the barrier)
    unsigned long long* dbg = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(peers.p[self_rank]) + flag_off + 64);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        dbg[0] = t;
    }
    const size_t n16 = bytes >> 4;  // segments are 16-byte multiples (launcher check)
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // every 16-byte chunk is loaded once and stored to all peers back to back (remote stores are posted: nothing waits per store)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = reinterpret_cast<const uint4*>(src)[i];
#pragma unroll
        for (int peer = 0; peer < DAD_MAX_WORLD; ++peer)
            if (peer < world && peer != self_rank) reinterpret_cast<uint4*>(reinterpret_cast<char*>(peers.p[peer]) + dst_off)[i] = v;
    }
    // ONE system-scope fence per CTA, after the CTA barrier (fences are cumulative over the stores the barrier orders before it);
    // a membar.sys in every thread serialises in the memory system and cost ~80 us per exchange at 8 ranks
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned prev = atomicAdd(counter, 1u);
        s_last = prev == gridDim.x - 1u;
        if (s_last) { *counter = 0u; s_seq = ++(*local_seq); }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const unsigned long long t0 = *reinterpret_cast<volatile unsigned long long*>(dbg);
        if (t > t0) dbg[1] += t - t0;
    }
    peer_flag_barrier(peers, world, self_rank, flag_off, s_seq);
    __syncthreads();
    __threadfence_system();
}

__global__ void __launch_bounds__(256) k_grad_sum_ranks(const float* __restrict__ stage, int world, int64_t n, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = __ldcg(stage + i);  // written by peers: bypass L1
    for (int r = 1; r < world; ++r) s += __ldcg(stage + (int64_t)r * n + i);
    out[i] = s;
}

}  // namespace

int launch_peer_push(const void* src, void* const* peers, int world, int self_rank, int include_self, size_t dst_off, size_t bytes, cudaStream_t st) {
    PeerPtrs pp;
    for (int r = 0; r < DAD_MAX_WORLD; ++r) pp.p[r] = r < world ? peers[r] : nullptr;
    int bx = (int)((bytes / 16 + 255) / 256);
    if (bx > 32) bx = 32;
    if (bx < 1) bx = 1;
    k_peer_push<<<dim3(bx, world), 256, 0, st>>>(reinterpret_cast<const char*>(src), pp, world, self_rank, include_self, dst_off, bytes);
    DAD_CHECK_LAUNCH("k_peer_push");
    return DAD_OK;
}
int launch_peer_barrier(void* const* peers, int world, int rank, size_t flag_off, unsigned* local_seq, cudaStream_t st) {
    PeerPtrs pp;
    for (int r = 0; r < DAD_MAX_WORLD; ++r) pp.p[r] = r < world ? peers[r] : nullptr;
    k_peer_barrier<<<1, 32, 0, st>>>(pp, world, rank, flag_off, local_seq);
    DAD_CHECK_LAUNCH("k_peer_barrier");
    return DAD_OK;
}
int launch_peer_push_barrier(const void* src, void* const* peers, int world, int self_rank, size_t dst_off, size_t bytes, size_t flag_off,
                             unsigned* local_seq, unsigned* counter, cudaStream_t st) {
    if ((bytes & 15u) || ((reinterpret_cast<uintptr_t>(src) | dst_off) & 15u))
        return dad_set_error(DAD_ERR_ARG, "peer_push_barrier: segment must be 16-byte aligned and a multiple of 16 bytes");
    PeerPtrs pp;
    for (int r = 0; r < DAD_MAX_WORLD; ++r) pp.p[r] = r < world ? peers[r] : nullptr;
    int bx = (int)((bytes / 16 + 255) / 256);
    if (bx > 148) bx = 148;
    if (bx < 1) bx = 1;
    k_peer_push_barrier<<<bx, 256, 0, st>>>(reinterpret_cast<const char*>(src), pp, world, self_rank, dst_off, bytes, flag_off, local_seq, counter);
    DAD_CHECK_LAUNCH("k_peer_push_barrier");
    return DAD_OK;
}
int launch_grad_sum_ranks(const float* stage, int world, int64_t n, float* out, cudaStream_t st) {
    k_grad_sum_ranks<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(stage, world, n, out);
    DAD_CHECK_LAUNCH("k_grad_sum_ranks");
    return DAD_OK;
}
