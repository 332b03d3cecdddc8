// Encoder kernels on the 5th-gen tensor cores (tcgen05, accumulators in TMEM, weights fed by TMA).
//
// Arithmetic ("f16x3"): every fp32 operand is scaled by a power of two (exact) and split on the fly into fp16 hi + fp16 lo
// (hi = fp16(x), lo = fp16(x - hi): 22 significant bits together) and the product is formed as  hi*hi + hi*lo + lo*hi
// with fp32 accumulation in TMEM (~2^-21 relative per product; the dropped lo*lo term is 2^-22).  Same tensor-pipe cost
// as a bf16 split (kind::f16 runs fp16 and bf16 at one rate) but 16x tighter, which is what keeps the ECDA gradient and
// the ReLU sign decisions inside the parity tolerance (DESIGN.md "Precision").  Scales: x by 2^3, W by 2^8 (fixed: |x| <
// 8188, |W| < 255 before saturation), g/len by a per-launch power of two derived from its absmax.  Only the dense
// contractions run here; everything else stays on CUDA cores.
//
// Forward  (I/model.py:18-41):  pre[N,256] = aug(x)[N,768] . W^T,  emb[b] = sum_t valid relu(pre + bias) / len_b
//   k_enc_fwd_tc<NV, CTA2>: persistent CTAs (one per SM), 128-frame x 256-unit tiles, K loop of 24 x 32 (64-byte rows, SWIZZLE_64B).
//   NV = 2 is the fused noisy forward (I/train.py:405-409,439): x_noisy is read ONCE, the converters form the weak and the strong
//   view of every element (one Philox block serves both) and two accumulators receive weak x teacher-W and strong x student-W.
//   CTA2 pairs two SMs (cluster of 2, tcgen05.mma.cta_group::2, M = 256 over the pair): every CTA converts ITS 128 frames and
//   TMA-loads HALF of every W k-block; the leader's one thread issues the MMAs for both, the peer's producers arrive on the
//   leader's barriers through the cluster address space, commits are multicast to both CTAs.  Warp roles (640 threads):
//     warp 0       TMA producer: W_hi / W_lo k-blocks (fp16) -> smem ring (2-6 stages)
//     warp 1       MMA issuer  : 6 tcgen05.mma (3 products x 2 k-steps) per k-block and view into TMEM accumulators
//     warp 2       TMEM allocator
//     warp 3       operand-cache store: the student view's fp16 hi A tiles, smem -> xs_hi by TMA (B operand of the backward)
//     warps 4-11   converters  : coalesced LDG of fp32 x (+ fused weak/strong augmentation, I/utils.py:328-375),
//                                fp16 hi/lo split, swizzled st.shared of the A operand(s) (K-major)
//     warps 12-19  epilogue    : tcgen05.ld -> bias + ReLU + padding mask -> 1-bit ReLU mask + per-32-frame pooled partial sums
//   k_pool_finalize then sums the partials of each utterance in fixed order and divides by its valid length.
#include "dad_common.cuh"
#include "tc_ptx.cuh"
#include "tc_host.cuh"
#include <cuda_fp16.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>


namespace {

// Timing experiments (DAD_TC_DBG=<bits> in the environment: 1 one product instead of three, 2 no W re-loads, 4 no noise, 8 no A
// stores -- wrong results, tools/gpu_dbg.sh) exist only in builds with -DDAD_TC_EXPERIMENTS (DAD_NVCC_FLAGS=-DDAD_TC_EXPERIMENTS
// python csrc/build.py --force); in the product build the switch is a compile-time zero and the branches vanish.
#ifdef DAD_TC_EXPERIMENTS
#define TC_DBG(p, bit) (((p).dbg & (bit)) != 0)
#else
#define TC_DBG(p, bit) false
#endif

constexpr int BM = 128, BK = 32, KB = DAD_D / BK;  // 24 k-blocks of 32
constexpr int FWD_THREADS = 896;                                          // 4 control + 16 converter + 8 epilogue warps
constexpr int FWD_CONV_WARPS = 16, FWD_EPI_WARP0 = 20, N_EPI_WARPS = 8;
constexpr int CONV_WARP0 = 4, N_CONV_WARPS = 8, EPI_WARP0 = 12;           // backward kernel (512 threads)
constexpr uint32_t A_BYTES = BM * BK * 2;  // 8 KB: one of hi / lo of one view, [128 frames][32 k] fp16, 64-byte rows, SWIZZLE_64B
constexpr uint32_t TMEM_COLS = 512;        // two 128 x 256 fp32 accumulators

// NV = views per launch (1: one accumulator per tile, double-buffered across tiles; 2: weak -> teacher W and strong -> student W from
// ONE read of x, both accumulators belong to the same tile).  CTA2: CTA pair, cta_group::2 MMA, every CTA stages half of the W rows.
template <int NV, bool CTA2>
struct FwdCfg {
    static constexpr uint32_t W_ROWS = CTA2 ? 128 : 256;
    static constexpr uint32_t W_BYTES = W_ROWS * BK * 2;  // 8 / 16 KB (one of hi / lo of one view)
    static constexpr uint32_t OFF_A = 0;                   // [view][hi, lo]
    static constexpr uint32_t OFF_W = NV * 2 * A_BYTES;    // [view][hi, lo]
    static constexpr uint32_t STAGE_BYTES = NV * 2 * (A_BYTES + W_BYTES);  // 48 / 32 / 96 / 64 KB
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;              // 4 / 6 / 2 / 3
    static constexpr uint32_t OFF_SCRATCH = STAGES * STAGE_BYTES;          // 8 epilogue warps x [32 rows][8 words + 1 pad]
    static constexpr uint32_t OFF_BIAS = OFF_SCRATCH + N_EPI_WARPS * 32 * 9 * 4;
    static constexpr uint32_t OFF_BARS = OFF_BIAS + 2 * DAD_H * 4;
    static constexpr uint32_t SMEM = OFF_BARS + 256 + 1024;  // + alignment slack
    static constexpr int MAX_STAGES = 6;
};

struct FwdView {
    const float* bias;  // [256] of the network this view feeds
    AugParams ap;
    float* part;        // [num_groups][2][256] pooled partial sums
    uint32_t* bits;     // [N][8] ReLU bits or null
};
struct FwdParams {
    const float* x;
    const uint8_t* pad;
    FwdView v[2];
    int shared_a;  // NV == 2: both accumulators read the A tiles of view 0 (one clean view, two weight sets: inference)
    int cache;  // operand cache of the LAST view (the student view): 0 off, 1 = its hi A tiles are also TMA-stored from shared memory
                //   to xs_hi [N][768] for the backward kernel, 2 = hi and lo
    int dbg;    // DAD_TC_DBG timing experiments (never set in production): 1 = one product instead of three
    int64_t N;
    int T;
    int num_tiles;
};

// Sum over the 32 lanes of each of the 16 per-lane values: on return lane L holds sum_lanes v[(L >> 1) & 15] (both lanes of a pair
// hold the same column; fixed butterfly order).  Stage with partner lane^off (off = 16, 8, 4, 2): the lane whose `off` bit is set
// keeps the upper half of its live values and sends the lower half; the last stage (off = 1) adds the two halves of the lane sum.
__device__ __forceinline__ float transpose_reduce16(float (&v)[16], const int lane) {
#pragma unroll
    for (int off = 16; off >= 2; off >>= 1) {
        const int n = off >> 1;  // live values after this stage
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float send = upper ? v[i] : v[i + n];
            const float keep = upper ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ---------------------------------------------------------------------------------------------------------------------
// Converter warp body of the forward kernel.  16 converter warps in two TEAMS of eight that alternate k-blocks (team = cw >> 3
// takes the k-blocks of its parity; KB = 24 is even, so the parity of a k-block is the parity of its pipeline slot as well).
// Inside a team, thread (w8 = cw & 7, lane): float4 column c4 = lane & 7 of the 32-wide k-block and the four tile rows
// w8*16 + g + 4 i (g = lane >> 3, i = 0..3):
//   - one warp-wide LDG.128 reads four fully used 128-byte row segments; the warp's stores of one i cover four CONSECUTIVE
//     64-byte tile rows (256 contiguous, swizzle-permuted bytes): two shared-memory wavefronts per 8-byte store, no conflicts;
//   - addresses advance incrementally (one 64-bit pointer per tile, compile-time offsets inside a group of NSET k-blocks); the
//     first version re-derived tile / row / bounds from the flat iteration index and spent ~45 of its ~200 instructions per
//     k-block on that (ncu source view), with two converter warps per scheduler at 55 % issue utilisation -- hence two teams:
//     twice the warps in flight, the same 16 elements per thread and k-block to amortise the wait / fence / arrive over;
//   - NSET register sets rotate over the team's k-blocks (12 per tile, a multiple of NSET): the loads of its next NSET-1
//     k-blocks, which may belong to the CTA's next tile, are in flight while one is converted;
//   - rows beyond N are loaded as zeros and otherwise treated like any row (their accumulator rows are never read, the operand-
//     cache store clips them); rows inside the temporal mask span of the strong view are multiplied by 0 instead of TC_XS.
// NV == 1: MODE 0 clean, 1 weak, 2 strong (I/utils.py:328-375).  NV == 2: view 0 = weak, view 1 = strong of the same x; in the
// Philox mode both views' normals come from the same Philox block (dad_common.cuh).  PHILOX: in-kernel noise, else supplied tensors.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pin_reg(uint32_t v) {  // opaque to the optimiser: computed once, kept in a register (no rematerialisation)
    asm volatile("" : "+r"(v));
    return v;
}
template <int NV, bool CTA2, int MODE, bool PHILOX, int NSET>
__device__ __forceinline__ void fwd_convert(const FwdParams& p, uint8_t* smem, uint64_t* full, uint64_t* afull, uint64_t* empty, const int cw,
                                            const int lane, const uint32_t rank, const int unit0, const int ustride, const int n_units) {
    using Cfg = FwdCfg<NV, CTA2>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int KT = KB / 2;          // k-blocks of one team per tile
    constexpr int GROUPS = KT / NSET;
    static_assert(KB % 2 == 0 && KT % NSET == 0, "a group of NSET team k-blocks must not straddle a tile");
    // NV == 2 with MODE == 0: ONE clean view feeds both accumulators (inference: student and teacher weights on the same x, dad_predict)
    constexpr bool SHARED = NV == 2 && MODE == 0;
    constexpr bool WEAK = (NV == 2 && !SHARED) || MODE == 1, STRONG = (NV == 2 && !SHARED) || MODE == 2;
    const int team = cw >> 3, w8 = cw & 7;
    const int c4 = lane & 7, g = lane >> 3;
    const int rbase = w8 * 16 + g;  // tile rows rbase + 4 i
    const uint32_t a_sw = ((uint32_t)rbase >> 1) & 3u, c16 = (uint32_t)c4 >> 1;
    const uint32_t sbase = tc::smem_u32(smem) + Cfg::OFF_A + (uint32_t)rbase * 64u + (uint32_t)(c4 & 1) * 8u;
    // ((rbase + 4 i) >> 1) & 3 = a_sw ^ ((i & 1) << 1): two distinct swizzled offsets, + i * 256
    const uint32_t so_even = pin_reg(sbase + ((c16 ^ a_sw) << 4)), so_odd = pin_reg(sbase + 256u + ((c16 ^ a_sw ^ 2u) << 4));
    const AugParams& apw = p.v[0].ap;        // weak view (NV == 2) or the single view
    const AugParams& aps = p.v[NV - 1].ap;   // strong view (NV == 2) or the single view
    const uint32_t full_addr = pin_reg(CTA2 ? tc::mapa(tc::smem_u32(full), 0) : tc::smem_u32(full));  // the (leader CTA's) full[]
    const uint32_t empty_addr = pin_reg(tc::smem_u32(empty)), afull_addr = pin_reg(tc::smem_u32(afull));
    const unsigned long long seed = ((NV == 2 || MODE != 0) && PHILOX) ? aug_seed(aps) : 0ull;
    const float rk_w = NOISE_UNIT_K * apw.sigma * apw.sigma, rk_s = NOISE_UNIT_K * aps.sigma * aps.sigma;
    const uint32_t blk = (uint32_t)(aps.stream >> 1);
    const bool odd_stream = (aps.stream & 1ull) != 0;  // NV == 1: which word pair of the block this view uses
    const bool has_keep = STRONG && aps.feat_keep != nullptr;
    const bool has_span = STRONG && aps.t_start != nullptr && aps.mask_len > 0;
    const bool cache = p.cache != 0;
    if (unit0 >= n_units) return;

    // a tile's state: pointer to (row rbase, column team*32 + c4*4) of the tile and the in-range bits of the four rows
    auto tile_of = [&](int u) { return CTA2 ? u * 2 + (int)rank : u; };
    auto tile_ptr = [&](int tile) { return p.x + ((int64_t)tile * BM + rbase) * DAD_D + team * BK + c4 * 4; };
    auto tile_valid = [&](int tile) -> uint32_t {
        const int64_t r0 = (int64_t)tile * BM + rbase;
        return (r0 < p.N ? 1u : 0u) | (r0 + 4 < p.N ? 2u : 0u) | (r0 + 8 < p.N ? 4u : 0u) | (r0 + 12 < p.N ? 8u : 0u);
    };
    auto load = [&](float4 (&dst)[4], const float* src, uint32_t valid) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            dst[i] = ((valid >> i) & 1u) ? __ldg(reinterpret_cast<const float4*>(src + (4 * i) * DAD_D)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto store_row = [&](uint32_t a_hi, const float4& v, float sc) {
        uint32_t h01, l01, h23, l23;
        tc::split2_f16(v.x * sc, v.y * sc, h01, l01);
        tc::split2_f16(v.z * sc, v.w * sc, h23, l23);
        tc::sts64(a_hi, h01, h23);
        tc::sts64(a_hi + A_BYTES, l01, l23);
    };

    uint32_t s = (uint32_t)team % STAGES, ph = 0;  // pipeline stage / phase of the team's k-block being converted (advance by 2)
    int u = unit0;
    int tile = tile_of(u);
    const float* ptr_cur = tile_ptr(tile);
    uint32_t v_cur = tile_valid(tile);
    float4 buf[NSET][4];
#pragma unroll
    for (int j = 0; j < NSET - 1; ++j) load(buf[j], ptr_cur + j * 2 * BK, v_cur);
    for (; u < n_units; u += ustride) {
        tile = tile_of(u);
        const bool more = u + ustride < n_units;
        const int tile_n = tile_of(u + ustride);
        const float* ptr_next = more ? tile_ptr(tile_n) : ptr_cur;
        const uint32_t v_next = more ? tile_valid(tile_n) : 0u;
        const int64_t n0 = (int64_t)tile * BM + rbase;  // local frame index of row 0 (row i: + 4 i)
        uint32_t zspan = 0;  // bit i: row i lies inside the temporal mask span (strong view)
        if (has_span) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t n = n0 + 4 * i;
                if (n < p.N) {
                    const uint32_t bb = (uint32_t)n / (uint32_t)p.T, t = (uint32_t)n - bb * (uint32_t)p.T;  // N < 2^31 on this path (launch check)
                    const int st0 = __ldg(aps.t_start + bb);
                    if ((int)t >= st0 && (int)t < st0 + aps.mask_len) zspan |= 1u << i;
                }
            }
        }
        const unsigned long long gn0 = (unsigned long long)n0 + (unsigned long long)aps.frame_offset;
        const float* nzw = (!PHILOX && NV == 2 && !SHARED) ? apw.noise + n0 * DAD_D + team * BK + c4 * 4 : nullptr;
        const float* nzs = (!PHILOX && (WEAK || STRONG)) ? aps.noise + n0 * DAD_D + team * BK + c4 * 4 : nullptr;
#pragma unroll 1
        for (int grp = 0; grp < GROUPS; ++grp) {
            const bool last = grp == GROUPS - 1;
#pragma unroll
            for (int j = 0; j < NSET; ++j) {
                const int k = grp * NSET + j;   // the team's k-th k-block of this tile: kb = 2 k + team
                // ---- prefetch the team's k-block k + NSET - 1 (of the next tile in the last group)
                if (j == 0 || !last) load(buf[(j + NSET - 1) % NSET], ptr_cur + (k + NSET - 1) * 2 * BK, v_cur);
                else load(buf[(j + NSET - 1) % NSET], ptr_next + (j - 1) * 2 * BK, v_next);
                // ---- convert k-block kb
                tc::mbar_wait_addr(empty_addr + s * 8u, ph ^ 1u);
                const uint32_t stage_off = s * Cfg::STAGE_BYTES;
                float4 keep = make_float4(1.f, 1.f, 1.f, 1.f);
                if (has_keep) keep = __ldg(reinterpret_cast<const float4*>(aps.feat_keep + (2 * k + team) * BK + c4 * 4));
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 vw = buf[j][i], vs = buf[j][i];
                    if (WEAK || STRONG) {
                        if (PHILOX) {
                            const uint4 w = noise_words(seed, blk, gn0 + 4 * i, (uint32_t)((2 * k + team) * (BK / 4) + c4));
                            NoiseBM4 z;
                            if (NV == 2 && !SHARED) {
                                bm4_from_words(w.x, w.y, rk_w, z);
                                vw = bm4_add(z, vw);
                                bm4_from_words(w.z, w.w, rk_s, z);
                                vs = bm4_add(z, vs);
                            } else {
                                bm4_from_words(odd_stream ? w.z : w.x, odd_stream ? w.w : w.y, rk_s, z);
                                vs = bm4_add(z, vs);
                            }
                        } else {
                            // parity mode: supplied unit normals, out = x + (n * sigma) with two fp32 roundings (bit-exact vs I/utils.py:330,338)
                            const bool inr = (v_cur >> i) & 1u;
                            const int64_t e = (int64_t)(4 * i) * DAD_D + k * 2 * BK;
                            if (NV == 2 && !SHARED) {
                                const float4 nw = inr ? __ldg(reinterpret_cast<const float4*>(nzw + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
                                vw.x = aug_add_noise(vw.x, nw.x, apw.sigma); vw.y = aug_add_noise(vw.y, nw.y, apw.sigma);
                                vw.z = aug_add_noise(vw.z, nw.z, apw.sigma); vw.w = aug_add_noise(vw.w, nw.w, apw.sigma);
                            }
                            const float4 ns = inr ? __ldg(reinterpret_cast<const float4*>(nzs + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            vs.x = aug_add_noise(vs.x, ns.x, aps.sigma); vs.y = aug_add_noise(vs.y, ns.y, aps.sigma);
                            vs.z = aug_add_noise(vs.z, ns.z, aps.sigma); vs.w = aug_add_noise(vs.w, ns.w, aps.sigma);
                        }
                        if (has_keep) {
                            vs.x = __fmul_rn(vs.x, keep.x); vs.y = __fmul_rn(vs.y, keep.y); vs.z = __fmul_rn(vs.z, keep.z); vs.w = __fmul_rn(vs.w, keep.w);
                        }
                    }
                    const uint32_t a_row = stage_off + ((i & 1) ? so_odd : so_even) + (uint32_t)(i >> 1) * 512u;  // rows + 4 i: + i * 256 bytes
                    const float sc_s = (STRONG && ((zspan >> i) & 1u)) ? 0.f : TC_XS;
                    if (NV == 2 && !SHARED) {
                        store_row(a_row, vw, TC_XS);
                        store_row(a_row + 2 * A_BYTES, vs, sc_s);
                    } else {
                        store_row(a_row, vs, sc_s);
                    }
                }
                tc::fence_proxy_async_smem();  // generic-proxy stores -> visible to tcgen05.mma / TMA (async proxy)
                __syncwarp();
                if (lane == 0) {
                    if (CTA2) tc::mbar_arrive_cluster(full_addr + s * 8u);
                    else tc::mbar_arrive_addr(full_addr + s * 8u);
                    if (cache) tc::mbar_arrive_addr(afull_addr + s * 8u);
                }
                s += 2;
                if (s >= (uint32_t)STAGES) { s -= STAGES; ph ^= 1u; }
            }
        }
        ptr_cur = ptr_next;
        v_cur = v_next;
    }
}

template <int NV, bool CTA2>
__global__ void __launch_bounds__(FWD_THREADS, 1)
k_enc_fwd_tc(const __grid_constant__ CUtensorMap tm_w0hi, const __grid_constant__ CUtensorMap tm_w0lo, const __grid_constant__ CUtensorMap tm_w1hi,
             const __grid_constant__ CUtensorMap tm_w1lo, const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo,
             FwdParams p) {
    using Cfg = FwdCfg<NV, CTA2>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
    uint64_t* full = bars;                          // [STAGES] leader CTA: 1 (TMA expect_tx) + 8 converter warps per CTA of the pair
    uint64_t* afull = bars + Cfg::MAX_STAGES;       // [STAGES] 8 local converter warps: the A tiles of THIS CTA are written (operand-cache store)
    uint64_t* empty = bars + 2 * Cfg::MAX_STAGES;   // [STAGES] 1 (tcgen05.commit, multicast to both CTAs of a pair) + 1 (operand-cache store warp)
    uint64_t* tfull = bars + 3 * Cfg::MAX_STAGES;   // [2]      1 (tcgen05.commit)
    uint64_t* tempty = tfull + 2;                   // [2]      leader CTA: 8 epilogue warps per CTA of the pair
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
    float* s_bias = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CTA2 ? tc::cluster_ctarank() : 0u;
    const int unit0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // tile (pair) index this CTA (pair) starts with
    const int ustride = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int n_units = CTA2 ? (p.num_tiles + 1) / 2 : p.num_tiles;         // a pair works on tiles 2u (leader) and 2u+1 (peer)
    constexpr uint32_t NCTA = CTA2 ? 2u : 1u;

    if (threadIdx.x < NV * DAD_H) s_bias[threadIdx.x] = p.v[threadIdx.x / DAD_H].bias[threadIdx.x % DAD_H];
    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tm_w0hi);
        tc::prefetch_tmap(&tm_w0lo);
        if (NV == 2) { tc::prefetch_tmap(&tm_w1hi); tc::prefetch_tmap(&tm_w1lo); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full[s], 1 + (FWD_CONV_WARPS / 2) * NCTA);  // one TEAM of eight converter warps (per CTA) fills a stage
            tc::mbar_init(&afull[s], FWD_CONV_WARPS / 2);
            tc::mbar_init(&empty[s], 1 + (p.cache ? 1 : 0));
        }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull[a], 1); tc::mbar_init(&tempty[a], N_EPI_WARPS * NCTA); }
        tc::fence_barrier_init();
    }
    if (CTA2) tc::cluster_sync();  // the peer's barriers exist before anyone arrives on them remotely
    else __syncthreads();
    if (warp == 2) {
        if (CTA2) { tc::tmem_alloc_pair(tmem_ptr, TMEM_COLS); tc::tmem_relinquish_pair(); }
        else { tc::tmem_alloc(tmem_ptr, TMEM_COLS); tc::tmem_relinquish(); }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // register budget follows the roles (setmaxnreg, one per role branch).  896 threads launch with 72 registers each and
    // setmaxnreg.inc can only draw on what the CTA's own warps released (NOT on the SM's unallocated registers), so the roles must
    // fit in 896 x 72 = 64 512:  128 x 32 (control warps 0-3) + 512 x 88 (converters) + 256 x 56 (epilogue) = 63 488

    if (warp == 0) {
        // ===================== TMA producer (weights; in a pair every CTA loads ITS half of the hidden units) =====================
        tc::reg_dec<32>();
        if (lane == 0) {
            const uint32_t full_remote = CTA2 ? tc::mapa(tc::smem_u32(full), 0) : 0u;
            uint32_t it = 0;
            for (int u = unit0; u < n_units; u += ustride)
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    tc::mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* w0 = smem + s * Cfg::STAGE_BYTES + Cfg::OFF_W;
                    if (!CTA2) {
                        tc::mbar_arrive_expect_tx(&full[s], NV * 2 * Cfg::W_BYTES);
                        tc::tma_load_2d(w0, &tm_w0hi, &full[s], kb * BK, 0);
                        tc::tma_load_2d(w0 + Cfg::W_BYTES, &tm_w0lo, &full[s], kb * BK, 0);
                        if (NV == 2) {
                            tc::tma_load_2d(w0 + 2 * Cfg::W_BYTES, &tm_w1hi, &full[s], kb * BK, 0);
                            tc::tma_load_2d(w0 + 3 * Cfg::W_BYTES, &tm_w1lo, &full[s], kb * BK, 0);
                        }
                    } else {
                        // the leader arms its barrier with the bytes of BOTH CTAs; the peer's copies complete on the leader's barrier too
                        if (rank == 0) tc::mbar_arrive_expect_tx(&full[s], 2 * NV * 2 * Cfg::W_BYTES);
                        const uint32_t fb = full_remote + s * 8u;
                        const int r0 = (int)rank * (int)Cfg::W_ROWS;
                        tc::tma_load_2d_pair(w0, &tm_w0hi, fb, kb * BK, r0);
                        tc::tma_load_2d_pair(w0 + Cfg::W_BYTES, &tm_w0lo, fb, kb * BK, r0);
                        if (NV == 2) {
                            tc::tma_load_2d_pair(w0 + 2 * Cfg::W_BYTES, &tm_w1hi, fb, kb * BK, r0);
                            tc::tma_load_2d_pair(w0 + 3 * Cfg::W_BYTES, &tm_w1lo, fb, kb * BK, r0);
                        }
                    }
                }
            // tail: every commit (possibly multicast from the leader) to this CTA's empty[] has landed before the CTA may exit
            for (int k = 0; k < STAGES; ++k, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                tc::mbar_wait(&empty[s], ph ^ 1);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (the leader CTA's for a pair) =====================
        tc::reg_dec<32>();
        if (!CTA2 || rank == 0) {
            constexpr uint32_t idesc = tc::make_idesc_f16(CTA2 ? 256 : 128, DAD_H, 0, 0);
            uint32_t it = 0, ti = 0;
            for (int u = unit0; u < n_units; u += ustride, ++ti) {
                const uint32_t acc = NV == 1 ? (ti & 1) : 0u, tph = NV == 1 ? ((ti >> 1) & 1) : (ti & 1);
                if (CTA2) tc::mbar_wait_cluster(&tempty[acc], tph ^ 1);
                else tc::mbar_wait(&tempty[acc], tph ^ 1);
                tc::tcgen05_fence_after();
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    if (CTA2) tc::mbar_wait_cluster(&full[s], ph);
                    else tc::mbar_wait(&full[s], ph);
                    tc::tcgen05_fence_after();
                    if (lane == 0) {
                        const uint32_t base = tc::smem_u32(smem + s * Cfg::STAGE_BYTES);
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            const uint32_t d_tmem = tmem_base + (NV == 1 ? acc : (uint32_t)v) * DAD_H;
#pragma unroll
                            for (int prod = 0; prod < 3; ++prod) {
                                if (TC_DBG(p, 1) && prod > 0) break;
                                const uint32_t a_off = Cfg::OFF_A + (p.shared_a ? 0 : v) * 2 * A_BYTES + ((prod == 2) ? A_BYTES : 0u);  // hi*hi, hi*lo, lo*hi
                                const uint32_t b_off = Cfg::OFF_W + v * 2 * Cfg::W_BYTES + ((prod == 1) ? Cfg::W_BYTES : 0u);
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k) {
                                    const uint64_t ad = tc::make_smem_desc_sw64(base + a_off + k * 32);
                                    const uint64_t bd = tc::make_smem_desc_sw64(base + b_off + k * 32);
                                    if (CTA2) tc::umma_f16_pair(d_tmem, ad, bd, idesc, (kb | prod | k) != 0);
                                    else tc::umma_f16(d_tmem, ad, bd, idesc, (kb | prod | k) != 0);
                                }
                            }
                        }
                        if (CTA2) {
                            tc::umma_commit_pair(&empty[s], 3);                     // smem slot free in BOTH CTAs once these MMAs retire
                            if (kb == KB - 1) tc::umma_commit_pair(&tfull[acc], 3);  // accumulators complete (each CTA holds its 128 rows)
                        } else {
                            tc::umma_commit(&empty[s]);
                            if (kb == KB - 1) tc::umma_commit(&tfull[acc]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= CONV_WARP0 && warp < CONV_WARP0 + FWD_CONV_WARPS) {
        // ===================== converters: fp32 x (+ augmentation) -> fp16 hi/lo A operand(s) =====================
        tc::reg_inc<88>();
        const int cw = warp - CONV_WARP0;
        const bool philox = p.v[NV - 1].ap.noise_src == 1;
        if (NV == 2 && p.shared_a) {
            fwd_convert<2, CTA2, 0, false, 3>(p, smem, full, afull, empty, cw, lane, rank, unit0, ustride, n_units);
        } else if (NV == 2) {
            if (philox) fwd_convert<2, CTA2, 2, true, 2>(p, smem, full, afull, empty, cw, lane, rank, unit0, ustride, n_units);
            else fwd_convert<2, CTA2, 2, false, 2>(p, smem, full, afull, empty, cw, lane, rank, unit0, ustride, n_units);
        } else {
            const int variant = p.v[0].ap.mode == 0 ? 0 : (p.v[0].ap.mode * 2 - 1 + (philox ? 1 : 0));  // 0 | 1,2 weak | 3,4 strong (supplied, philox)
            switch (variant) {
                case 0: fwd_convert<1, CTA2, 0, false, 3>(p, smem, full, afull, empty, cw, lane, rank, unit0, ustride, n_units); break;
                case 1: fwd_convert<1, CTA2, 1, false, 2>(p, smem, full, afull, empty, cw, lane, rank, unit0, ustride, n_units); break;
                case 2: fwd_convert<1, CTA2, 1, true, 2>(p, smem, full, afull, empty, cw, lane, rank, unit0, ustride, n_units); break;
                case 3: fwd_convert<1, CTA2, 2, false, 2>(p, smem, full, afull, empty, cw, lane, rank, unit0, ustride, n_units); break;
                default: fwd_convert<1, CTA2, 2, true, 2>(p, smem, full, afull, empty, cw, lane, rank, unit0, ustride, n_units); break;
            }
        }
    } else if (warp >= FWD_EPI_WARP0) {
        // ===================== epilogue: TMEM -> bias + ReLU + mask -> ReLU bits + pooled partial sums =====================
        // Lane = frame row (tcgen05.ld 32x32b), 32 hidden units per chunk.  The per-utterance column sums over the 32 rows are a
        // register transpose-reduce: 31 shuffles + 31 adds per 32 x 32 block leave lane j with the sum of column j -- no shared
        // memory, no serial chain.  Eight warps: warp % 4 = TMEM lane quadrant; NV == 2: warps 20-23 drain the weak / teacher
        // accumulator, 24-27 the strong / student one (both belong to the same tile and are NOT double-buffered: the drain is the
        // exposed part of a tile, so it is spread over twice the warps); NV == 1: the two warps of a quadrant take four chunks each.
        tc::reg_dec<56>();
        const int ew = warp - FWD_EPI_WARP0;
        const int q = ew & 3, sel = ew >> 2;
        const int view = NV == 2 ? sel : 0;
        const int c_begin = NV == 2 ? 0 : sel * 4, c_end = NV == 2 ? 8 : sel * 4 + 4;
        const FwdView& V = p.v[view];
        const uint32_t wstage = tc::smem_u32(smem + Cfg::OFF_SCRATCH) + (uint32_t)(ew * 32 * 9 + lane * 9) * 4u;  // [32 rows][8 words + 1 pad]
        const uint32_t bias_a = tc::smem_u32(s_bias + view * DAD_H);
        const uint32_t tempty_remote = CTA2 ? tc::mapa(tc::smem_u32(tempty), 0) : 0u;
        uint32_t ti = 0;
        for (int u = unit0; u < n_units; u += ustride, ++ti) {
            const int tile = CTA2 ? u * 2 + (int)rank : u;
            const bool live = tile < p.num_tiles;  // the peer's tile of the last pair may not exist
            const uint32_t acc = NV == 1 ? (ti & 1) : 0u, tph = NV == 1 ? ((ti >> 1) & 1) : (ti & 1);
            const int64_t n0 = (int64_t)tile * BM + q * 32;
            const int64_t n = n0 + lane;
            const bool in = n < p.N;
            const bool valid = in && !(p.pad && p.pad[n]);
            const int64_t b_first = n0 / p.T;
            const int64_t split64 = (b_first + 1) * p.T - n0;  // rows [0,split) belong to utterance b_first, the rest to b_first+1 (T >= 32)
            const int split = split64 > 32 ? 32 : (int)split64;
            const bool first = lane < split;
            tc::mbar_wait(&tfull[acc], tph);
            tc::tcgen05_fence_after();
            const int64_t gidx = (int64_t)tile * 4 + q;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (NV == 1 ? acc : (uint32_t)view) * DAD_H;
#pragma unroll 1  // keep the body small: the unrolled version was instruction-fetch bound
            for (int c = c_begin; c < c_end; ++c) {
                uint32_t w = 0;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {  // 16 hidden units at a time: r[16] + v[16] keep the epilogue warps at 64 registers
                    const int col = c * 32 + hf * 16;
                    uint32_t r[16];
                    tc::tmem_ld_32x16(taddr + col, r);
                    tc::tmem_ld_wait();
                    float v[16];
#pragma unroll
                    for (int j4 = 0; j4 < 16; j4 += 4) {
                        const uint4 bs = tc::lds128(bias_a + (uint32_t)(col + j4) * 4u);
                        const float bj[4] = {__uint_as_float(bs.x), __uint_as_float(bs.y), __uint_as_float(bs.z), __uint_as_float(bs.w)};
#pragma unroll
                        for (int uu = 0; uu < 4; ++uu) {
                            const float pre = fmaf(__uint_as_float(r[j4 + uu]), TC_INV_XW, bj[uu]);  // exact power-of-two unscale, then + bias
                            const bool pos = pre > 0.f && valid;  // padded frames carry no ReLU bits (nothing of them reaches the backward)
                            w |= (pos ? 1u : 0u) << (hf * 16 + j4 + uu);
                            v[j4 + uu] = (pos && first) ? pre : 0.f;
                        }
                    }
                    const float s0 = transpose_reduce16(v, lane);
                    float s1 = 0.f;
                    if (split < 32) {  // the 32 rows straddle two utterances: a second masked reduction (warp-uniform branch)
#pragma unroll
                        for (int j4 = 0; j4 < 16; j4 += 4) {
                            const uint4 bs = tc::lds128(bias_a + (uint32_t)(col + j4) * 4u);
                            const float bj[4] = {__uint_as_float(bs.x), __uint_as_float(bs.y), __uint_as_float(bs.z), __uint_as_float(bs.w)};
#pragma unroll
                            for (int uu = 0; uu < 4; ++uu) {
                                const float pre = fmaf(__uint_as_float(r[j4 + uu]), TC_INV_XW, bj[uu]);
                                v[j4 + uu] = (pre > 0.f && valid && !first) ? pre : 0.f;
                            }
                        }
                        s1 = transpose_reduce16(v, lane);
                    }
                    if (live && (lane & 1) == 0) {  // lane L holds column (L >> 1) of this half
                        V.part[(gidx * 2 + 0) * DAD_H + col + (lane >> 1)] = s0;
                        V.part[(gidx * 2 + 1) * DAD_H + col + (lane >> 1)] = s1;
                    }
                }
                tc::sts32(wstage + (uint32_t)(c - c_begin) * 4u, w);
            }
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTA2) tc::mbar_arrive_cluster(tempty_remote + acc * 8u);
                else tc::mbar_arrive(&tempty[acc]);
            }
            if (V.bits && in) {
                uint4* dst = reinterpret_cast<uint4*>(V.bits + n * DAD_HW);
                if (NV == 2) {
                    dst[0] = make_uint4(tc::lds32(wstage), tc::lds32(wstage + 4), tc::lds32(wstage + 8), tc::lds32(wstage + 12));
                    dst[1] = make_uint4(tc::lds32(wstage + 16), tc::lds32(wstage + 20), tc::lds32(wstage + 24), tc::lds32(wstage + 28));
                } else {
                    dst[sel] = make_uint4(tc::lds32(wstage), tc::lds32(wstage + 4), tc::lds32(wstage + 8), tc::lds32(wstage + 12));
                }
            }
            __syncwarp();
        }
    } else if (warp == 3) {
        // ===================== operand-cache store: the student view's A tiles, shared -> global by TMA =====================
        // The hi A tile of a stage ([128 frames][32 k] fp16, SWIZZLE_64B) is exactly one TMA box of xs_hi [N][768] (and the lo tile
        // of xs_lo when the 3-product backward is selected); the slot is handed back (second arrival on empty[s]) once the bulk
        // store has finished READING shared memory.  Frames beyond N are clipped by the tensor map.
        tc::reg_dec<32>();
        if (p.cache && lane == 0) {
            tc::prefetch_tmap(&tm_xhi);
            if (p.cache > 1) tc::prefetch_tmap(&tm_xlo);
            uint32_t it = 0;
            for (int u = unit0; u < n_units; u += ustride) {
                const int tile = CTA2 ? u * 2 + (int)rank : u;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    tc::mbar_wait(&afull[s], ph);
                    const uint8_t* a = smem + s * Cfg::STAGE_BYTES + Cfg::OFF_A + (NV - 1) * 2 * A_BYTES;
                    if (tile < p.num_tiles) {
                        tc::tma_store_2d(&tm_xhi, a, kb * BK, tile * BM);
                        if (p.cache > 1) tc::tma_store_2d(&tm_xlo, a + A_BYTES, kb * BK, tile * BM);
                        tc::bulk_commit();
                        tc::bulk_wait_read0();
                    }
                    tc::mbar_arrive(&empty[s]);
                }
            }
            tc::bulk_wait0();
        }
    } else {
        tc::reg_dec<32>();  // warp 2 (TMEM allocator)
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (CTA2) tc::cluster_sync();  // nobody frees TMEM / leaves while the pair's MMAs, remote arrivals or multicast commits are in flight
    if (warp == 2) {
        tc::tcgen05_fence_after();
        if (CTA2) tc::tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// emb[b,:] = (sum of the 32-frame partials of utterance b, in frame order) / max(len_b, 1);  grid = (B, views), 256 threads
struct PoolArgs {
    const float* part[2];
    float* emb[2];
    float* lens_out;  // written by view 0
};
__global__ void __launch_bounds__(DAD_H) k_pool_finalize(PoolArgs a, const uint8_t* __restrict__ pad, int T) {
    __shared__ float s_w[DAD_H / 32];
    __shared__ float s_len;
    const int b = blockIdx.x, h = threadIdx.x;
    const float* __restrict__ part = a.part[blockIdx.y];
    float* __restrict__ emb = a.emb[blockIdx.y];
    // partial sums first (their loads do not depend on the length): ascending 32-frame groups, eight loads in flight.  A group
    // that straddles two utterances keeps the earlier one in slot 0 and the later one in slot 1 (T >= 32: never more than two),
    // so only this utterance's FIRST group can be a slot-1 read.
    const int64_t r0 = (int64_t)b * T, r1 = r0 + T - 1;
    const int64_t g0 = r0 >> 5, g1 = r1 >> 5;
    float s = __ldg(part + (g0 * 2 + (g0 * 32 == r0 ? 0 : 1)) * DAD_H + h);
    for (int64_t g = g0 + 1; g <= g1; g += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = g + u <= g1 ? __ldg(part + (g + u) * 2 * DAD_H + h) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (g + u <= g1) s += v[u];
    }
    if (pad) {
        float c = 0.f;
        for (int t = h; t < T; t += DAD_H) c += pad[(int64_t)b * T + t] ? 0.f : 1.f;
        c = warp_sum(c);
        if ((h & 31) == 0) s_w[h >> 5] = c;
        __syncthreads();
        if (h == 0) {
            float tot = 0.f;
            for (int w = 0; w < DAD_H / 32; ++w) tot += s_w[w];
            s_len = fmaxf(tot, 1.0f);
        }
    } else if (h == 0) {
        s_len = (float)T;
    }
    __syncthreads();
    emb[(int64_t)b * DAD_H + h] = s / s_len;
    if (h == 0 && blockIdx.y == 0 && a.lens_out) a.lens_out[b] = s_len;
}

// CTA pairs (cta_group::2): process-wide switch, DAD_TC_CTA2=0|1|2 in the environment or dad_tc_set_cta_pair(): 0 never, 1 always,
// 2 (default) = where it pays: the clean / inference launches are bound by shared-memory operand traffic of the MMAs and gain from
// sharing the W k-blocks across the pair; the noisy launches are bound by their converters (Philox + Box-Muller), and coupling two
// SMs to one barrier only adds the slower SM's jitter to both (measured: DESIGN.md 5.1).
int g_cta2 = -1;
int cta2_mode() {
    if (g_cta2 < 0) {
        const char* e = getenv("DAD_TC_CTA2");
        g_cta2 = e ? atoi(e) : 2;
        if (g_cta2 < 0 || g_cta2 > 2) g_cta2 = 2;
    }
    return g_cta2;
}

// fp16 [rows][768] row-major, box = [box_rows][32 cols] (64 bytes inner), SWIZZLE_64B
int make_tmap_f16_sw64(CUtensorMap* m, const void* base, uint64_t rows, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return dad_set_error(DAD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {DAD_D, rows};
    cuuint64_t strides[1] = {DAD_D * 2};
    cuuint32_t box[2] = {BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return dad_set_error(DAD_ERR_CUDA, "cuTensorMapEncodeTiled (SWIZZLE_64B) failed (%d)", (int)r);
    return DAD_OK;
}

template <int NV, bool CTA2>
int launch_fwd_variant(const CUtensorMap (&tw)[4], const CUtensorMap& txh, const CUtensorMap& txl, const FwdParams& p, cudaStream_t st) {
    using Cfg = FwdCfg<NV, CTA2>;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_enc_fwd_tc<NV, CTA2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        if (e != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "encoder_fwd(tc): cannot opt in to %u B shared memory: %s", Cfg::SMEM, cudaGetErrorString(e));
        attr = true;
    }
    const int sms = sm_count();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute at[1];
    if (CTA2) {
        const int pairs = (p.num_tiles + 1) / 2, max_pairs = sms / 2;
        cfg.gridDim = dim3(2 * (pairs < max_pairs ? pairs : max_pairs));
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
    } else {
        cfg.gridDim = dim3(p.num_tiles < sms ? p.num_tiles : sms);
    }
    cfg.blockDim = dim3(FWD_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_enc_fwd_tc<NV, CTA2>, tw[0], tw[1], tw[2], tw[3], txh, txl, p);
    if (e != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "k_enc_fwd_tc<%d,%d> launch: %s", NV, (int)CTA2, cudaGetErrorString(e));
    return DAD_OK;
}


// =====================================================================================================================
// Fused noisy forward, TRANSPOSED over a CTA pair (k_enc_fwd_tp): the training step's weak -> teacher and strong -> student views.
//   pre^T[256 units, 128 frames] = W[256, 768] . aug(x)^T :  the WEIGHTS are the A operand (M = 256 hidden units over the pair: every
//   CTA TMA-loads and owns the W rows of ITS 128 units) and the converted frames are the B operand (N = 128 frames per tile: every
//   CTA converts 64 of them; tcgen05.mma.cta_group::2 shares B across the pair).  Each CTA therefore holds, per view, a
//   [128 units (TMEM lanes)] x [128 frames (columns)] accumulator = 128 columns: two views x TWO tiles fit the 512 TMEM columns, so
//   the epilogue of tile i runs under the MMAs / conversions of tile i+1 (the [frames x units] orientation needs all 512 columns for
//   one tile, which left the drain exposed: ncu showed the converters waiting 41 % of the time).  A stage is 16 KB of x (64 frames x
//   32 k, two views, hi/lo) + 32 KB of W (128 units x 32 k, two nets, hi/lo) = 48 KB: FOUR stages instead of two.
//   The epilogue is also cheaper in this orientation: a lane is a hidden unit, so the pooled sum over frames is a serial per-lane
//   accumulation (no shuffles), the bias is one register, and the ReLU word of (frame, 32 units) is one warp ballot.
//   Same values as k_enc_fwd_tc<2, *>: same Philox counters, same f16x3 products in the same (k-block, product, k-step) order.
// =====================================================================================================================
// NF = frames converted by one CTA per tile: 64 (N = 128 frames per pair tile: two tiles of accumulators in TMEM, four 48 KB stages) or
// 128 (N = 256: every MMA re-reads its 4 KB W slice for 128 instead of 64 cycles of math and a W k-block is TMA-loaded once per 256
// instead of 128 frames -- a third less shared-memory traffic per frame; one tile of accumulators fills the 512 TMEM columns, so the
// MMAs of tile i+1 wait for the (cheap) epilogue of tile i while the converters run ahead into the three 64 KB stages).
template <int NF>
struct TpCfg {
    static_assert(NF == 64 || NF == 128, "frames per CTA and tile");
    static constexpr int CF = NF, TF = 2 * NF;                // frames per CTA / per pair tile (UMMA N)
    static constexpr int HALVES = NF / 64;                     // a converter team walks a k-block in 64-row halves
    static constexpr uint32_t XB = NF * BK * 2;                // 4 / 8 KB: hi or lo of one view of this CTA's frames, [NF][32] fp16, SWIZZLE_64B
    static constexpr uint32_t WB = 128 * BK * 2;               // 8 KB: hi or lo of one net, this CTA's 128 hidden units
    static constexpr uint32_t OFF_X = 0, OFF_W = 4 * XB;       // [view][hi, lo] / [net][hi, lo]
    static constexpr uint32_t STAGE = 4 * XB + 4 * WB;         // 48 / 64 KB
    static constexpr int STAGES = NF == 64 ? 4 : 3;
    static constexpr int NBUF = NF == 64 ? 2 : 1;              // accumulator sets (2 views x TF columns each) in the 512 TMEM columns
    static constexpr int NSET = NF == 64 ? 3 : 4;              // converter landing sets (units of 64 rows x 32 k in flight)
    static constexpr uint32_t OFF_SCRATCH = STAGES * STAGE;    // 8 epilogue warps x 32 ballot words
    static constexpr uint32_t OFF_BARS = OFF_SCRATCH + N_EPI_WARPS * 128;
    static constexpr uint32_t SMEM = OFF_BARS + 256 + 1024;
    static constexpr int MAX_STAGES = 4;
};

// converter warps: two teams of eight alternate k-blocks; thread (w8, lane) owns float4 column c4 = lane & 7 of the 32-wide k-block
// and, per 64-row half hh, the rows 64 hh + w8*8 + (lane >> 3) + 4 i (i = 0, 1) of the CTA's frames.  A "unit" is one half of one
// k-block (64 rows x 32 k); NSET register sets rotate over the team's units, the loads of the next NSET-1 units are in flight.
template <int NF, bool PHILOX, int TEAMS>
__device__ __forceinline__ void tp_convert(const FwdParams& p, uint8_t* smem, uint64_t* full, uint64_t* afull, uint64_t* empty, const int cw,
                                           const int lane, const uint32_t rank, const int unit0, const int ustride, const int n_units) {
    using Cfg = TpCfg<NF>;
    constexpr int H = Cfg::HALVES, STAGES = Cfg::STAGES;
    constexpr int NSET = TEAMS == 3 ? 2 : Cfg::NSET;  // three teams run at 64 registers per thread: two landing sets
    constexpr int KT = KB / TEAMS;             // k-blocks of one team per tile (kb = TEAMS k + team)
    static_assert(KB % TEAMS == 0, "teams take the k-blocks round robin");
    constexpr int UT = KT * H;                 // units of one team per tile
    constexpr int GROUPS = UT / NSET, KPG = NSET / H;  // k-blocks per group
    static_assert(UT % NSET == 0 && NSET % H == 0, "a group of NSET units must not straddle a tile and must hold whole k-blocks");
    const int team = cw >> 3, w8 = cw & 7;
    const int c4 = lane & 7, g = lane >> 3;
    const int rbase = w8 * 8 + g;  // rows 64 hh + rbase + 4 i
    const uint32_t a_sw = ((uint32_t)rbase >> 1) & 3u, c16 = (uint32_t)c4 >> 1;
    const uint32_t sbase = tc::smem_u32(smem) + Cfg::OFF_X + (uint32_t)rbase * 64u + (uint32_t)(c4 & 1) * 8u;
    const uint32_t so0 = pin_reg(sbase + ((c16 ^ a_sw) << 4)), so1 = pin_reg(sbase + 256u + ((c16 ^ a_sw ^ 2u) << 4));
    const AugParams& apw = p.v[0].ap;
    const AugParams& aps = p.v[1].ap;
    const uint32_t full_addr = pin_reg(tc::mapa(tc::smem_u32(full), 0));  // the leader CTA's full[]
    const uint32_t empty_addr = pin_reg(tc::smem_u32(empty)), afull_addr = pin_reg(tc::smem_u32(afull));
    const unsigned long long seed = PHILOX ? aug_seed(aps) : 0ull;
    const float rk_w = NOISE_UNIT_K * apw.sigma * apw.sigma, rk_s = NOISE_UNIT_K * aps.sigma * aps.sigma;
    const uint32_t blk = (uint32_t)(aps.stream >> 1);
    const bool has_keep = aps.feat_keep != nullptr;
    const bool has_span = aps.t_start != nullptr && aps.mask_len > 0;
    const bool cache = p.cache != 0;
    if (unit0 >= n_units) return;

    auto frame0 = [&](int u) { return (int64_t)u * Cfg::TF + (int64_t)rank * Cfg::CF + rbase; };  // this thread's row 0 of tile u
    auto tile_ptr = [&](int u) { return p.x + frame0(u) * DAD_D + team * BK + c4 * 4; };  // + k * TEAMS * BK per team k-block
    auto tile_valid = [&](int u) -> uint32_t {  // bit hh*2 + i: row 64 hh + 4 i of this thread lies inside the batch
        const int64_t r0 = frame0(u);
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 2 * H; ++b) v |= (r0 + 64 * (b >> 1) + 4 * (b & 1) < p.N ? 1u : 0u) << b;
        return v;
    };
    // unit (k, hh) of a tile: + k team-k-blocks of 2*BK floats, + hh * 64 rows
    auto load = [&](float4 (&dst)[2], const float* tile, int k, int hh, uint32_t valid) {
        const float* src = tile + k * TEAMS * BK + (int64_t)hh * 64 * DAD_D;
#pragma unroll
        for (int i = 0; i < 2; ++i)
            dst[i] = ((valid >> (hh * 2 + i)) & 1u) ? __ldg(reinterpret_cast<const float4*>(src + (4 * i) * DAD_D)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto store_row = [&](uint32_t a_hi, const float4& v, float sc) {
        uint32_t h01, l01, h23, l23;
        tc::split2_f16(v.x * sc, v.y * sc, h01, l01);
        tc::split2_f16(v.z * sc, v.w * sc, h23, l23);
        tc::sts64(a_hi, h01, h23);
        tc::sts64(a_hi + Cfg::XB, l01, l23);
    };

    uint32_t it = (uint32_t)team;  // running k-block index of the team's k-block being converted (advance by TEAMS): stage it % STAGES
    int u = unit0;
    const float* ptr_cur = tile_ptr(u);
    uint32_t v_cur = tile_valid(u);
    float4 buf[NSET][2];
#pragma unroll
    for (int j = 0; j < NSET - 1; ++j) load(buf[j], ptr_cur, j / H, j % H, v_cur);
    for (; u < n_units; u += ustride) {
        const bool more = u + ustride < n_units;
        const float* ptr_next = more ? tile_ptr(u + ustride) : ptr_cur;
        const uint32_t v_next = more ? tile_valid(u + ustride) : 0u;
        const int64_t n0 = frame0(u);
        uint32_t zspan = 0;  // bit hh*2 + i: the row lies inside the temporal mask span (strong view)
        if (has_span) {
#pragma unroll
            for (int b = 0; b < 2 * H; ++b) {
                const int64_t n = n0 + 64 * (b >> 1) + 4 * (b & 1);
                if (n < p.N) {
                    const uint32_t bb = (uint32_t)n / (uint32_t)p.T, t = (uint32_t)n - bb * (uint32_t)p.T;
                    const int st0 = __ldg(aps.t_start + bb);
                    if ((int)t >= st0 && (int)t < st0 + aps.mask_len) zspan |= 1u << b;
                }
            }
        }
        const unsigned long long gn0 = (unsigned long long)n0 + (unsigned long long)aps.frame_offset;
        const float* nzw = !PHILOX ? apw.noise + n0 * DAD_D + team * BK + c4 * 4 : nullptr;
        const float* nzs = !PHILOX ? aps.noise + n0 * DAD_D + team * BK + c4 * 4 : nullptr;
#pragma unroll 1
        for (int grp = 0; grp < GROUPS; ++grp) {
            const bool last = grp == GROUPS - 1;
#pragma unroll
            for (int j = 0; j < NSET; ++j) {
                const int k = grp * KPG + j / H;  // the team's k-th k-block of this tile: kb = TEAMS k + team
                const int hh = j % H;
                const int kb = TEAMS * k + team;
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                // ---- prefetch unit q + NSET - 1 (of the next tile in the last group)
                {
                    const int jn = (j + NSET - 1) % NSET;  // slot being refilled = the one consumed in the previous iteration
                    if (j == 0) load(buf[jn], ptr_cur, grp * KPG + (NSET - 1) / H, (NSET - 1) % H, v_cur);
                    else if (!last) load(buf[jn], ptr_cur, (grp + 1) * KPG + (j - 1) / H, (j - 1) % H, v_cur);
                    else load(buf[jn], ptr_next, (j - 1) / H, (j - 1) % H, v_next);
                }
                float4 keep = make_float4(1.f, 1.f, 1.f, 1.f);
                if (has_keep) keep = __ldg(reinterpret_cast<const float4*>(aps.feat_keep + kb * BK + c4 * 4));
                if (hh == 0) tc::mbar_wait_addr(empty_addr + s * 8u, ph ^ 1u);
                const uint32_t stage_off = s * Cfg::STAGE + (uint32_t)hh * 4096u;  // 64 rows x 64 bytes per half
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int rb = hh * 2 + i;
                    float4 vw = buf[j][i], vs = buf[j][i];
                    if (PHILOX) {
                        if (!TC_DBG(p, 4)) {
                            const uint4 w = noise_words(seed, blk, gn0 + 64 * hh + 4 * i, (uint32_t)(kb * (BK / 4) + c4));
                            NoiseBM4 z;
                            bm4_from_words(w.x, w.y, rk_w, z);
                            vw = bm4_add(z, vw);
                            bm4_from_words(w.z, w.w, rk_s, z);
                            vs = bm4_add(z, vs);
                        }
                    } else {
                        // parity mode: supplied unit normals, out = x + (n * sigma) with two fp32 roundings (bit-exact vs I/utils.py:330,338)
                        const bool inr = (v_cur >> rb) & 1u;
                        const int64_t e = (int64_t)(64 * hh + 4 * i) * DAD_D + k * TEAMS * BK;
                        const float4 nw = inr ? __ldg(reinterpret_cast<const float4*>(nzw + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        vw.x = aug_add_noise(vw.x, nw.x, apw.sigma); vw.y = aug_add_noise(vw.y, nw.y, apw.sigma);
                        vw.z = aug_add_noise(vw.z, nw.z, apw.sigma); vw.w = aug_add_noise(vw.w, nw.w, apw.sigma);
                        const float4 ns = inr ? __ldg(reinterpret_cast<const float4*>(nzs + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        vs.x = aug_add_noise(vs.x, ns.x, aps.sigma); vs.y = aug_add_noise(vs.y, ns.y, aps.sigma);
                        vs.z = aug_add_noise(vs.z, ns.z, aps.sigma); vs.w = aug_add_noise(vs.w, ns.w, aps.sigma);
                    }
                    if (has_keep) {
                        vs.x = __fmul_rn(vs.x, keep.x); vs.y = __fmul_rn(vs.y, keep.y); vs.z = __fmul_rn(vs.z, keep.z); vs.w = __fmul_rn(vs.w, keep.w);
                    }
                    const uint32_t a_row = stage_off + (i ? so1 : so0);
                    const float sc_s = ((zspan >> rb) & 1u) ? 0.f : TC_XS;
                    if (!TC_DBG(p, 8) || vw.x == 12345.f) {
                        store_row(a_row, vw, TC_XS);
                        store_row(a_row + 2 * Cfg::XB, vs, sc_s);
                    }
                }
                if (hh == H - 1) {
                    tc::fence_proxy_async_smem();  // generic-proxy stores -> visible to tcgen05.mma / TMA (async proxy)
                    __syncwarp();
                    if (lane == 0) {
                        tc::mbar_arrive_cluster(full_addr + s * 8u);
                        if (cache) tc::mbar_arrive_addr(afull_addr + s * 8u);
                    }
                    it += TEAMS;
                }
            }
        }
        ptr_cur = ptr_next;
        v_cur = v_next;
    }
}

// TEAMS = 2: 896 threads (4 control + 16 converter + 8 epilogue warps, setmaxnreg 32 / 88 / 56).  TEAMS = 3: 1024 threads at 64
// registers (4 control + 24 converter + 4 epilogue warps, each epilogue warp drains both views of its lane quadrant): the converters
// are bound by their own latencies (Philox rounds, MUFU chains) with four warps per scheduler -- six warps give the schedulers more
// to choose from -- and the epilogue, hidden under the next tile, has time to spare.
template <int NF, int TEAMS>
__global__ void __launch_bounds__(128 + TEAMS * 256 + (TEAMS == 3 ? 128 : 256), 1)
k_enc_fwd_tp(const __grid_constant__ CUtensorMap tm_w0hi, const __grid_constant__ CUtensorMap tm_w0lo, const __grid_constant__ CUtensorMap tm_w1hi,
             const __grid_constant__ CUtensorMap tm_w1lo, const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo,
             FwdParams p) {
    using Cfg = TpCfg<NF>;
    constexpr int STAGES = Cfg::STAGES, NBUF = Cfg::NBUF;
    constexpr uint32_t TF = Cfg::TF;
    constexpr int CONV_WARPS = TEAMS * 8, EPI_W0 = CONV_WARP0 + CONV_WARPS, EPI_WARPS = TEAMS == 3 ? 4 : 8, VIEWS_PER_EPI = 8 / EPI_WARPS;
    constexpr bool SETREG = TEAMS == 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
    uint64_t* full = bars;                          // [STAGES] leader CTA: 1 (TMA expect_tx) + 8 converter warps of each CTA
    uint64_t* afull = bars + Cfg::MAX_STAGES;       // [STAGES] 8 local converter warps (operand-cache store)
    uint64_t* empty = bars + 2 * Cfg::MAX_STAGES;   // [STAGES] 1 (tcgen05.commit, multicast to both CTAs) + 1 (operand-cache store warp)
    uint64_t* tfull = bars + 3 * Cfg::MAX_STAGES;   // [2] 1 (tcgen05.commit, multicast)
    uint64_t* tempty = tfull + 2;                   // [2] leader CTA: 8 epilogue warps of each CTA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int unit0 = (int)(blockIdx.x >> 1), ustride = (int)(gridDim.x >> 1);
    const int n_units = (int)ceil_div64(p.N, TF);  // pair tiles

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tm_w0hi); tc::prefetch_tmap(&tm_w0lo); tc::prefetch_tmap(&tm_w1hi); tc::prefetch_tmap(&tm_w1lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full[s], 1 + 8 * 2);  // one team of eight converter warps per CTA fills a stage
            tc::mbar_init(&afull[s], 8);
            tc::mbar_init(&empty[s], 1 + (p.cache ? 1 : 0));
        }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull[a], 1); tc::mbar_init(&tempty[a], EPI_WARPS * 2); }
        tc::fence_barrier_init();
    }
    tc::cluster_sync();  // the peer's barriers exist before anyone arrives on them remotely
    if (warp == 2) { tc::tmem_alloc_pair(tmem_ptr, TMEM_COLS); tc::tmem_relinquish_pair(); }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer: this CTA's 128 hidden units of both nets, hi / lo =====================
        if (SETREG) tc::reg_dec<32>();
        if (lane == 0) {
            const uint32_t full_remote = tc::mapa(tc::smem_u32(full), 0);
            const int r0 = (int)rank * 128;
            uint32_t it = 0;
            for (int u = unit0; u < n_units; u += ustride)
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    tc::mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* w0 = smem + s * Cfg::STAGE + Cfg::OFF_W;
                    if (TC_DBG(p, 2) && it >= (uint32_t)STAGES) {  // timing experiment: W stays what the first fill left
                        if (rank == 0) tc::mbar_arrive(&full[s]);
                        continue;
                    }
                    if (rank == 0) tc::mbar_arrive_expect_tx(&full[s], 2 * 4 * Cfg::WB);  // the bytes of BOTH CTAs land on the leader's barrier
                    const uint32_t fb = full_remote + s * 8u;
                    tc::tma_load_2d_pair(w0, &tm_w0hi, fb, kb * BK, r0);
                    tc::tma_load_2d_pair(w0 + Cfg::WB, &tm_w0lo, fb, kb * BK, r0);
                    tc::tma_load_2d_pair(w0 + 2 * Cfg::WB, &tm_w1hi, fb, kb * BK, r0);
                    tc::tma_load_2d_pair(w0 + 3 * Cfg::WB, &tm_w1lo, fb, kb * BK, r0);
                }
            for (int k = 0; k < STAGES; ++k, ++it) {  // every (multicast) commit to this CTA's empty[] has landed before the CTA may exit
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                tc::mbar_wait(&empty[s], ph ^ 1);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA): A = W rows (M = 256 over the pair), B = converted frames (N = TF) ==========
        if (SETREG) tc::reg_dec<32>();
        if (rank == 0) {
            constexpr uint32_t idesc = tc::make_idesc_f16(256, (int)TF, 0, 0);
            uint32_t it = 0, ti = 0;
            for (int u = unit0; u < n_units; u += ustride, ++ti) {
                const uint32_t ab = NBUF == 2 ? (ti & 1) : 0u, tph = NBUF == 2 ? ((ti >> 1) & 1) : (ti & 1);
                tc::mbar_wait_cluster(&tempty[ab], tph ^ 1);
                tc::tcgen05_fence_after();
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    tc::mbar_wait_cluster(&full[s], ph);
                    tc::tcgen05_fence_after();
                    if (lane == 0) {
                        const uint32_t base = tc::smem_u32(smem + s * Cfg::STAGE);
#pragma unroll
                        for (int v = 0; v < 2; ++v) {
                            const uint32_t d_tmem = tmem_base + ab * 2u * TF + (uint32_t)v * TF;
#pragma unroll
                            for (int prod = 0; prod < 3; ++prod) {  // x_hi*W_hi, x_hi*W_lo, x_lo*W_hi
                                if (TC_DBG(p, 1) && prod > 0) break;
                                const uint32_t w_off = Cfg::OFF_W + v * 2 * Cfg::WB + ((prod == 1) ? Cfg::WB : 0u);
                                const uint32_t x_off = Cfg::OFF_X + v * 2 * Cfg::XB + ((prod == 2) ? Cfg::XB : 0u);
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k) {
                                    const uint64_t ad = tc::make_smem_desc_sw64(base + w_off + k * 32);
                                    const uint64_t bd = tc::make_smem_desc_sw64(base + x_off + k * 32);
                                    tc::umma_f16_pair(d_tmem, ad, bd, idesc, (kb | prod | k) != 0);
                                }
                            }
                        }
                        tc::umma_commit_pair(&empty[s], 3);
                        if (kb == KB - 1) tc::umma_commit_pair(&tfull[ab], 3);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= CONV_WARP0 && warp < EPI_W0) {
        if (SETREG) tc::reg_inc<88>();
        const int cw = warp - CONV_WARP0;
        if (p.v[1].ap.noise_src == 1) tp_convert<NF, true, TEAMS>(p, smem, full, afull, empty, cw, lane, rank, unit0, ustride, n_units);
        else tp_convert<NF, false, TEAMS>(p, smem, full, afull, empty, cw, lane, rank, unit0, ustride, n_units);
    } else if (warp >= EPI_W0) {
        // ===================== epilogue: lane = hidden unit, column = frame =====================
        // warp (q = ew & 3, view = ew >> 2) drains TMEM lanes 32 q .. 32 q + 31 (units rank*128 + 32 q + lane) of its view: per
        // 32-frame group a serial per-lane sum of relu(pre) over the valid frames (slot 0: frames of the group's first utterance,
        // slot 1: of the second, T >= 32) and one ballot per frame = the ReLU word (frame, units 32 (4 rank + q) ..) of the backward.
        if (SETREG) tc::reg_dec<56>();
        const int ew = warp - EPI_W0;
        const int q = ew & 3, view0 = EPI_WARPS == 8 ? (ew >> 2) : 0;
        const int unit = (int)rank * 128 + q * 32 + lane;
        const uint32_t wword = rank * 4u + (uint32_t)q;
        const uint32_t scratch = tc::smem_u32(smem + Cfg::OFF_SCRATCH) + (uint32_t)ew * 128u;
        const uint32_t tempty_remote = tc::mapa(tc::smem_u32(tempty), 0);
        uint32_t ti = 0;
        for (int u = unit0; u < n_units; u += ustride, ++ti) {
            const uint32_t ab = NBUF == 2 ? (ti & 1) : 0u, tph = NBUF == 2 ? ((ti >> 1) & 1) : (ti & 1);
            tc::mbar_wait(&tfull[ab], tph);
            tc::tcgen05_fence_after();
#pragma unroll 1
            for (int vi = 0; vi < VIEWS_PER_EPI; ++vi) {
            const int view = view0 + vi;
            const FwdView& V = p.v[view];
            const float bias = __ldg(V.bias + unit);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 2u * TF + (uint32_t)view * TF;
#pragma unroll 1
            for (int kg = 0; kg < (int)TF / 32; ++kg) {
                const int64_t n0g = (int64_t)u * TF + 32 * kg;
                const int64_t n = n0g + lane;
                const bool in = n < p.N;
                const uint32_t validw = __ballot_sync(0xffffffffu, in && !(p.pad && p.pad[n]));
                const int64_t split64 = (n0g / p.T + 1) * p.T - n0g;  // frames [0, split) of the group belong to its first utterance
                const int split = split64 > 32 ? 32 : (int)split64;
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t r[16];
                    tc::tmem_ld_32x16(taddr + (uint32_t)(32 * kg + 16 * hf), r);
                    tc::tmem_ld_wait();
                    if (split >= 16 * hf + 16 || split <= 16 * hf) {  // the 16 frames belong to one utterance (warp-uniform)
                        float sa = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float pre = fmaf(__uint_as_float(r[j]), TC_INV_XW, bias);  // exact power-of-two unscale, then + bias
                            const bool pos = pre > 0.f && ((validw >> (16 * hf + j)) & 1u);  // padded frames carry no ReLU bits
                            const uint32_t bal = __ballot_sync(0xffffffffu, pos);
                            if (lane == 0) tc::sts32(scratch + (uint32_t)(16 * hf + j) * 4u, bal);
                            sa += pos ? pre : 0.f;
                        }
                        if (split >= 16 * hf + 16) s0 += sa; else s1 += sa;
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float pre = fmaf(__uint_as_float(r[j]), TC_INV_XW, bias);
                            const bool pos = pre > 0.f && ((validw >> (16 * hf + j)) & 1u);
                            const uint32_t bal = __ballot_sync(0xffffffffu, pos);
                            if (lane == 0) tc::sts32(scratch + (uint32_t)(16 * hf + j) * 4u, bal);
                            const float val = pos ? pre : 0.f;
                            if (16 * hf + j < split) s0 += val; else s1 += val;
                        }
                    }
                }
                __syncwarp();
                const uint32_t myw = tc::lds32(scratch + (uint32_t)lane * 4u);  // the ReLU word of frame n0g + lane
                __syncwarp();
                if (V.bits && in) V.bits[n * DAD_HW + wword] = myw;
                const int64_t gidx = n0g >> 5;
                V.part[(gidx * 2 + 0) * DAD_H + unit] = s0;
                V.part[(gidx * 2 + 1) * DAD_H + unit] = s1;
            }
            }
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(tempty_remote + ab * 8u);
        }
    } else if (warp == 3) {
        // ===================== operand-cache store: the strong view's hi (/ lo) tiles of THIS CTA's frames, shared -> global by TMA =====
        if (SETREG) tc::reg_dec<32>();
        if (p.cache && lane == 0) {
            tc::prefetch_tmap(&tm_xhi);
            if (p.cache > 1) tc::prefetch_tmap(&tm_xlo);
            uint32_t it = 0;
            for (int u = unit0; u < n_units; u += ustride) {
                const int64_t f0 = (int64_t)u * TF + (int64_t)rank * Cfg::CF;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    tc::mbar_wait(&afull[s], ph);
                    const uint8_t* a = smem + s * Cfg::STAGE + Cfg::OFF_X + 2 * Cfg::XB;
                    if (f0 < p.N) {
                        tc::tma_store_2d(&tm_xhi, a, kb * BK, (int)f0);
                        if (p.cache > 1) tc::tma_store_2d(&tm_xlo, a + Cfg::XB, kb * BK, (int)f0);
                        tc::bulk_commit();
                        tc::bulk_wait_read0();
                    }
                    tc::mbar_arrive(&empty[s]);
                }
            }
            tc::bulk_wait0();
        }
    } else {
        if (SETREG) tc::reg_dec<32>();  // warp 2 (TMEM allocator)
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::cluster_sync();  // nobody frees TMEM / leaves while the pair's MMAs, remote arrivals or multicast commits are in flight
    if (warp == 2) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
}

// frames per CTA and tile of the transposed kernel: 64; DAD_TC_TP_NF=128 selects the N = 256 form (kept for experiments)
int tp_frames_per_cta(int64_t N) {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("DAD_TC_TP_NF"); forced = e ? atoi(e) : 0; }
    if (forced == 64 || forced == 128) return forced;
    (void)N;
    return 64;  // measured (B=4096, T=200): 2.28-2.36 ms with 64 frames per CTA, 2.44-2.55 ms with 128
}

// converter teams of the transposed kernel: 2; DAD_TC_TP_TEAMS=3 selects the 24-converter-warp form (kept for experiments: measured
// SLOWER, 2.49-2.62 ms against 2.18-2.41 ms at B=4096, T=200 -- more warps in flight do not help a kernel that runs at the board's
// power limit, and at 64 registers the converter body loses instruction-level parallelism)
int tp_teams() {
    static int t = -1;
    if (t < 0) { const char* e = getenv("DAD_TC_TP_TEAMS"); t = e ? atoi(e) : 2; if (t != 2 && t != 3) t = 2; }
    return t;
}

template <int NF, int TEAMS>
int launch_fwd_tp(const CUtensorMap (&tw)[4], const CUtensorMap& txh, const CUtensorMap& txl, const FwdParams& p, cudaStream_t st) {
    using Cfg = TpCfg<NF>;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_enc_fwd_tp<NF, TEAMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        if (e != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "aug_encoder_fwd(tp): cannot opt in to %u B shared memory: %s", Cfg::SMEM, cudaGetErrorString(e));
        attr = true;
    }
    const int units = (int)ceil_div64(p.N, Cfg::TF), max_pairs = sm_count() / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute at[1];
    cfg.gridDim = dim3(2 * (units < max_pairs ? units : max_pairs));
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.blockDim = dim3(128 + TEAMS * 256 + (TEAMS == 3 ? 128 : 256));
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_enc_fwd_tp<NF, TEAMS>, tw[0], tw[1], tw[2], tw[3], txh, txl, p);
    if (e != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "k_enc_fwd_tp<%d> launch: %s", NF, cudaGetErrorString(e));
    return DAD_OK;
}

}  // namespace

void enc_tc_set_cta_pair(int mode) { g_cta2 = (mode < 0 || mode > 2) ? 2 : mode; }
int enc_tc_get_cta_pair() { return cta2_mode(); }
// products of the dW GEMM (2: x_hi only, 3: fp32-grade): process-wide, DAD_TC_BWD_PRODUCTS=2|3 or dad_tc_set_bwd_products()
static int g_bwd_products = -1;
void enc_tc_set_bwd_products(int n) { g_bwd_products = n == 2 ? 2 : 3; }
int enc_tc_get_bwd_products() {
    if (g_bwd_products < 0) {
        const char* e = getenv("DAD_TC_BWD_PRODUCTS");
        g_bwd_products = (e && atoi(e) == 2) ? 2 : 3;
    }
    return g_bwd_products;
}

// pooled partials of ONE view: [tiles * 4 groups][2 slots][256] (tiles rounded up to a whole pair)
static size_t fwd_part_bytes(int64_t B, int64_t T) {
    const int64_t tiles = (ceil_div64(B * T, BM) + 1) / 2 * 2;
    return (size_t)(tiles * 4 * 2 * DAD_H * sizeof(float));
}
size_t enc_fwd_tc_ws_bytes(int64_t B, int64_t T) { return 2 * fwd_part_bytes(B, T); }

// One launch for one view (W1 == null) or for the weak -> teacher (W0, ap0) and strong -> student (W1, ap1) views of the same x.
// xs_hi / xs_lo: operand cache of the LAST view (hi only: xs_lo == null).
int launch_enc_fwd_tc(const float* x, const uint8_t* pad, const void* W0_hi, const void* W0_lo, const float* bias0, const AugParams& ap0, float* emb0,
                      uint32_t* bits0, const void* W1_hi, const void* W1_lo, const float* bias1, const AugParams* ap1, float* emb1, uint32_t* bits1,
                      void* xs_hi, void* xs_lo, float* lens_out, void* ws, size_t ws_bytes, int64_t B, int64_t T, cudaStream_t st) {
    const int nv = W1_hi ? 2 : 1;
    if (T < 32) return dad_set_error(DAD_ERR_UNSUPPORTED, "encoder_fwd: the tcgen05 path needs T >= 32 (got %lld); use DAD_GEMM_FP32", (long long)T);
    if (B * T >= (1ll << 31)) return dad_set_error(DAD_ERR_UNSUPPORTED, "encoder_fwd: the tcgen05 path indexes frames with 32 bits (B*T = %lld)", (long long)(B * T));
    uintptr_t al = reinterpret_cast<uintptr_t>(W0_hi) | reinterpret_cast<uintptr_t>(W0_lo) | reinterpret_cast<uintptr_t>(W1_hi) | reinterpret_cast<uintptr_t>(W1_lo) |
                   reinterpret_cast<uintptr_t>(xs_hi) | reinterpret_cast<uintptr_t>(xs_lo);
    if (al & 15u) return dad_set_error(DAD_ERR_ARG, "encoder_fwd: W_hi/W_lo/xs_hi/xs_lo must be 16-byte aligned (TMA)");
    if (ws_bytes < enc_fwd_tc_ws_bytes(B, T)) return dad_set_error(DAD_ERR_WS, "encoder_fwd(tc): workspace too small");
    const bool shared = nv == 2 && ap1 && ap0.mode == 0 && ap1->mode == 0;  // one clean view, two weight sets (dad_predict)
    if (nv == 2 && !shared) {
        if (!ap1 || ap0.mode != 1 || ap1->mode != 2) return dad_set_error(DAD_ERR_ARG, "aug_encoder_fwd: view 0 must be the weak and view 1 the strong augmentation");
        if (ap0.noise_src != ap1->noise_src) return dad_set_error(DAD_ERR_ARG, "aug_encoder_fwd: both views must use the same noise source");
        if (ap0.noise_src == 1 && (ap0.seed != ap1->seed || ap0.step_ptr != ap1->step_ptr || ap0.frame_offset != ap1->frame_offset ||
                                   ap0.stream + 1 != ap1->stream || (ap0.stream & 1ull)))
            return dad_set_error(DAD_ERR_ARG, "aug_encoder_fwd: Philox views must share seed / step / frame offset and use streams (2k, 2k+1)");
    }
    const bool noisy = ap0.mode != 0;
    const bool tp = nv == 2 && !shared && cta2_mode() == 2;  // training's fused noisy forward: the transposed CTA-pair kernel
    const int tp_nf = tp ? tp_frames_per_cta(B * T) : 0;
    const bool cta2 = tp || cta2_mode() == 1 || (cta2_mode() == 2 && !noisy);
    const uint32_t wrows = cta2 ? 128 : 256;
    CUtensorMap tw[4], txh, txl;
    memset(tw, 0, sizeof(tw));
    memset(&txh, 0, sizeof(txh));
    memset(&txl, 0, sizeof(txl));
    int rc = make_tmap_f16_sw64(&tw[0], W0_hi, DAD_H, wrows);
    if (!rc) rc = make_tmap_f16_sw64(&tw[1], W0_lo, DAD_H, wrows);
    if (!rc && nv == 2) rc = make_tmap_f16_sw64(&tw[2], W1_hi, DAD_H, wrows);
    if (!rc && nv == 2) rc = make_tmap_f16_sw64(&tw[3], W1_lo, DAD_H, wrows);
    if (rc) return rc;
    FwdParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.pad = pad;
    const size_t pb = fwd_part_bytes(B, T);
    p.v[0].bias = bias0; p.v[0].ap = ap0; p.v[0].part = reinterpret_cast<float*>(ws); p.v[0].bits = bits0;
    if (nv == 2) {
        p.v[1].bias = bias1; p.v[1].ap = *ap1; p.v[1].part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + pb); p.v[1].bits = bits1;
    }
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("DAD_TC_DBG"); dbg = e ? atoi(e) : 0; }
        p.dbg = dbg;
    }
    p.shared_a = shared ? 1 : 0;
    p.cache = xs_hi ? (xs_lo ? 2 : 1) : 0;
    if (p.cache) {
        rc = make_tmap_f16_sw64(&txh, xs_hi, (uint64_t)(B * T), tp ? (uint32_t)tp_nf : BM);
        if (!rc && p.cache > 1) rc = make_tmap_f16_sw64(&txl, xs_lo, (uint64_t)(B * T), tp ? (uint32_t)tp_nf : BM);
        if (rc) return rc;
    }
    p.N = B * T; p.T = (int)T; p.num_tiles = (int)ceil_div64(B * T, BM);
    if (tp && tp_nf == 128) rc = launch_fwd_tp<128, 2>(tw, txh, txl, p, st);
    else if (tp) rc = tp_teams() == 3 ? launch_fwd_tp<64, 3>(tw, txh, txl, p, st) : launch_fwd_tp<64, 2>(tw, txh, txl, p, st);
    else if (nv == 2) rc = cta2 ? launch_fwd_variant<2, true>(tw, txh, txl, p, st) : launch_fwd_variant<2, false>(tw, txh, txl, p, st);
    else rc = cta2 ? launch_fwd_variant<1, true>(tw, txh, txl, p, st) : launch_fwd_variant<1, false>(tw, txh, txl, p, st);
    if (rc) return rc;
    PoolArgs pa;
    pa.part[0] = p.v[0].part; pa.part[1] = p.v[1].part;
    pa.emb[0] = emb0; pa.emb[1] = emb1;
    pa.lens_out = lens_out;
    k_pool_finalize<<<dim3((unsigned)B, nv), DAD_H, 0, st>>>(pa, pad, (int)T);
    DAD_CHECK_LAUNCH("k_pool_finalize");
    return DAD_OK;
}

// =====================================================================================================================
// Backward:  dW[256,768] = sum_n dpre[n,:] (x) aug(x)[n,:],   dpre[n,h] = relu_bit(n,h) * valid(n) * g[b(n),h] / len_b
//   = a [256 x N] . [N x 768] GEMM whose reduction axis is the frame axis: BOTH operands are MN-major in shared memory
//   (frame rows, 128-byte lines of 64 contiguous h / d elements, SWIZZLE_128B), so nothing is transposed.
//   grid = (3 d-slices of 256, S frame-range splits); per CTA two 128 x 256 fp32 accumulators (all 512 TMEM columns).
//   The A operand is never read from memory as numbers: it is SELECTED from a per-utterance table of pre-split
//   scaled g/len values (G_hi/G_lo, fp16) by the 1-bit ReLU mask of the forward.  Split partials are reduced in fixed order
//   by dad_grad_reduce (deterministic, no fp32 atomics).
// =====================================================================================================================
namespace {

constexpr int BWD_BK = 32, BWD_THREADS = 512;
constexpr uint32_t BOP_BYTES = DAD_H * BWD_BK * 2;                  // 16 KB: one operand (hi or lo), 256 x 32 fp16
constexpr uint32_t BOFF_AHI = 0, BOFF_ALO = BOP_BYTES, BOFF_BHI = 2 * BOP_BYTES, BOFF_BLO = 3 * BOP_BYTES;
// PRODS = 3 (default): dpre_hi*x_hi + dpre_hi*x_lo + dpre_lo*x_hi (fp32-grade, as the forward).  PRODS = 2 (opt-in,
// dad_tc_set_bwd_products): the B operand is x_hi only -- dpre_hi*x_hi + dpre_lo*x_hi: one third less tensor work, no B_lo tile in
// shared memory (a fourth pipeline stage instead) and an operand cache of half the size (xs_hi only).  Dropping x_lo perturbs every
// term of the frame sum by an independent relative 2^-12 (the fp16 rounding of x); on gradient elements whose terms cancel that is
// a relative error of ~1e-4 where fp32 has 1e-7, and the teacher-forced weight tolerance of the parity tests (1e-5 after Adam on
// elements with |g| >= 1e-6) does NOT hold (measured 1e-5 .. 8e-4) -- hence not the default.
template <int PRODS>
struct BwdCfg {
    static constexpr uint32_t STAGE_BYTES = (PRODS == 3 ? 4 : 3) * BOP_BYTES;  // A_hi, A_lo, B_hi (, B_lo) = 48 / 64 KB
    static constexpr int STAGES = PRODS == 3 ? 3 : 4;
    static constexpr uint32_t OFF_LUT = STAGES * STAGE_BYTES;   // 256 x uint4: ReLU byte -> 8 fp16 lane masks
    static constexpr uint32_t OFF_BARS = OFF_LUT + 256 * 16;
    static constexpr uint32_t SMEM = OFF_BARS + 128 + 1024;
};
constexpr uint32_t MN_LBO = (BWD_BK / 8) * 1024;                    // pitch between 64-element M/N chunks (4 k-groups of 1 KB)
constexpr uint32_t MN_SBO = 1024;                                   // pitch between 8-frame groups

struct BwdParams {
    const float* x;
    const uint8_t* pad;
    AugParams ap;
    const uint32_t* bits;            // [N][8]
    const uint16_t* g_hi;            // [B][256]  fp16(s * g/len)
    const uint16_t* g_lo;            // [B][256]  fp16(s * g/len - hi)
    const float* unscale;            // [1] 1 / (s * TC_XS): applied to the accumulators in the epilogue (exact power of two)
    float* dWp;                      // [S][256][768]
    int64_t N;
    int T;
    int B;
    int utt_per_split;               // utterances per frame-range split (grid.y); k-blocks are aligned to utterance starts
    int use_xs;                      // B operand comes from the forward's operand cache by TMA (tm_xhi / tm_xlo); x / ap are not touched
};

constexpr int ABSMAX_GRID = 256;  // == DAD_H: k_bwd_prep folds the partials with one thread each

// per-CTA partial max of |g/len| (no atomics: the table scale must be bit-reproducible)
__global__ void __launch_bounds__(DAD_H) k_bwd_absmax(const float* __restrict__ g_emb, const float* __restrict__ lens, float* __restrict__ part, int B) {
    __shared__ float s_w[DAD_H / 32];
    float m = 0.f;
    for (int b = blockIdx.x; b < B; b += gridDim.x) m = fmaxf(m, fabsf(g_emb[(int64_t)b * DAD_H + threadIdx.x] / lens[b]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < DAD_H / 32; ++w) m = fmaxf(m, s_w[w]);
        part[blockIdx.x] = m;
    }
}

// G_hi/G_lo = fp16 split of s * g_emb / len (division as in autograd of sum/len), one thread per (b,h).
// s = 2^(13-e) with 2^e <= absmax < 2^(e+1): the largest entry lands in [2^13, 2^14), 18 octaves of entries keep all 22 bits.
__global__ void __launch_bounds__(DAD_H) k_bwd_prep(const float* __restrict__ g_emb, const float* __restrict__ lens, const float* __restrict__ part,
                                                    uint16_t* __restrict__ g_hi, uint16_t* __restrict__ g_lo, float* __restrict__ unscale) {
    __shared__ float s_scale;
    __shared__ float s_m[DAD_H / 32];
    {
        const float m = warp_max(part[threadIdx.x]);  // max is order-independent: the scale is bit-reproducible
        if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        float m = threadIdx.x < DAD_H / 32 ? s_m[threadIdx.x] : 0.f;
        m = warp_max(m);
        if (threadIdx.x == 0) {
            float sc = 1.0f;
            if (m > 0.f && m < 3.0e38f) {
                int e = (int)((__float_as_uint(m) >> 23) & 0xFFu) - 127;  // floor(log2 m) for normal m
                if (e < -100) e = -100;                                     // keep s and 1/s finite fp32 normals
                sc = __uint_as_float((uint32_t)(127 + 13 - e) << 23);
            }
            s_scale = sc;
            if (blockIdx.x == 0) unscale[0] = 1.0f / (sc * TC_XS);
        }
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * DAD_H + threadIdx.x;
    const float v = (g_emb[i] / lens[blockIdx.x]) * s_scale;
    unsigned short h, l;
    tc::split1_f16(v, h, l);
    g_hi[i] = h;
    g_lo[i] = l;
}

// db[h] = sum over valid frames of relu_bit(n,h) * g[b,h]/len_b = sum_b (g[b,h]/len_b) * count_b[h].
// `wpu` (1, 2, 4 or 8) WARPS per utterance, each counting the ReLU bits of a contiguous 1/wpu of the frames.  Lane (w = lane & 7,
// q = lane >> 3) walks frames q, q+4, ... of the warp's range for its 32-bit mask word w and counts all 32 hidden units at once
// in a VERTICAL counter: bit plane i holds bit i of the 32 counts.  Eight frames enter through a carry-save adder tree (7 CSAs =
// 14 LOP3 update the ones / twos / fours planes and emit one carry word that ripples through the upper planes): ~3-5 integer ops
// per frame, against 48 for sixteen registers of paired 16-bit counters ((word >> k) & 0x00010001 ...) and 256 for warp ballots.
// The four q-lanes of a word add their vertical counters with a shuffle butterfly (bit-serial full adder per plane), each then
// expands a quarter of the counts to integers; the warps of an utterance are summed through shared memory (integer: order-free).
// contrib[b][h] = count * g/len, and k_bwd_db_fold sums utterance groups in fixed order into the split partials consumed by
// dad_grad_reduce.  NPL planes per lane: frames per lane < 2^NPL.
// At small batch the kernel is a chain of dependent L2 round trips per warp (25 us at 512 utterances per data-parallel rank with
// one warp per utterance, 10 us with eight): the launcher picks wpu so that the grid has enough warps to cover the latency.
constexpr int DB_WARPS = 8;
static_assert(DB_WARPS * 32 == DAD_H, "k_bwd_db: one thread per hidden unit in the epilogue");
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
    const uint32_t s = a ^ b ^ c, cy = (a & b) | (c & (a | b));
    sum = s;
    carry = cy;
}
template <int NPL>
__global__ void __launch_bounds__(DB_WARPS * 32) k_bwd_db(const uint32_t* __restrict__ bits, const uint8_t* __restrict__ pad, const float* __restrict__ g_emb,
                                                         const float* __restrict__ lens, float* __restrict__ contrib, int T, int B, int wpu) {
    __shared__ int s_cnt[DB_WARPS][DAD_H];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int upc = DB_WARPS / wpu;                   // utterances per CTA
    const int b = blockIdx.x * upc + warp / wpu;      // this warp's utterance (may be >= B in the last CTA)
    const int sub = warp % wpu;
    const int w = lane & 7, q = lane >> 3;
    const int per = ((T + wpu - 1) / wpu + 3) & ~3;   // frames per warp, a multiple of the 4 frames a warp reads at once
    const int tb = sub * per;
    const int te = b < B ? (tb + per < T ? tb + per : T) : 0;
    uint32_t P[NPL + 2];                              // + 2 planes for the sum over the four q-lanes
#pragma unroll
    for (int i = 0; i < NPL + 2; ++i) P[i] = 0u;
    const int64_t n0 = (int64_t)b * T;
    for (int t0 = tb + q; t0 < te; t0 += 32) {  // 8 independent loads in flight per lane
        uint32_t x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int t = t0 + 4 * u;
            const int64_t n = n0 + t;
            x[u] = t < te ? __ldg(bits + n * DAD_HW + w) : 0u;  // a warp reads 4 frames x 32 B = 128 contiguous bytes
            if (pad && t < te && pad[n]) x[u] = 0u;
        }
        uint32_t t0_, t1_, t2_, t3_, f0, f1, e0;
        csa(P[0], x[0], x[1], P[0], t0_);
        csa(P[0], x[2], x[3], P[0], t1_);
        csa(P[1], t0_, t1_, P[1], f0);
        csa(P[0], x[4], x[5], P[0], t2_);
        csa(P[0], x[6], x[7], P[0], t3_);
        csa(P[1], t2_, t3_, P[1], f1);
        csa(P[2], f0, f1, P[2], e0);
#pragma unroll
        for (int i = 3; i < NPL; ++i) {  // ripple the eights carry through the upper planes
            const uint32_t cy = P[i] & e0;
            P[i] ^= e0;
            e0 = cy;
        }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {  // add the vertical counters of the four q-lanes (all four end up with the total)
        uint32_t cy = 0u;
#pragma unroll
        for (int i = 0; i < NPL + 2; ++i) {
            const uint32_t y = __shfl_xor_sync(0xffffffffu, P[i], o);
            csa(P[i], y, cy, P[i], cy);
        }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // this lane expands bits k and k + 16 of its word for k = 4q .. 4q+3
        const int k = 4 * q + kk;
        uint32_t acc = 0u;
#pragma unroll
        for (int i = 0; i < NPL + 2; ++i) acc += ((P[i] >> k) & 0x00010001u) << i;  // counts < 2^16: the two halves never meet
        s_cnt[warp][w * 32 + k] = (int)(acc & 0xFFFFu);
        s_cnt[warp][w * 32 + 16 + k] = (int)(acc >> 16);
    }
    __syncthreads();
    const int h = threadIdx.x;
    for (int j = 0; j < upc; ++j) {
        const int bj = blockIdx.x * upc + j;
        if (bj >= B) break;
        int cnt = 0;
        for (int ww = 0; ww < wpu; ++ww) cnt += s_cnt[j * wpu + ww][h];
        contrib[(int64_t)bj * DAD_H + h] = (float)cnt * (g_emb[(int64_t)bj * DAD_H + h] / lens[bj]);
    }
}
__global__ void __launch_bounds__(DAD_H) k_bwd_db_fold(const float* __restrict__ contrib, float* __restrict__ dbp, int B, int per) {
    const int b0 = blockIdx.x * per;
    const int b1 = b0 + per < B ? b0 + per : B;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int b = b0;
    for (; b + 3 < b1; b += 4) {
        s0 += contrib[(int64_t)b * DAD_H + threadIdx.x];
        s1 += contrib[(int64_t)(b + 1) * DAD_H + threadIdx.x];
        s2 += contrib[(int64_t)(b + 2) * DAD_H + threadIdx.x];
        s3 += contrib[(int64_t)(b + 3) * DAD_H + threadIdx.x];
    }
    for (; b < b1; ++b) s0 += contrib[(int64_t)b * DAD_H + threadIdx.x];
    dbp[(int64_t)blockIdx.x * DAD_H + threadIdx.x] = (s0 + s1) + (s2 + s3);
}

// ---------------------------------------------------------------------------------------------------------------------
// Converter warp body of the backward kernel.  The k-blocks (32 frames) are aligned to UTTERANCE starts: utterance b is covered by
// KPU = ceil(T / 32) k-blocks, the last one padded with zero rows of the A operand (its B rows are whatever follows in memory:
// they meet zeros).  Inside a k-block every frame therefore belongs to one utterance, so the pre-split fp16 g/len chunk of that
// utterance (this lane's 8 hidden units hq = lane: two 16-byte registers sets, fetched once per utterance) serves all of it.
// Warp cw owns the 4 consecutive frames cw*4 .. cw*4+3 of every k-block, for BOTH operands:
//   A operand (dpre^T, 256 hidden units): the frame's ReLU byte for this lane's 8 units -> a shared-memory LUT of 256 uint4
//     (8 fp16 lane masks) -> AND with the g/len chunk -> two 16-byte stores.  Padded frames carry zero ReLU bits (the forward
//     kernels write them so) and frames beyond the utterance are zeroed here: no padding-mask reads, no per-frame utterance lookup.
//     The first version spent ~490 instructions per k-block and thread on this (ncu: cached backward 0.75 ms, tensor pipe 45 %
//     active); this one ~80.
//   B operand (aug(x), 256 d columns of this CTA's slice), unless the operand cache feeds it by TMA (XS): lane -> float4 columns
//     lane and lane+32 (two coalesced 512-byte loads per row), augmentation regenerated exactly as in the forward, fp16 hi (/ lo).
// The ReLU words (and the x rows) of the next k-blocks are in flight while one is converted (NSET register sets).
// ---------------------------------------------------------------------------------------------------------------------
struct BwdRegs {
    float4 x[8];     // [r4*2 + half]  (unused with the operand cache)
    uint32_t w[4];   // ReLU mask word holding this lane's byte, per frame (0 beyond the utterance)
};

template <int PRODS, int MODE, bool PHILOX, bool XS>
__device__ __forceinline__ void bwd_convert(const BwdParams& p, uint8_t* smem, uint64_t* full, uint64_t* empty, const int cw, const int lane,
                                            const int dslice, const int b_begin, const int b_end) {
    using Cfg = BwdCfg<PRODS>;
    constexpr int BWD_STAGES = Cfg::STAGES;
    constexpr uint32_t BSTAGE_BYTES = Cfg::STAGE_BYTES;
    const AugParams& ap = p.ap;
    const uint32_t smem_a = tc::smem_u32(smem);
    const uint32_t lut = smem_a + Cfg::OFF_LUT;
    const unsigned long long seed = (MODE != 0 && PHILOX) ? aug_seed(ap) : 0ull;
    const float radius_k = NOISE_UNIT_K * ap.sigma * ap.sigma;
    const int T = p.T;
    const int KPU = (T + BWD_BK - 1) / BWD_BK;
    // shared-memory offsets (MN-major, SWIZZLE_128B): frame r = cw*4 + j -> 8-frame group cw>>1, row-in-atom r7 = (cw&1)*4 + j
    const uint32_t r7b = (uint32_t)(cw & 1) * 4u;
    const uint32_t xoff0 = (uint32_t)(lane >> 4) * MN_LBO + (uint32_t)(cw >> 1) * MN_SBO + (uint32_t)(lane & 1) * 8u;  // + 2*MN_LBO for the second half
    const uint32_t xc = (uint32_t)(lane & 15) >> 1;
    const uint32_t aoff0 = (uint32_t)(lane >> 3) * MN_LBO + (uint32_t)(cw >> 1) * MN_SBO;
    const uint32_t ac = (uint32_t)lane & 7u;
    const uint32_t bshift = (uint32_t)(lane & 3) * 8u;
    float4 keep0 = make_float4(1.f, 1.f, 1.f, 1.f), keep1 = keep0;
    const bool has_keep = !XS && MODE == 2 && ap.feat_keep != nullptr;
    if (has_keep) {
        keep0 = __ldg(reinterpret_cast<const float4*>(ap.feat_keep + dslice * 256 + lane * 4));
        keep1 = __ldg(reinterpret_cast<const float4*>(ap.feat_keep + dslice * 256 + (lane + 32) * 4));
    }
    const int nkb = (b_end - b_begin) * KPU;
    if (nkb <= 0) return;

    // flat k-block index it -> (utterance, first frame-in-utterance of this warp): advanced incrementally by the LOAD stream
    int lb = b_begin, lk = 0;  // utterance / k-block-in-utterance of the next load
    auto load = [&](BwdRegs& R) {
        const int t0 = lk * BWD_BK + cw * 4;
        const int64_t n0 = (int64_t)lb * T + t0;
        const uint32_t* bw = p.bits + n0 * DAD_HW + (lane >> 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) R.w[j] = (t0 + j < T) ? __ldg(bw + j * DAD_HW) : 0u;
        if (!XS) {
            const float* src = p.x + n0 * DAD_D + dslice * 256 + lane * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool in = n0 + j < p.N;  // rows of the NEXT utterance are read (they meet zero A rows); rows beyond N are not
                R.x[2 * j] = in ? __ldg(reinterpret_cast<const float4*>(src + j * DAD_D)) : make_float4(0.f, 0.f, 0.f, 0.f);
                R.x[2 * j + 1] = in ? __ldg(reinterpret_cast<const float4*>(src + j * DAD_D + 128)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (++lk == KPU) { lk = 0; ++lb; }
    };
    auto tab_ldg = [&](const uint16_t* base, int b) { return __ldg(reinterpret_cast<const uint4*>(base + (int64_t)b * DAD_H + lane * 8)); };

    uint32_t s = 0, ph = 0;
    int cb = b_begin, ck = 0;  // utterance / k-block-in-utterance being converted
    uint4 gh = tab_ldg(p.g_hi, cb), gl = tab_ldg(p.g_lo, cb);
    uint4 gh_n = gh, gl_n = gl;
    auto convert = [&](BwdRegs& R) {
        if (ck == 0 && cb + 1 < b_end) { gh_n = tab_ldg(p.g_hi, cb + 1); gl_n = tab_ldg(p.g_lo, cb + 1); }  // next utterance's chunk: a whole utterance ahead
        tc::mbar_wait(&empty[s], ph ^ 1u);
        const uint32_t st = smem_a + s * BSTAGE_BYTES;
        // ---- A operand: dpre^T chunks = g/len chunk of the utterance, masked by the frame's ReLU byte (LUT: byte -> 8 fp16 masks)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t byte = (R.w[j] >> bshift) & 0xFFu;
            const uint4 m = tc::lds128(lut + byte * 16u);
            const uint32_t r7 = r7b + (uint32_t)j;
            const uint32_t off = aoff0 + r7 * 128u + ((ac ^ r7) << 4);
            tc::sts128(st + BOFF_AHI + off, make_uint4(gh.x & m.x, gh.y & m.y, gh.z & m.z, gh.w & m.w));
            tc::sts128(st + BOFF_ALO + off, make_uint4(gl.x & m.x, gl.y & m.y, gl.z & m.z, gl.w & m.w));
        }
        // ---- B operand: aug(x) rows -> fp16 hi (/ lo) (scaled by TC_XS); with the operand cache (XS) the TMA warp loads it instead
        if (!XS) {
            const int t0 = ck * BWD_BK + cw * 4;
            const int64_t n0 = (int64_t)cb * T + t0;
            uint32_t zx = 0;  // rows that contribute zeros: inside the temporal mask span of the strong view
            if (MODE == 2 && ap.t_start && ap.mask_len > 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int t = t0 + j, bb = cb;
                    if (t >= T) { t -= T; ++bb; }
                    if (bb < p.B) {
                        const int s0 = __ldg(ap.t_start + bb);
                        if (t >= s0 && t < s0 + ap.mask_len) zx |= 1u << j;
                    }
                }
            }
            if (MODE != 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float4& va = R.x[2 * j + half];
                        const int f = lane + 32 * half;
                        if (PHILOX) {
                            NoiseBM4 z;
                            philox_bm4(seed, (uint32_t)ap.stream, (unsigned long long)(n0 + j) + (unsigned long long)ap.frame_offset, (uint32_t)(dslice * 64 + f), radius_k, z);
                            va = bm4_add(z, va);
                        } else {
                            const float* nz = ap.noise + (n0 + j) * DAD_D + dslice * 256 + f * 4;
                            const float4 na = (n0 + j < p.N) ? __ldg(reinterpret_cast<const float4*>(nz)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            va.x = aug_add_noise(va.x, na.x, ap.sigma); va.y = aug_add_noise(va.y, na.y, ap.sigma);
                            va.z = aug_add_noise(va.z, na.z, ap.sigma); va.w = aug_add_noise(va.w, na.w, ap.sigma);
                        }
                        if (has_keep) {
                            const float4 kp = half ? keep1 : keep0;
                            va.x = __fmul_rn(va.x, kp.x); va.y = __fmul_rn(va.y, kp.y); va.z = __fmul_rn(va.z, kp.z); va.w = __fmul_rn(va.w, kp.w);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t r7 = r7b + (uint32_t)j;
                const uint32_t roff = xoff0 + r7 * 128u + ((xc ^ r7) << 4);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const float4 v = R.x[2 * j + half];
                    const float sc = ((zx >> j) & 1u) ? 0.f : TC_XS;
                    uint32_t h01, l01, h23, l23;
                    tc::split2_f16(v.x * sc, v.y * sc, h01, l01);
                    tc::split2_f16(v.z * sc, v.w * sc, h23, l23);
                    const uint32_t off = roff + (uint32_t)half * 2u * MN_LBO;
                    tc::sts64(st + BOFF_BHI + off, h01, h23);
                    if (PRODS == 3) tc::sts64(st + BOFF_BLO + off, l01, l23);
                }
            }
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&full[s]);
        if (++s == (uint32_t)BWD_STAGES) { s = 0; ph ^= 1u; }
        if (++ck == KPU) { ck = 0; ++cb; gh = gh_n; gl = gl_n; }
    };

    BwdRegs RA, RB, RC;
    load(RA);
    if (nkb > 1) load(RB);
    for (int kb = 0; kb < nkb; kb += 3) {
        if (kb + 2 < nkb) load(RC);
        convert(RA);
        if (kb + 1 < nkb) {
            if (kb + 3 < nkb) load(RA);
            convert(RB);
        }
        if (kb + 2 < nkb) {
            if (kb + 4 < nkb) load(RB);
            convert(RC);
        }
    }
}

template <int PRODS>
__global__ void __launch_bounds__(BWD_THREADS, 1)
k_enc_bwd_tc(const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo, BwdParams p) {
    using Cfg = BwdCfg<PRODS>;
    constexpr int BWD_STAGES = Cfg::STAGES;
    constexpr uint32_t BSTAGE_BYTES = Cfg::STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
    uint64_t* full = bars;                  // [STAGES] 8 converter warps (+ 1 TMA expect_tx with the operand cache)
    uint64_t* empty = bars + BWD_STAGES;    // [STAGES] 1 (tcgen05.commit)
    uint64_t* tfull = bars + 2 * BWD_STAGES;  // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * BWD_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dslice = blockIdx.x, split = blockIdx.y;
    const int b_begin = split * p.utt_per_split;
    const int b_end = b_begin + p.utt_per_split < p.B ? b_begin + p.utt_per_split : p.B;
    const int KPU = (p.T + BWD_BK - 1) / BWD_BK;
    const int nkb = b_end > b_begin ? (b_end - b_begin) * KPU : 0;

    if (threadIdx.x < 256) {  // ReLU byte -> eight 16-bit lane masks (bit k set -> 0xFFFF in half-word k)
        const uint32_t v = threadIdx.x;
        uint32_t m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = ((v >> (2 * k)) & 1u ? 0x0000FFFFu : 0u) | ((v >> (2 * k + 1)) & 1u ? 0xFFFF0000u : 0u);
        *reinterpret_cast<uint4*>(smem + Cfg::OFF_LUT + v * 16) = make_uint4(m[0], m[1], m[2], m[3]);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < BWD_STAGES; ++s) { tc::mbar_init(&full[s], N_CONV_WARPS + (p.use_xs ? 1 : 0)); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(tfull, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_ptr, TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer (operand cache -> B operand) =====================
        // one box = 32 frames x 64 d-columns (128-byte rows, SWIZZLE_128B) = exactly one MN-major 64-column chunk of the stage;
        // frames beyond N are zero-filled by the tensor map
        tc::reg_dec<40>();
        if (p.use_xs && lane == 0) {
            tc::prefetch_tmap(&tm_xhi);
            if (PRODS == 3) tc::prefetch_tmap(&tm_xlo);
            int kb = 0;
            for (int b = b_begin; b < b_end; ++b)
                for (int kl = 0; kl < KPU; ++kl, ++kb) {
                    const uint32_t s = kb % BWD_STAGES, ph = (kb / BWD_STAGES) & 1;
                    tc::mbar_wait(&empty[s], ph ^ 1);
                    tc::mbar_arrive_expect_tx(&full[s], (PRODS == 3 ? 2 : 1) * BOP_BYTES);
                    const int row = b * p.T + kl * BWD_BK;
                    uint8_t* st = smem + s * BSTAGE_BYTES;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        tc::tma_load_2d(st + BOFF_BHI + c * MN_LBO, &tm_xhi, &full[s], dslice * 256 + c * 64, row);
                        if (PRODS == 3) tc::tma_load_2d(st + BOFF_BLO + c * MN_LBO, &tm_xlo, &full[s], dslice * 256 + c * 64, row);
                    }
                }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        tc::reg_dec<40>();
        constexpr uint32_t idesc = tc::make_idesc_f16(128, DAD_H, 1, 1);  // A and B both MN-major
        for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t s = kb % BWD_STAGES, ph = (kb / BWD_STAGES) & 1;
            tc::mbar_wait(&full[s], ph);
            tc::tcgen05_fence_after();
            if (lane == 0) {
                const uint32_t base = tc::smem_u32(smem + s * BSTAGE_BYTES);
#pragma unroll
                for (int mh = 0; mh < 2; ++mh)
#pragma unroll
                    for (int prod = 0; prod < PRODS; ++prod) {
                        // 3 products: hi*hi, hi*lo, lo*hi;  2 products: hi*hi, lo*hi (B = x_hi only)
                        const uint32_t a_off = ((prod == PRODS - 1) ? BOFF_ALO : BOFF_AHI) + mh * 2 * MN_LBO;
                        const uint32_t b_off = (PRODS == 3 && prod == 1) ? BOFF_BLO : BOFF_BHI;
#pragma unroll
                        for (int ks = 0; ks < BWD_BK / 16; ++ks) {
                            const uint64_t ad = tc::make_smem_desc(base + a_off + ks * 2 * MN_SBO, MN_LBO, MN_SBO);
                            const uint64_t bd = tc::make_smem_desc(base + b_off + ks * 2 * MN_SBO, MN_LBO, MN_SBO);
                            tc::umma_f16(tmem_base + mh * DAD_H, ad, bd, idesc, (kb | prod | ks) != 0);
                        }
                    }
                tc::umma_commit(&empty[s]);
                if (kb == nkb - 1) tc::umma_commit(tfull);
            }
            __syncwarp();
        }
    } else if (warp >= CONV_WARP0 && warp < CONV_WARP0 + N_CONV_WARPS) {
        // ===================== converters =====================
        tc::reg_inc<200>();  // 128 x 40 (control) + 256 x 200 (converters: three landing sets) + 128 x 64 (epilogue) <= 64 K
        const int cw = warp - CONV_WARP0;
        const int variant = p.use_xs ? 5 : (p.ap.mode == 0 ? 0 : (p.ap.mode * 2 - 1 + (p.ap.noise_src == 1 ? 1 : 0)));
        switch (variant) {
            case 0: bwd_convert<PRODS, 0, false, false>(p, smem, full, empty, cw, lane, dslice, b_begin, b_end); break;
            case 1: bwd_convert<PRODS, 1, false, false>(p, smem, full, empty, cw, lane, dslice, b_begin, b_end); break;
            case 2: bwd_convert<PRODS, 1, true, false>(p, smem, full, empty, cw, lane, dslice, b_begin, b_end); break;
            case 3: bwd_convert<PRODS, 2, false, false>(p, smem, full, empty, cw, lane, dslice, b_begin, b_end); break;
            case 4: bwd_convert<PRODS, 2, true, false>(p, smem, full, empty, cw, lane, dslice, b_begin, b_end); break;
            default: bwd_convert<PRODS, 0, false, true>(p, smem, full, empty, cw, lane, dslice, b_begin, b_end); break;  // A operand only
        }
    } else if (warp >= EPI_WARP0) {
        // ===================== epilogue: TMEM -> split-K partial dW =====================
        tc::reg_dec<64>();
        const int q = warp - EPI_WARP0;
        float* out = p.dWp + (int64_t)split * DAD_H * DAD_D;
        const float unscale = __ldg(p.unscale);
        if (nkb > 0) {
            tc::mbar_wait(tfull, 0);
            tc::tcgen05_fence_after();
        }
#pragma unroll 1
        for (int mh = 0; mh < 2; ++mh) {
            const int h = mh * 128 + q * 32 + lane;
            float* row = out + (int64_t)h * DAD_D + dslice * 256;
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
                uint32_t r[32];
                if (nkb > 0) {
                    tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + mh * DAD_H + c * 32, r);
                    tc::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(row + c * 32 + j) = make_float4(__uint_as_float(r[j]) * unscale, __uint_as_float(r[j + 1]) * unscale,
                                                                                __uint_as_float(r[j + 2]) * unscale, __uint_as_float(r[j + 3]) * unscale);
            }
        }
    } else {
        tc::reg_dec<40>();  // warps 2, 3: same warpgroup as the TMA / MMA warps
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

static int bwd_utt_per_split(int64_t B) { return (int)ceil_div64(B, 49); }
int enc_bwd_tc_splits(int64_t B) {
    // grid = 3 d-slices x S frame-range splits of WHOLE utterances: one wave of 147 CTAs at S = 49 (B >= 49)
    if (B < 1) return 1;
    return (int)ceil_div64(B, bwd_utt_per_split(B));
}
static inline size_t bwd_align(size_t v) { return (v + 255) & ~(size_t)255; }
struct BwdCarve { size_t g_hi, g_lo, scal, contrib, total; };
static BwdCarve bwd_carve(int64_t B) {
    BwdCarve c;
    size_t o = 0;
    c.g_hi = o; o = bwd_align(o + (size_t)B * DAD_H * sizeof(uint16_t));
    c.g_lo = o; o = bwd_align(o + (size_t)B * DAD_H * sizeof(uint16_t));
    c.scal = o; o = bwd_align(o + (ABSMAX_GRID + 64) * sizeof(float));   // [0,ABSMAX_GRID) absmax partials, [ABSMAX_GRID] unscale
    c.contrib = o; o = bwd_align(o + (size_t)B * DAD_H * sizeof(float));
    c.total = o;
    return c;
}
size_t enc_bwd_tc_ws_bytes(int64_t B, int64_t) { return bwd_carve(B).total; }

static int g_bwd_lane = 0;
int launch_enc_bwd_tc(const float* x, const uint8_t* pad, const AugParams& ap, const uint32_t* bits, const void* xs_hi, const void* xs_lo,
                      const float* g_emb, const float* lens, float* dWp, float* dbp, int splits, void* ws, size_t ws_bytes, int64_t B, int64_t T,
                      cudaStream_t st) {
    const BwdCarve cv = bwd_carve(B);
    if (ws_bytes < cv.total) return dad_set_error(DAD_ERR_WS, "encoder_bwd(tc): workspace too small");
    if (T < 32) return dad_set_error(DAD_ERR_UNSUPPORTED, "encoder_bwd: the tcgen05 path needs T >= 32 (got %lld); use DAD_GEMM_FP32", (long long)T);
    if (B * T >= (1ll << 31)) return dad_set_error(DAD_ERR_UNSUPPORTED, "encoder_bwd: the tcgen05 path indexes frames with 32 bits (B*T = %lld)", (long long)(B * T));
    const int prods = enc_tc_get_bwd_products();
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_enc_bwd_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdCfg<2>::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_enc_bwd_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdCfg<3>::SMEM);
        if (e != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "encoder_bwd(tc): cannot opt in to %u B shared memory: %s", BwdCfg<2>::SMEM, cudaGetErrorString(e));
        attr = true;
    }
    uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
    uint16_t* g_hi = reinterpret_cast<uint16_t*>(w8 + cv.g_hi);
    uint16_t* g_lo = reinterpret_cast<uint16_t*>(w8 + cv.g_lo);
    float* scal = reinterpret_cast<float*>(w8 + cv.scal);
    float* contrib = reinterpret_cast<float*>(w8 + cv.contrib);
    // the bias gradient (ReLU bits x g/len, no GEMM operand) runs on the library's helper stream beside absmax -> prep -> dW GEMM
    int wpu = 1;  // warps per utterance: enough warps in the grid to cover the load latency (~16 per SM)
    while (wpu < DB_WARPS && B * wpu < (int64_t)sm_count() * 16) wpu *= 2;
    const unsigned db_grid = (unsigned)ceil_div64(B * wpu, DB_WARPS);
    const int64_t frames_per_lane = ceil_div64(ceil_div64(T, wpu) + 3, 4) + 1;
    if (frames_per_lane >= 16384)
        return dad_set_error(DAD_ERR_UNSUPPORTED, "encoder_bwd: T = %lld frames is beyond the bias-gradient counters (T < 65536)", (long long)T);
    if (splits != enc_bwd_tc_splits(B)) return dad_set_error(DAD_ERR_ARG, "encoder_bwd(tc): splits must equal dad_encoder_bwd_splits()");
    SideLane* lane = side_lane(g_bwd_lane, st);
    const bool forked = side_fork(lane, st);
    cudaStream_t sdb = forked ? lane->side : st;
    if (frames_per_lane < 256)
        k_bwd_db<8><<<db_grid, DB_WARPS * 32, 0, sdb>>>(bits, pad, g_emb, lens, contrib, (int)T, (int)B, wpu);
    else
        k_bwd_db<14><<<db_grid, DB_WARPS * 32, 0, sdb>>>(bits, pad, g_emb, lens, contrib, (int)T, (int)B, wpu);
    DAD_CHECK_LAUNCH("k_bwd_db");
    k_bwd_db_fold<<<splits, DAD_H, 0, sdb>>>(contrib, dbp, (int)B, (int)ceil_div64(B, splits));
    DAD_CHECK_LAUNCH("k_bwd_db_fold");
    k_bwd_absmax<<<ABSMAX_GRID, DAD_H, 0, st>>>(g_emb, lens, scal, (int)B);
    DAD_CHECK_LAUNCH("k_bwd_absmax");
    k_bwd_prep<<<(unsigned)B, DAD_H, 0, st>>>(g_emb, lens, scal, g_hi, g_lo, scal + ABSMAX_GRID);
    DAD_CHECK_LAUNCH("k_bwd_prep");
    BwdParams p;
    p.x = x; p.pad = pad; p.ap = ap; p.bits = bits; p.g_hi = g_hi; p.g_lo = g_lo; p.unscale = scal + ABSMAX_GRID; p.dWp = dWp;
    p.N = B * T; p.T = (int)T; p.B = (int)B;
    p.utt_per_split = bwd_utt_per_split(B);
    CUtensorMap tm_hi, tm_lo;
    memset(&tm_hi, 0, sizeof(tm_hi));
    memset(&tm_lo, 0, sizeof(tm_lo));
    p.use_xs = xs_hi ? 1 : 0;
    if (p.use_xs) {
        if (prods == 3 && !xs_lo) return dad_set_error(DAD_ERR_ARG, "encoder_bwd: the 3-product backward needs xs_lo next to xs_hi");
        if ((reinterpret_cast<uintptr_t>(xs_hi) | reinterpret_cast<uintptr_t>(xs_lo)) & 15u)
            return dad_set_error(DAD_ERR_ARG, "encoder_bwd: xs_hi/xs_lo must be 16-byte aligned (TMA)");
        int rc = make_tmap_f16(&tm_hi, xs_hi, (uint64_t)p.N, DAD_D, BWD_BK);
        if (rc) return rc;
        if (prods == 3) rc = make_tmap_f16(&tm_lo, xs_lo, (uint64_t)p.N, DAD_D, BWD_BK);
        if (rc) return rc;
    }
    if (prods == 3) k_enc_bwd_tc<3><<<dim3(DAD_D / 256, splits), BWD_THREADS, BwdCfg<3>::SMEM, st>>>(tm_hi, tm_lo, p);
    else k_enc_bwd_tc<2><<<dim3(DAD_D / 256, splits), BWD_THREADS, BwdCfg<2>::SMEM, st>>>(tm_hi, tm_lo, p);
    DAD_CHECK_LAUNCH("k_enc_bwd_tc");
    if (forked && !side_join(lane, st)) return dad_set_error(DAD_ERR_CUDA, "encoder_bwd: helper-stream join failed: %s", cudaGetErrorString(cudaGetLastError()));
    g_bwd_lane ^= 1;  // consecutive calls (clean view, strong view) alternate between the two helper lanes
    return DAD_OK;
}
