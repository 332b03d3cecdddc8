// Encoder kernels on the 5th-gen tensor cores (tcgen05, accumulators in TMEM, weights fed by TMA).
//
// Arithmetic: every fp32 operand is split on the fly into bf16 hi + bf16 lo (hi = bf16(x), lo = bf16(x - hi)) and the
// product is formed as  hi*hi + hi*lo + lo*hi  with fp32 accumulation in TMEM ("bf16x3", ~2^-17 relative per product;
// parity evidence in DESIGN.md).  Only the dense contractions run here; everything else stays on CUDA cores.
//
// Forward  (I/model.py:18-41):  pre[N,256] = aug(x)[N,768] . W^T,  emb[b] = sum_t valid relu(pre + bias) / len_b
//   persistent CTAs (one per SM), 128-frame x 256-unit tiles, K loop of 12 x 64.  Warp roles:
//     warp 0      TMA producer: W_hi / W_lo k-blocks (bf16, SWIZZLE_128B) -> smem ring
//     warp 1      MMA issuer : 12 tcgen05.mma (3 products x 4 k-steps) per k-block into one of two TMEM accumulators
//     warp 2      TMEM allocator
//     warps 4-11  converters : coalesced LDG of fp32 x (+ fused weak/strong augmentation, I/utils.py:328-375),
//                              bf16 hi/lo split, swizzled st.shared of the A operand (K-major, SWIZZLE_128B)
//     warps 12-15 epilogue   : tcgen05.ld -> bias + ReLU + padding mask -> 1-bit ReLU mask + per-32-frame pooled partial sums
//   k_pool_finalize then sums the partials of each utterance in fixed order and divides by its valid length.
#include "dad_common.cuh"
#include "tc_ptx.cuh"
#include <cuda_bf16.h>
#include <mutex>

namespace {

constexpr int BM = 128, BK = 64, KB = DAD_D / BK, STAGES = 2;
constexpr int FWD_THREADS = 512;
constexpr int CONV_WARP0 = 4, N_CONV_WARPS = 8, EPI_WARP0 = 12;
constexpr uint32_t A_BYTES = BM * BK * 2;        // 16 KB  (one of hi / lo)
constexpr uint32_t W_BYTES = DAD_H * BK * 2;     // 32 KB
constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;  // 96 KB
constexpr uint32_t OFF_AHI = 0, OFF_ALO = A_BYTES, OFF_WHI = 2 * A_BYTES, OFF_WLO = 2 * A_BYTES + W_BYTES;
constexpr uint32_t OFF_SCRATCH = STAGES * STAGE_BYTES;                 // 4 warps x [32][33] f32
constexpr uint32_t OFF_BIAS = OFF_SCRATCH + 4 * 32 * 33 * 4;
constexpr uint32_t OFF_BARS = OFF_BIAS + DAD_H * 4;
constexpr uint32_t FWD_SMEM = OFF_BARS + 128 + 1024;                   // + alignment slack
constexpr uint32_t TMEM_COLS = 512;                                     // two 128 x 256 fp32 accumulators

struct FwdParams {
    const float* x;
    const uint8_t* pad;
    const float* bias;
    AugParams ap;
    float* part;        // [num_groups][2][256]
    uint32_t* bits;     // [N][8] or null
    int64_t N;
    int T;
    int num_tiles;
};

__global__ void __launch_bounds__(FWD_THREADS, 1)
k_enc_fwd_tc(const __grid_constant__ CUtensorMap tm_whi, const __grid_constant__ CUtensorMap tm_wlo, FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
    uint64_t* full = bars;            // [STAGES]  1 (TMA expect_tx) + 8 (converter warps)
    uint64_t* empty = bars + 2;       // [STAGES]  1 (tcgen05.commit)
    uint64_t* tfull = bars + 4;       // [2]       1 (tcgen05.commit)
    uint64_t* tempty = bars + 6;      // [2]       4 (epilogue warps)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
    float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < DAD_H) s_bias[threadIdx.x] = p.bias[threadIdx.x];
    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tm_whi);
        tc::prefetch_tmap(&tm_wlo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1 + N_CONV_WARPS); tc::mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull[a], 1); tc::mbar_init(&tempty[a], 4); }
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
        // ===================== TMA producer (weights) =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    tc::mbar_wait(&empty[s], ph ^ 1);
                    tc::mbar_arrive_expect_tx(&full[s], 2 * W_BYTES);
                    tc::tma_load_2d(smem + s * STAGE_BYTES + OFF_WHI, &tm_whi, &full[s], kb * BK, 0);
                    tc::tma_load_2d(smem + s * STAGE_BYTES + OFF_WLO, &tm_wlo, &full[s], kb * BK, 0);
                }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = tc::make_idesc_bf16(BM, DAD_H, 0, 0);
        uint32_t it = 0, ti = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++ti) {
            const uint32_t acc = ti & 1, tph = (ti >> 1) & 1;
            tc::mbar_wait(&tempty[acc], tph ^ 1);
            tc::tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * DAD_H;
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                tc::mbar_wait(&full[s], ph);
                tc::tcgen05_fence_after();
                if (lane == 0) {
                    const uint32_t base = tc::smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
                    for (int prod = 0; prod < 3; ++prod) {
                        const uint32_t a_off = (prod == 2) ? OFF_ALO : OFF_AHI;  // hi*hi, hi*lo, lo*hi
                        const uint32_t b_off = (prod == 1) ? OFF_WLO : OFF_WHI;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t ad = tc::make_smem_desc(base + a_off + k * 32, 16, 1024);
                            const uint64_t bd = tc::make_smem_desc(base + b_off + k * 32, 16, 1024);
                            tc::umma_bf16(d_tmem, ad, bd, idesc, (kb | prod | k) != 0);
                        }
                    }
                    tc::umma_commit(&empty[s]);                  // smem slot free once these MMAs retire
                    if (kb == KB - 1) tc::umma_commit(&tfull[acc]);  // accumulator complete
                }
                __syncwarp();
            }
        }
    } else if (warp >= CONV_WARP0 && warp < CONV_WARP0 + N_CONV_WARPS) {
        // ===================== converters: fp32 x (+ augmentation) -> bf16 hi/lo A operand =====================
        const int cw = warp - CONV_WARP0;
        const int kq = lane & 15, rsub = lane >> 4;  // float4 index inside the 64-float k-block row segment; row parity
        const AugParams& ap = p.ap;
        const unsigned long long seed = (ap.mode && ap.noise_src == 1) ? aug_seed(ap) : 0ull;
        const int mode = ap.mode;
        uint32_t it = 0;
        float4 cur[8], nxt[8];
        uint32_t zero_cur = 0, zero_nxt = 0;  // bit i: row i of this thread is out of range (or temporally masked) -> contributes zeros

        auto row_flags = [&](int tile) -> uint32_t {
            uint32_t z = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t n = (int64_t)tile * BM + cw * 16 + 2 * i + rsub;
                bool zr = n >= p.N;
                if (!zr && mode == 2 && ap.t_start && ap.mask_len > 0) {
                    const int b = (int)(n / p.T), t = (int)(n - (int64_t)b * p.T);
                    const int s = __ldg(ap.t_start + b);
                    zr = (t >= s && t < s + ap.mask_len);
                }
                z |= (zr ? 1u : 0u) << i;
            }
            return z;
        };
        auto issue_loads = [&](int tile, int kb, uint32_t zmask, float4 (&dst)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t n = (int64_t)tile * BM + cw * 16 + 2 * i + rsub;
                if (n < p.N && !((zmask >> i) & 1u))
                    dst[i] = __ldg(reinterpret_cast<const float4*>(p.x + n * DAD_D + kb * BK + kq * 4));
                else
                    dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };

        int tile = blockIdx.x;
        if (tile < p.num_tiles) {
            zero_cur = row_flags(tile);
            issue_loads(tile, 0, zero_cur, cur);
        }
        for (; tile < p.num_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < KB; ++kb, ++it) {
                // prefetch the next k-block (possibly of the next tile) into registers before converting this one
                int ntile = tile, nkb = kb + 1;
                if (nkb == KB) { nkb = 0; ntile = tile + gridDim.x; }
                const bool has_next = ntile < p.num_tiles;
                if (has_next) {
                    zero_nxt = (nkb == 0) ? row_flags(ntile) : zero_cur;
                    issue_loads(ntile, nkb, zero_nxt, nxt);
                }
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                tc::mbar_wait(&empty[s], ph ^ 1);
                uint8_t* a_hi = smem + s * STAGE_BYTES + OFF_AHI;
                uint8_t* a_lo = smem + s * STAGE_BYTES + OFF_ALO;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = cw * 16 + 2 * i + rsub;
                    float4 v = cur[i];
                    const bool zr = (zero_cur >> i) & 1u;
                    if (mode != 0 && !zr) {
                        const int64_t n = (int64_t)tile * BM + r;
                        const int64_t e = n * DAD_D + kb * BK + kq * 4;
                        float4 nz;
                        if (ap.noise_src == 1) nz = philox_normal4(seed, ap.stream, (unsigned long long)(e >> 2));
                        else nz = __ldg(reinterpret_cast<const float4*>(ap.noise + e));
                        v.x = aug_add_noise(v.x, nz.x, ap.sigma); v.y = aug_add_noise(v.y, nz.y, ap.sigma);
                        v.z = aug_add_noise(v.z, nz.z, ap.sigma); v.w = aug_add_noise(v.w, nz.w, ap.sigma);
                        if (mode == 2 && ap.feat_keep) {
                            const float4 kp = __ldg(reinterpret_cast<const float4*>(ap.feat_keep + kb * BK + kq * 4));
                            v.x = __fmul_rn(v.x, kp.x); v.y = __fmul_rn(v.y, kp.y); v.z = __fmul_rn(v.z, kp.z); v.w = __fmul_rn(v.w, kp.w);
                        }
                    }
                    if (zr) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    uint32_t h01, l01, h23, l23;
                    tc::split2_bf16(v.x, v.y, h01, l01);
                    tc::split2_bf16(v.z, v.w, h23, l23);
                    const uint32_t off = tc::sw128_offset(r, kq >> 1) + (kq & 1) * 8;
                    *reinterpret_cast<uint2*>(a_hi + off) = make_uint2(h01, h23);
                    *reinterpret_cast<uint2*>(a_lo + off) = make_uint2(l01, l23);
                }
                tc::fence_proxy_async_smem();  // generic-proxy stores -> visible to tcgen05.mma (async proxy)
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&full[s]);
                if (has_next) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
                    zero_cur = zero_nxt;
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===================== epilogue: TMEM -> bias + ReLU + mask -> ReLU bits + pooled partial sums =====================
        const int q = warp - EPI_WARP0;  // == warp % 4: this warp may touch TMEM lanes [32q, 32q+32)
        float* scratch = reinterpret_cast<float*>(smem + OFF_SCRATCH) + q * 32 * 33;
        uint32_t ti = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++ti) {
            const uint32_t acc = ti & 1, tph = (ti >> 1) & 1;
            const int64_t n0 = (int64_t)tile * BM + q * 32;
            const int64_t n = n0 + lane;
            const bool in = n < p.N;
            const bool valid = in && !(p.pad && p.pad[n]);
            const int64_t b_first = n0 / p.T;
            int64_t split64 = (b_first + 1) * p.T - n0;  // rows [0,split) belong to utterance b_first, the rest to b_first+1 (T >= 32)
            const int split = split64 > 32 ? 32 : (int)split64;
            tc::mbar_wait(&tfull[acc], tph);
            tc::tcgen05_fence_after();
            uint32_t words[8];
            const int64_t g = (int64_t)tile * 4 + q;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t r[32];
                tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * DAD_H + c * 32, r);
                tc::tmem_ld_wait();
                uint32_t w = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float pre = __uint_as_float(r[j]) + s_bias[c * 32 + j];
                    const bool pos = pre > 0.f;
                    w |= (pos ? 1u : 0u) << j;
                    scratch[lane * 33 + j] = (pos && valid) ? pre : 0.f;
                }
                words[c] = w;
                __syncwarp();
                float s0 = 0.f, s1 = 0.f;
                for (int rr = 0; rr < split; ++rr) s0 += scratch[rr * 33 + lane];
                for (int rr = split; rr < 32; ++rr) s1 += scratch[rr * 33 + lane];
                p.part[(g * 2 + 0) * DAD_H + c * 32 + lane] = s0;
                p.part[(g * 2 + 1) * DAD_H + c * 32 + lane] = s1;
                __syncwarp();
            }
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[acc]);
            if (p.bits && in) {
                uint4* dst = reinterpret_cast<uint4*>(p.bits + n * DAD_HW);
                dst[0] = make_uint4(words[0], words[1], words[2], words[3]);
                dst[1] = make_uint4(words[4], words[5], words[6], words[7]);
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// emb[b,:] = (sum of the 32-frame partials of utterance b, in frame order) / max(len_b, 1);  grid = B, 256 threads
__global__ void __launch_bounds__(DAD_H) k_pool_finalize(const float* __restrict__ part, const uint8_t* __restrict__ pad, float* __restrict__ emb,
                                                         float* __restrict__ lens_out, int T) {
    __shared__ float s_w[DAD_H / 32];
    __shared__ float s_len;
    const int b = blockIdx.x, h = threadIdx.x;
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
    const int64_t r0 = (int64_t)b * T, r1 = r0 + T - 1;
    float s = 0.f;
    for (int64_t g = r0 / 32; g <= r1 / 32; ++g) {
        const int slot = ((g * 32) / T == b) ? 0 : 1;
        s += __ldg(part + (g * 2 + slot) * DAD_H + h);
    }
    emb[(int64_t)b * DAD_H + h] = s / s_len;
    if (h == 0 && lens_out) lens_out[b] = s_len;
}

// ---------------------------------------------------------------- host: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(f);
    });
    return fn;
}

// bf16 [rows][cols] row-major, box = [box_rows][64 cols] (128 bytes inner), SWIZZLE_128B
int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return dad_set_error(DAD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return dad_set_error(DAD_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return DAD_OK;
}

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace

size_t enc_fwd_tc_ws_bytes(int64_t B, int64_t T) {
    const int64_t tiles = ceil_div64(B * T, BM);
    return (size_t)(tiles * 4 * 2 * DAD_H * sizeof(float));
}

int launch_enc_fwd_tc(const float* x, const uint8_t* pad, const void* W_hi, const void* W_lo, const float* bias, const AugParams& ap, float* emb,
                      uint32_t* bits, float* lens_out, void* ws, size_t ws_bytes, int64_t B, int64_t T, cudaStream_t st) {
    if (T < 32) return dad_set_error(DAD_ERR_UNSUPPORTED, "encoder_fwd: the tcgen05 path needs T >= 32 (got %lld); use DAD_GEMM_FP32", (long long)T);
    if ((reinterpret_cast<uintptr_t>(W_hi) | reinterpret_cast<uintptr_t>(W_lo)) & 15u)
        return dad_set_error(DAD_ERR_ARG, "encoder_fwd: W_hi/W_lo must be 16-byte aligned (TMA)");
    if (ws_bytes < enc_fwd_tc_ws_bytes(B, T)) return dad_set_error(DAD_ERR_WS, "encoder_fwd(tc): workspace too small");
    CUtensorMap tm_hi, tm_lo;
    int rc = make_tmap_bf16(&tm_hi, W_hi, DAD_H, DAD_D, DAD_H);
    if (rc) return rc;
    rc = make_tmap_bf16(&tm_lo, W_lo, DAD_H, DAD_D, DAD_H);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_enc_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
        if (e != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "encoder_fwd(tc): cannot opt in to %u B shared memory: %s", FWD_SMEM, cudaGetErrorString(e));
        attr = true;
    }
    FwdParams p;
    p.x = x; p.pad = pad; p.bias = bias; p.ap = ap; p.part = reinterpret_cast<float*>(ws); p.bits = bits;
    p.N = B * T; p.T = (int)T; p.num_tiles = (int)ceil_div64(B * T, BM);
    const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    k_enc_fwd_tc<<<grid, FWD_THREADS, FWD_SMEM, st>>>(tm_hi, tm_lo, p);
    DAD_CHECK_LAUNCH("k_enc_fwd_tc");
    k_pool_finalize<<<(unsigned)B, DAD_H, 0, st>>>(p.part, pad, emb, lens_out, (int)T);
    DAD_CHECK_LAUNCH("k_pool_finalize");
    return DAD_OK;
}

// =====================================================================================================================
// Backward:  dW[256,768] = sum_n dpre[n,:] (x) aug(x)[n,:],   dpre[n,h] = relu_bit(n,h) * valid(n) * g[b(n),h] / len_b
//   = a [256 x N] . [N x 768] GEMM whose reduction axis is the frame axis: BOTH operands are MN-major in shared memory
//   (frame rows, 128-byte lines of 64 contiguous h / d elements, SWIZZLE_128B), so nothing is transposed.
//   grid = (3 d-slices of 256, S frame-range splits); per CTA two 128 x 256 fp32 accumulators (all 512 TMEM columns).
//   The A operand is never read from memory as numbers: it is SELECTED from a per-utterance table of pre-split
//   g/len values (G_hi/G_lo, bf16) by the 1-bit ReLU mask of the forward.  Split partials are reduced in fixed order
//   by dad_grad_reduce (deterministic, no fp32 atomics).
// =====================================================================================================================
namespace {

constexpr int BWD_BK = 32, BWD_STAGES = 3, BWD_THREADS = 512;
constexpr uint32_t BOP_BYTES = DAD_H * BWD_BK * 2;                  // 16 KB: one operand (hi or lo), 256 x 32 bf16
constexpr uint32_t BSTAGE_BYTES = 4 * BOP_BYTES;                    // A_hi, A_lo, B_hi, B_lo = 64 KB
constexpr uint32_t BOFF_AHI = 0, BOFF_ALO = BOP_BYTES, BOFF_BHI = 2 * BOP_BYTES, BOFF_BLO = 3 * BOP_BYTES;
constexpr uint32_t BOFF_BARS = BWD_STAGES * BSTAGE_BYTES;
constexpr uint32_t BWD_SMEM = BOFF_BARS + 128 + 1024;
constexpr uint32_t MN_LBO = (BWD_BK / 8) * 1024;                    // pitch between 64-element M/N chunks (4 k-groups of 1 KB)
constexpr uint32_t MN_SBO = 1024;                                   // pitch between 8-frame groups

struct BwdParams {
    const float* x;
    const uint8_t* pad;
    AugParams ap;
    const uint32_t* bits;            // [N][8]
    const __nv_bfloat16* g_hi;       // [B][256]  bf16(g/len)
    const __nv_bfloat16* g_lo;       // [B][256]  bf16(g/len - hi)
    float* dWp;                      // [S][256][768]
    int64_t N;
    int T;
    int64_t rows_per_split;          // multiple of 32
};

// G_hi/G_lo = bf16 split of g_emb / len (division as in autograd of sum/len), one thread per (b,h)
__global__ void __launch_bounds__(DAD_H) k_bwd_prep(const float* __restrict__ g_emb, const float* __restrict__ lens, __nv_bfloat16* __restrict__ g_hi,
                                                    __nv_bfloat16* __restrict__ g_lo) {
    const int64_t i = (int64_t)blockIdx.x * DAD_H + threadIdx.x;
    const float v = g_emb[i] / lens[blockIdx.x];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    g_hi[i] = h;
    g_lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// db[h]: sum over valid frames of relu_bit * g/len (fp32, exact selection).  grid = (S, sub): CTA (s, u) covers the u-th
// slice of split s's frame range with 4 independent frames in flight per thread; k_bwd_db_fold sums the sub-partials.
__global__ void __launch_bounds__(DAD_H) k_bwd_db(const uint32_t* __restrict__ bits, const uint8_t* __restrict__ pad, const float* __restrict__ g_emb,
                                                  const float* __restrict__ lens, float* __restrict__ sub_part, int64_t N, int T,
                                                  int64_t rows_per_split, int64_t rows_per_sub) {
    const int h = threadIdx.x;
    const int64_t s0 = (int64_t)blockIdx.x * rows_per_split;
    const int64_t s1 = s0 + rows_per_split < N ? s0 + rows_per_split : N;
    int64_t n0 = s0 + (int64_t)blockIdx.y * rows_per_sub;
    int64_t n1 = n0 + rows_per_sub < s1 ? n0 + rows_per_sub : s1;
    float acc = 0.f;
    for (int64_t n = n0; n < n1; n += 4) {
        uint32_t w[4];
        bool on[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t m = n + u;
            on[u] = m < n1 && !(pad && pad[m]);
            w[u] = on[u] ? __ldg(bits + m * DAD_HW + (h >> 5)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (on[u] && ((w[u] >> (h & 31)) & 1u)) {
                const int64_t b = (n + u) / T;
                acc += __ldg(g_emb + b * DAD_H + h) / __ldg(lens + b);
            }
        }
    }
    sub_part[((int64_t)blockIdx.x * gridDim.y + blockIdx.y) * DAD_H + h] = acc;
}
__global__ void __launch_bounds__(DAD_H) k_bwd_db_fold(const float* __restrict__ sub_part, float* __restrict__ dbp, int nsub) {
    float s = 0.f;
    for (int u = 0; u < nsub; ++u) s += sub_part[((int64_t)blockIdx.x * nsub + u) * DAD_H + threadIdx.x];
    dbp[(int64_t)blockIdx.x * DAD_H + threadIdx.x] = s;
}

__global__ void __launch_bounds__(BWD_THREADS, 1) k_enc_bwd_tc(BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BOFF_BARS);
    uint64_t* full = bars;                  // [3] 8 converter warps
    uint64_t* empty = bars + BWD_STAGES;    // [3] 1 (tcgen05.commit)
    uint64_t* tfull = bars + 2 * BWD_STAGES;  // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * BWD_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dslice = blockIdx.x, split = blockIdx.y;
    const int64_t n_begin = (int64_t)split * p.rows_per_split;
    const int64_t n_end = n_begin + p.rows_per_split < p.N ? n_begin + p.rows_per_split : p.N;
    const int nkb = n_end > n_begin ? (int)ceil_div64(n_end - n_begin, BWD_BK) : 0;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < BWD_STAGES; ++s) { tc::mbar_init(&full[s], N_CONV_WARPS); tc::mbar_init(&empty[s], 1); }
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

    if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = tc::make_idesc_bf16(128, DAD_H, 1, 1);  // A and B both MN-major
        for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t s = kb % BWD_STAGES, ph = (kb / BWD_STAGES) & 1;
            tc::mbar_wait(&full[s], ph);
            tc::tcgen05_fence_after();
            if (lane == 0) {
                const uint32_t base = tc::smem_u32(smem + s * BSTAGE_BYTES);
#pragma unroll
                for (int mh = 0; mh < 2; ++mh)
#pragma unroll
                    for (int prod = 0; prod < 3; ++prod) {
                        const uint32_t a_off = ((prod == 2) ? BOFF_ALO : BOFF_AHI) + mh * 2 * MN_LBO;  // hi*hi, hi*lo, lo*hi
                        const uint32_t b_off = (prod == 1) ? BOFF_BLO : BOFF_BHI;
#pragma unroll
                        for (int ks = 0; ks < BWD_BK / 16; ++ks) {
                            const uint64_t ad = tc::make_smem_desc(base + a_off + ks * 2 * MN_SBO, MN_LBO, MN_SBO);
                            const uint64_t bd = tc::make_smem_desc(base + b_off + ks * 2 * MN_SBO, MN_LBO, MN_SBO);
                            tc::umma_bf16(tmem_base + mh * DAD_H, ad, bd, idesc, (kb | prod | ks) != 0);
                        }
                    }
                tc::umma_commit(&empty[s]);
                if (kb == nkb - 1) tc::umma_commit(tfull);
            }
            __syncwarp();
        }
    } else if (warp >= CONV_WARP0 && warp < CONV_WARP0 + N_CONV_WARPS) {
        // ===================== converters =====================
        const int cw = warp - CONV_WARP0;
        const int ctid = threadIdx.x - CONV_WARP0 * 32;  // 0..255
        const AugParams& ap = p.ap;
        const unsigned long long seed = (ap.mode && ap.noise_src == 1) ? aug_seed(ap) : 0ull;
        const int mode = ap.mode;
        // x operand: this thread converts float4 #f (f = half*32 + lane) of rows cw*4 + i/2, i = 0..7
        // A operand: this thread generates the 16-byte chunks hq = hc + 8j (j = 0..3) of frame row ar
        const int ar = ctid >> 3, hc = ctid & 7;
        float4 cur[8], nxt[8];
        uint32_t zc = 0, zn = 0;
        uint32_t wcur[4], wnxt[4];
        int64_t bcur = 0, bnxt = 0;

        auto x_flags = [&](int kb) -> uint32_t {
            uint32_t z = 0;
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) {
                const int64_t n = n_begin + (int64_t)kb * BWD_BK + cw * 4 + r4;
                bool zr = n >= n_end;
                if (!zr && mode == 2 && ap.t_start && ap.mask_len > 0) {
                    const int b = (int)(n / p.T), t = (int)(n - (int64_t)b * p.T);
                    const int s = __ldg(ap.t_start + b);
                    zr = (t >= s && t < s + ap.mask_len);
                }
                z |= (zr ? 1u : 0u) << r4;
            }
            return z;
        };
        auto x_loads = [&](int kb, uint32_t z, float4 (&dst)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r4 = i >> 1, f = (i & 1) * 32 + lane;
                const int64_t n = n_begin + (int64_t)kb * BWD_BK + cw * 4 + r4;
                if (n < n_end && !((z >> r4) & 1u)) dst[i] = __ldg(reinterpret_cast<const float4*>(p.x + n * DAD_D + dslice * 256 + f * 4));
                else dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto a_loads = [&](int kb, uint32_t (&w)[4], int64_t& b) {
            const int64_t n = n_begin + (int64_t)kb * BWD_BK + ar;
            const bool on = n < n_end && !(p.pad && p.pad[n]);
            b = n < n_end ? n / p.T : 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int h = (hc + 8 * j) * 8;
                const uint32_t word = on ? __ldg(p.bits + n * DAD_HW + (h >> 5)) : 0u;
                w[j] = (word >> (h & 31)) & 0xFFu;
            }
        };

        if (nkb > 0) {
            zc = x_flags(0);
            x_loads(0, zc, cur);
            a_loads(0, wcur, bcur);
        }
        for (int kb = 0; kb < nkb; ++kb) {
            const bool has_next = kb + 1 < nkb;
            if (has_next) {
                zn = x_flags(kb + 1);
                x_loads(kb + 1, zn, nxt);
                a_loads(kb + 1, wnxt, bnxt);
            }
            const uint32_t s = kb % BWD_STAGES, ph = (kb / BWD_STAGES) & 1;
            tc::mbar_wait(&empty[s], ph ^ 1);
            uint8_t* st = smem + s * BSTAGE_BYTES;
            // ---- A operand: dpre^T chunks selected from the pre-split g/len table by the ReLU bits
            {
                const uint32_t rofs = (uint32_t)(ar >> 3) * MN_SBO;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int hq = hc + 8 * j;  // 16-byte chunk index along h (0..31)
                    const uint32_t byte = wcur[j];
                    uint4 gh = make_uint4(0, 0, 0, 0), gl = make_uint4(0, 0, 0, 0);
                    if (byte) {
                        gh = __ldg(reinterpret_cast<const uint4*>(p.g_hi + bcur * DAD_H + hq * 8));
                        gl = __ldg(reinterpret_cast<const uint4*>(p.g_lo + bcur * DAD_H + hq * 8));
                        const uint32_t m0 = ((byte & 1u) ? 0x0000FFFFu : 0u) | ((byte & 2u) ? 0xFFFF0000u : 0u);
                        const uint32_t m1 = ((byte & 4u) ? 0x0000FFFFu : 0u) | ((byte & 8u) ? 0xFFFF0000u : 0u);
                        const uint32_t m2 = ((byte & 16u) ? 0x0000FFFFu : 0u) | ((byte & 32u) ? 0xFFFF0000u : 0u);
                        const uint32_t m3 = ((byte & 64u) ? 0x0000FFFFu : 0u) | ((byte & 128u) ? 0xFFFF0000u : 0u);
                        gh.x &= m0; gh.y &= m1; gh.z &= m2; gh.w &= m3;
                        gl.x &= m0; gl.y &= m1; gl.z &= m2; gl.w &= m3;
                    }
                    const uint32_t off = (uint32_t)(hq >> 3) * MN_LBO + rofs + tc::sw128_offset(ar & 7, hq & 7);
                    *reinterpret_cast<uint4*>(st + BOFF_AHI + off) = gh;
                    *reinterpret_cast<uint4*>(st + BOFF_ALO + off) = gl;
                }
            }
            // ---- B operand: aug(x) rows -> bf16 hi/lo
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r4 = i >> 1, f = (i & 1) * 32 + lane;
                const int r = cw * 4 + r4;
                float4 v = cur[i];
                const bool zr = (zc >> r4) & 1u;
                if (mode != 0 && !zr) {
                    const int64_t n = n_begin + (int64_t)kb * BWD_BK + r;
                    const int d = dslice * 256 + f * 4;
                    const int64_t e = n * DAD_D + d;
                    float4 nz;
                    if (ap.noise_src == 1) nz = philox_normal4(seed, ap.stream, (unsigned long long)(e >> 2));
                    else nz = __ldg(reinterpret_cast<const float4*>(ap.noise + e));
                    v.x = aug_add_noise(v.x, nz.x, ap.sigma); v.y = aug_add_noise(v.y, nz.y, ap.sigma);
                    v.z = aug_add_noise(v.z, nz.z, ap.sigma); v.w = aug_add_noise(v.w, nz.w, ap.sigma);
                    if (mode == 2 && ap.feat_keep) {
                        const float4 kp = __ldg(reinterpret_cast<const float4*>(ap.feat_keep + d));
                        v.x = __fmul_rn(v.x, kp.x); v.y = __fmul_rn(v.y, kp.y); v.z = __fmul_rn(v.z, kp.z); v.w = __fmul_rn(v.w, kp.w);
                    }
                }
                if (zr) v = make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t h01, l01, h23, l23;
                tc::split2_bf16(v.x, v.y, h01, l01);
                tc::split2_bf16(v.z, v.w, h23, l23);
                const uint32_t off = (uint32_t)(f >> 4) * MN_LBO + (uint32_t)(r >> 3) * MN_SBO + tc::sw128_offset(r & 7, (f & 15) >> 1) + (f & 1) * 8;
                *reinterpret_cast<uint2*>(st + BOFF_BHI + off) = make_uint2(h01, h23);
                *reinterpret_cast<uint2*>(st + BOFF_BLO + off) = make_uint2(l01, l23);
            }
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full[s]);
            if (has_next) {
#pragma unroll
                for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) wcur[j] = wnxt[j];
                zc = zn;
                bcur = bnxt;
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===================== epilogue: TMEM -> split-K partial dW =====================
        const int q = warp - EPI_WARP0;
        float* out = p.dWp + (int64_t)split * DAD_H * DAD_D;
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
                    *reinterpret_cast<uint4*>(row + c * 32 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

int enc_bwd_tc_splits(int64_t N) {
    // grid = 3 d-slices x S: one wave of 147 CTAs at S = 49; keep >= 128 frames per split for small batches
    int64_t s = N / 128;
    if (s > 49) s = 49;
    if (s < 1) s = 1;
    return (int)s;
}
constexpr int DB_MAX_SUB = 32;
size_t enc_bwd_tc_ws_bytes(int64_t B, int64_t) {
    return (size_t)(2 * B * DAD_H * sizeof(__nv_bfloat16)) + 256 + (size_t)49 * DB_MAX_SUB * DAD_H * sizeof(float);
}

int launch_enc_bwd_tc(const float* x, const uint8_t* pad, const AugParams& ap, const uint32_t* bits, const float* g_emb, const float* lens,
                      float* dWp, float* dbp, int splits, void* ws, size_t ws_bytes, int64_t B, int64_t T, cudaStream_t st) {
    if (ws_bytes < enc_bwd_tc_ws_bytes(B, T)) return dad_set_error(DAD_ERR_WS, "encoder_bwd(tc): workspace too small");
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_enc_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
        if (e != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "encoder_bwd(tc): cannot opt in to %u B shared memory: %s", BWD_SMEM, cudaGetErrorString(e));
        attr = true;
    }
    __nv_bfloat16* g_hi = reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* g_lo = g_hi + B * DAD_H;
    k_bwd_prep<<<(unsigned)B, DAD_H, 0, st>>>(g_emb, lens, g_hi, g_lo);
    DAD_CHECK_LAUNCH("k_bwd_prep");
    BwdParams p;
    p.x = x; p.pad = pad; p.ap = ap; p.bits = bits; p.g_hi = g_hi; p.g_lo = g_lo; p.dWp = dWp;
    p.N = B * T; p.T = (int)T;
    p.rows_per_split = ceil_div64(ceil_div64(p.N, splits), BWD_BK) * BWD_BK;
    k_enc_bwd_tc<<<dim3(DAD_D / 256, splits), BWD_THREADS, BWD_SMEM, st>>>(p);
    DAD_CHECK_LAUNCH("k_enc_bwd_tc");
    float* sub_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + (((size_t)(2 * B * DAD_H * sizeof(__nv_bfloat16)) + 255) & ~(size_t)255));
    int nsub = (int)(p.rows_per_split / 64);
    nsub = nsub < 1 ? 1 : (nsub > DB_MAX_SUB ? DB_MAX_SUB : nsub);
    const int64_t rows_per_sub = ceil_div64(p.rows_per_split, nsub);
    k_bwd_db<<<dim3(splits, nsub), DAD_H, 0, st>>>(bits, pad, g_emb, lens, sub_part, p.N, (int)T, p.rows_per_split, rows_per_sub);
    DAD_CHECK_LAUNCH("k_bwd_db");
    k_bwd_db_fold<<<splits, DAD_H, 0, st>>>(sub_part, dbp, nsub);
    DAD_CHECK_LAUNCH("k_bwd_db_fold");
    return DAD_OK;
}
