// ECDA pairwise Gram matrices on the 5th-gen tensor cores (tcgen05, accumulators in TMEM, both operands fed by TMA).
//
// G[i][j] = z'_i . z'_j for the class-centred member embeddings z' = z - mean_c (the squared distances of the multi-kernel
// Gaussian MMD, I/utils.py:532-541, are then |z'_i|^2 + |z'_j|^2 - 2 G_ij -- SURVEY.md Appendix A "one Gram GEMM").
// Same f16x3 arithmetic as the encoder kernels (k_enc_tc.cu): z' is scaled by 2^8 and split into fp16 hi + lo by k_ecda_pack
// (22 significant bits), the product is hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM, the epilogue unscales by 2^-16.
//
//   work item  = one 128-row x 256-column block of one class (rows this rank owns x all members of the class), K = 256;
//   persistent CTAs (one per SM), items strided over the grid; per item 4 k-blocks of 64; 256 threads:
//     warp 0      TMA producer : A (128 x 64) and B (256 x 64) tiles of Z_hi / Z_lo, SWIZZLE_128B, 2-stage ring (96 KB / stage)
//     warp 1      MMA issuer  : 12 tcgen05.mma (3 products x 4 k-steps) per k-block into one of two 128 x 256 TMEM accumulators
//     warp 2      TMEM allocator
//     warps 4-7   epilogue    : tcgen05.ld -> exact 2^-16 unscale -> G (fp32, row-major), overlapped with the next item's MMAs
#include "dad_common.cuh"
#include "dad_pack.cuh"
#include "tc_ptx.cuh"
#include "tc_host.cuh"
#include <string.h>

namespace {

constexpr int GBM = 128, GBN = 256, GBK = 64, GKB = DAD_H / GBK, GSTAGES = 2;
constexpr uint32_t GA_BYTES = GBM * GBK * 2;   // 16 KB
constexpr uint32_t GB_BYTES = GBN * GBK * 2;   // 32 KB
constexpr uint32_t GSTAGE = 2 * GA_BYTES + 2 * GB_BYTES;  // 96 KB
constexpr uint32_t GOFF_AHI = 0, GOFF_ALO = GA_BYTES, GOFF_BHI = 2 * GA_BYTES, GOFF_BLO = 2 * GA_BYTES + GB_BYTES;
constexpr uint32_t GOFF_BARS = GSTAGES * GSTAGE;
constexpr uint32_t GRAM_SMEM = GOFF_BARS + 128 + 1024;
constexpr int GRAM_THREADS = 256;
constexpr float GRAM_UNSCALE = 1.0f / (ECDA_ZS * ECDA_ZS);

__global__ void __launch_bounds__(GRAM_THREADS, 1)
k_gram_tc(const __grid_constant__ CUtensorMap tm_ah, const __grid_constant__ CUtensorMap tm_al, const __grid_constant__ CUtensorMap tm_bh,
          const __grid_constant__ CUtensorMap tm_bl, const GramItem* __restrict__ items, const int* __restrict__ n_items_ptr, float* __restrict__ G,
          int64_t ldg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GOFF_BARS);
    uint64_t* full = bars;        // [2] 1 (TMA expect_tx)
    uint64_t* empty = bars + 2;   // [2] 1 (tcgen05.commit)
    uint64_t* tfull = bars + 4;   // [2] 1 (tcgen05.commit)
    uint64_t* tempty = bars + 6;  // [2] 4 (epilogue warps)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = *n_items_ptr;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tm_ah); tc::prefetch_tmap(&tm_al); tc::prefetch_tmap(&tm_bh); tc::prefetch_tmap(&tm_bl);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < GSTAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull[a], 1); tc::mbar_init(&tempty[a], 4); }
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_ptr, 512);
        tc::tmem_relinquish();
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
                const GramItem gi = items[w];
                for (int kb = 0; kb < GKB; ++kb, ++it) {
                    const uint32_t s = it % GSTAGES, ph = (it / GSTAGES) & 1;
                    tc::mbar_wait(&empty[s], ph ^ 1);
                    tc::mbar_arrive_expect_tx(&full[s], GSTAGE);
                    uint8_t* st = smem + s * GSTAGE;
                    tc::tma_load_2d(st + GOFF_AHI, &tm_ah, &full[s], kb * GBK, gi.a_row);
                    tc::tma_load_2d(st + GOFF_ALO, &tm_al, &full[s], kb * GBK, gi.a_row);
                    tc::tma_load_2d(st + GOFF_BHI, &tm_bh, &full[s], kb * GBK, gi.b_row);
                    tc::tma_load_2d(st + GOFF_BLO, &tm_bl, &full[s], kb * GBK, gi.b_row);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = tc::make_idesc_f16(GBM, GBN, 0, 0);
        uint32_t it = 0, ti = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++ti) {
            const uint32_t acc = ti & 1, tph = (ti >> 1) & 1;
            tc::mbar_wait(&tempty[acc], tph ^ 1);
            tc::tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * GBN;
            for (int kb = 0; kb < GKB; ++kb, ++it) {
                const uint32_t s = it % GSTAGES, ph = (it / GSTAGES) & 1;
                tc::mbar_wait(&full[s], ph);
                tc::tcgen05_fence_after();
                if (lane == 0) {
                    const uint32_t base = tc::smem_u32(smem + s * GSTAGE);
#pragma unroll
                    for (int prod = 0; prod < 3; ++prod) {
                        const uint32_t a_off = (prod == 2) ? GOFF_ALO : GOFF_AHI;  // hi*hi, hi*lo, lo*hi
                        const uint32_t b_off = (prod == 1) ? GOFF_BLO : GOFF_BHI;
#pragma unroll
                        for (int k = 0; k < GBK / 16; ++k) {
                            const uint64_t ad = tc::make_smem_desc(base + a_off + k * 32, 16, 1024);
                            const uint64_t bd = tc::make_smem_desc(base + b_off + k * 32, 16, 1024);
                            tc::umma_f16(d_tmem, ad, bd, idesc, (kb | prod | k) != 0);
                        }
                    }
                    tc::umma_commit(&empty[s]);
                    if (kb == GKB - 1) tc::umma_commit(&tfull[acc]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> G =====================
        const int q = warp - 4;  // == warp % 4: TMEM lanes [32q, 32q+32)
        uint32_t ti = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++ti) {
            const GramItem gi = items[w];
            const uint32_t acc = ti & 1, tph = (ti >> 1) & 1;
            tc::mbar_wait(&tfull[acc], tph);
            tc::tcgen05_fence_after();
            float* row = G + (int64_t)(gi.g_row + q * 32 + lane) * ldg + gi.g_col;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * GBN;
#pragma unroll 1
            for (int c = 0; c < GBN / 32; ++c) {
                uint32_t r[32];
                tc::tmem_ld_32x32(taddr + c * 32, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(row + c * 32 + j) = make_float4(__uint_as_float(r[j]) * GRAM_UNSCALE, __uint_as_float(r[j + 1]) * GRAM_UNSCALE,
                                                                                __uint_as_float(r[j + 2]) * GRAM_UNSCALE, __uint_as_float(r[j + 3]) * GRAM_UNSCALE);
            }
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[acc]);
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_pz_tc: the second ECDA GEMM, (P Z')[i][h] = sum_j P_ij z'_j[h], on tcgen05.  A = P rows (K-major: the member index j is the
// contiguous axis of P), B = Z' rows [j][h] = MN-major (h contiguous), so neither operand is transposed in memory -- the same
// operand-major trick as the encoder backward (k_enc_tc.cu).  One work item = 128 owned rows x all 256 features, K = the class's
// members in k-blocks of 64; P columns beyond the class are zeros (k_ecda_p writes whole 256-column blocks), so ragged class
// sizes need no tail handling.  Same warp roles / barriers / f16x3 products as k_gram_tc.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GRAM_THREADS, 1)
k_pz_tc(const __grid_constant__ CUtensorMap tm_ph, const __grid_constant__ CUtensorMap tm_pl, const __grid_constant__ CUtensorMap tm_zh,
        const __grid_constant__ CUtensorMap tm_zl, const PzItem* __restrict__ items, const int* __restrict__ n_items_ptr, float* __restrict__ accz,
        int g_rows_alloc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GOFF_BARS);
    uint64_t* full = bars;
    uint64_t* empty = bars + 2;
    uint64_t* tfull = bars + 4;
    uint64_t* tempty = bars + 6;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = *n_items_ptr;
    constexpr uint32_t B_CHUNK = GBK * 128;  // 8 KB: one 64-feature chunk of a 64-member k-block (MN-major B)

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tm_ph); tc::prefetch_tmap(&tm_pl); tc::prefetch_tmap(&tm_zh); tc::prefetch_tmap(&tm_zl);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < GSTAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull[a], 1); tc::mbar_init(&tempty[a], 4); }
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_ptr, 512);
        tc::tmem_relinquish();
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
                const PzItem pi = items[w];
                for (int kk = 0; kk < pi.n_kb; ++kk, ++it) {
                    const int kb = pi.kb0 + kk;
                    const uint32_t s = it % GSTAGES, ph = (it / GSTAGES) & 1;
                    tc::mbar_wait(&empty[s], ph ^ 1);
                    tc::mbar_arrive_expect_tx(&full[s], GSTAGE);
                    uint8_t* st = smem + s * GSTAGE;
                    tc::tma_load_2d(st + GOFF_AHI, &tm_ph, &full[s], kb * GBK, pi.g_row);
                    tc::tma_load_2d(st + GOFF_ALO, &tm_pl, &full[s], kb * GBK, pi.g_row);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        tc::tma_load_2d(st + GOFF_BHI + c * B_CHUNK, &tm_zh, &full[s], c * 64, pi.b_row + kb * GBK);
                        tc::tma_load_2d(st + GOFF_BLO + c * B_CHUNK, &tm_zl, &full[s], c * 64, pi.b_row + kb * GBK);
                    }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = tc::make_idesc_f16(GBM, GBN, 0, 1);  // A K-major, B MN-major
        uint32_t it = 0, ti = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++ti) {
            const int n_kb = items[w].n_kb;
            const uint32_t acc = ti & 1, tph = (ti >> 1) & 1;
            tc::mbar_wait(&tempty[acc], tph ^ 1);
            tc::tcgen05_fence_after();
            if (n_kb == 0) {  // empty split-K slice (class smaller than the slice count): nothing to accumulate, the epilogue writes zeros
                if (lane == 0) tc::umma_commit(&tfull[acc]);
                __syncwarp();
                continue;
            }
            const uint32_t d_tmem = tmem_base + acc * GBN;
            for (int kb = 0; kb < n_kb; ++kb, ++it) {
                const uint32_t s = it % GSTAGES, ph = (it / GSTAGES) & 1;
                tc::mbar_wait(&full[s], ph);
                tc::tcgen05_fence_after();
                if (lane == 0) {
                    const uint32_t base = tc::smem_u32(smem + s * GSTAGE);
#pragma unroll
                    for (int prod = 0; prod < 3; ++prod) {
                        const uint32_t a_off = (prod == 2) ? GOFF_ALO : GOFF_AHI;
                        const uint32_t b_off = (prod == 1) ? GOFF_BLO : GOFF_BHI;
#pragma unroll
                        for (int k = 0; k < GBK / 16; ++k) {
                            const uint64_t ad = tc::make_smem_desc(base + a_off + k * 32, 16, 1024);
                            const uint64_t bd = tc::make_smem_desc(base + b_off + k * 2048, B_CHUNK, 1024);  // LBO = chunk pitch, SBO = 8-member group
                            tc::umma_f16(d_tmem, ad, bd, idesc, (kb | prod | k) != 0);
                        }
                    }
                    tc::umma_commit(&empty[s]);
                    if (kb == n_kb - 1) tc::umma_commit(&tfull[acc]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        uint32_t ti = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++ti) {
            const PzItem pi = items[w];
            const uint32_t acc = ti & 1, tph = (ti >> 1) & 1;
            tc::mbar_wait(&tfull[acc], tph);
            tc::tcgen05_fence_after();
            float* row = accz + ((int64_t)pi.slice * g_rows_alloc + pi.g_row + q * 32 + lane) * DAD_H;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * GBN;
            const float unscale = pi.n_kb > 0 ? 1.0f / ECDA_ZS : 0.f;  // the P scale (per class) is removed by k_ecda_fold
#pragma unroll 1
            for (int c = 0; c < GBN / 32; ++c) {
                uint32_t r[32];
                if (pi.n_kb > 0) {
                    tc::tmem_ld_32x32(taddr + c * 32, r);
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
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[acc]);
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// Z_hi / Z_lo rows in LIST ORDER (class by class, clean members then noisy members): z' = z - mean of the row's class, scaled 2^8
__global__ void __launch_bounds__(256) k_ecda_pack(EcdaArgs a) {
    const EcdaStats* st = a.stats;
    if (!st->active || !st->any_feasible) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pos = blockIdx.x * 8 + warp;
    const int last_c = st->n_classes - 1;
    const int n_total = st->list_off[last_c] + st->n_s[last_c] + st->n_t[last_c];
    if (pos >= n_total) return;
    int c = 0;
    for (int cc = 1; cc <= last_c; ++cc)
        if (pos >= st->list_off[cc]) c = cc;
    const int Bg = a.pack.B_local * a.pack.world;
    const int r = a.lists[pos];
    const float* z = r < Bg ? pack_emb_clean(a.pack, r) : pack_emb_strong(a.pack, r - Bg);
    const float4 v0 = *reinterpret_cast<const float4*>(z + lane * 8), v1 = *reinterpret_cast<const float4*>(z + lane * 8 + 4);
    const float4 m0 = *reinterpret_cast<const float4*>(a.mu_all + c * DAD_H + lane * 8), m1 = *reinterpret_cast<const float4*>(a.mu_all + c * DAD_H + lane * 8 + 4);
    {
        const float d[8] = {v0.x - m0.x, v0.y - m0.y, v0.z - m0.z, v0.w - m0.w, v1.x - m1.x, v1.y - m1.y, v1.z - m1.z, v1.w - m1.w};
        float s2 = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) s2 = fmaf(d[u], d[u], s2);
        s2 = warp_sum(s2);
        if (lane == 0) {
            const int dom = r >= Bg ? 1 : 0;
            a.pos_nsq[pos] = s2;
            a.pos_om[pos] = (dom && a.class_aware) ? pack_score(a.pack, r - Bg) : 1.f;
            a.pos_dom[pos] = dom;
        }
    }
    uint32_t h[4], l[4];
    tc::split2_f16((v0.x - m0.x) * ECDA_ZS, (v0.y - m0.y) * ECDA_ZS, h[0], l[0]);
    tc::split2_f16((v0.z - m0.z) * ECDA_ZS, (v0.w - m0.w) * ECDA_ZS, h[1], l[1]);
    tc::split2_f16((v1.x - m1.x) * ECDA_ZS, (v1.y - m1.y) * ECDA_ZS, h[2], l[2]);
    tc::split2_f16((v1.z - m1.z) * ECDA_ZS, (v1.w - m1.w) * ECDA_ZS, h[3], l[3]);
    *reinterpret_cast<uint4*>(a.z_hi + (int64_t)pos * DAD_H + lane * 8) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(a.z_lo + (int64_t)pos * DAD_H + lane * 8) = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace

int launch_ecda_gram(const EcdaArgs& a, cudaStream_t st) {
    const int R = 2 * a.pack.B_local * a.pack.world;
    k_ecda_pack<<<ceil_div(R, 8), 256, 0, st>>>(a);
    DAD_CHECK_LAUNCH("k_ecda_pack");
    CUtensorMap ah, al, bh, bl;
    int rc = make_tmap_f16(&ah, a.z_hi, (uint64_t)R, DAD_H, GBM);
    if (!rc) rc = make_tmap_f16(&al, a.z_lo, (uint64_t)R, DAD_H, GBM);
    if (!rc) rc = make_tmap_f16(&bh, a.z_hi, (uint64_t)R, DAD_H, GBN);
    if (!rc) rc = make_tmap_f16(&bl, a.z_lo, (uint64_t)R, DAD_H, GBN);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_gram_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAM_SMEM);
        if (e != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "ecda gram: cannot opt in to %u B shared memory: %s", GRAM_SMEM, cudaGetErrorString(e));
        attr = true;
    }
    int grid = a.max_gram_items < sm_count() ? a.max_gram_items : sm_count();
    if (grid < 1) grid = 1;
    k_gram_tc<<<grid, GRAM_THREADS, GRAM_SMEM, st>>>(ah, al, bh, bl, a.gram_items, &a.stats->n_gram_items, a.G, a.ldg);
    DAD_CHECK_LAUNCH("k_gram_tc");
    return DAD_OK;
}

int launch_ecda_pz(const EcdaArgs& a, cudaStream_t st) {
    const int R = 2 * a.pack.B_local * a.pack.world;
    CUtensorMap ph, pl, zh, zl;
    int rc = make_tmap_f16(&ph, a.p_hi, (uint64_t)a.g_rows_alloc, (uint64_t)a.ldg, GBM);
    if (!rc) rc = make_tmap_f16(&pl, a.p_lo, (uint64_t)a.g_rows_alloc, (uint64_t)a.ldg, GBM);
    if (!rc) rc = make_tmap_f16(&zh, a.z_hi, (uint64_t)R, DAD_H, GBK);
    if (!rc) rc = make_tmap_f16(&zl, a.z_lo, (uint64_t)R, DAD_H, GBK);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_pz_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAM_SMEM);
        if (e != cudaSuccess) return dad_set_error(DAD_ERR_CUDA, "ecda pz: cannot opt in to %u B shared memory: %s", GRAM_SMEM, cudaGetErrorString(e));
        attr = true;
    }
    int grid = a.max_pz_items < sm_count() ? a.max_pz_items : sm_count();
    if (grid < 1) grid = 1;
    k_pz_tc<<<grid, GRAM_THREADS, GRAM_SMEM, st>>>(ph, pl, zh, zl, a.pz_items, &a.stats->n_pz_items, a.accz, a.g_rows_alloc);
    DAD_CHECK_LAUNCH("k_pz_tc");
    return DAD_OK;
}
