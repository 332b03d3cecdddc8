// tcgen05 / TMEM / TMA encoder kernels (bf16 hi/lo split, 3 products, fp32 accumulate) -- placeholder translation
// unit until the tensor-core path lands; DAD_GEMM_TC_BF16X3 reports "unsupported" rather than falling back.
#include "dad_common.cuh"
size_t enc_fwd_tc_ws_bytes(int64_t, int64_t) { return 256; }
int launch_enc_fwd_tc(const float*, const uint8_t*, const void*, const void*, const float*, const AugParams&, float*, uint32_t*, float*,
                      void*, size_t, int64_t, int64_t, cudaStream_t) {
    return dad_set_error(DAD_ERR_UNSUPPORTED, "encoder_fwd: DAD_GEMM_TC_BF16X3 not built");
}
int enc_bwd_tc_splits(int64_t) { return 1; }
size_t enc_bwd_tc_ws_bytes(int64_t, int64_t) { return 256; }
int launch_enc_bwd_tc(const float*, const uint8_t*, const AugParams&, const uint32_t*, const float*, const float*, float*, float*, int,
                      void*, size_t, int64_t, int64_t, cudaStream_t) {
    return dad_set_error(DAD_ERR_UNSUPPORTED, "encoder_bwd: DAD_GEMM_TC_BF16X3 not built");
}
