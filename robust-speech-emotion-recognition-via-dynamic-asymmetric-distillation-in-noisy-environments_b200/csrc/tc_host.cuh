// Host-side helpers of the tcgen05 kernels: TMA tensor-map construction (driver entry point looked up at run time) and SM count.
#pragma once
#include "dad_common.cuh"
#include <cuda.h>
#include <mutex>

namespace {

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

// fp16 [rows][cols] row-major, box = [box_rows][64 cols] (128 bytes inner), SWIZZLE_128B
int make_tmap_f16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return dad_set_error(DAD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
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

// A library-internal helper stream per device for work that is INDEPENDENT of a launcher's main chain (e.g. the bias gradient beside
// the dW GEMM): fork() makes the helper stream wait for everything enqueued on `st` so far, join() makes `st` wait for the helper
// stream.  Plain event record / wait: legal under stream capture (becomes a parallel branch of the captured graph).
struct SideLane {
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int dev = -1;
};
inline SideLane* side_lane(int which, cudaStream_t st) {
    static SideLane lanes[16][2];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    SideLane& l = lanes[dev][which & 1];
    if (!l.side) {
        // never create streams / events while `st` is being captured (resource creation is not a capturable operation): a launcher
        // whose very first call on this device happens under capture simply stays on its main stream for that graph
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return nullptr;
        if (cudaStreamCreateWithFlags(&l.side, cudaStreamNonBlocking) != cudaSuccess) { l.side = nullptr; return nullptr; }
        if (cudaEventCreateWithFlags(&l.ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&l.ev_join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        l.dev = dev;
    }
    return &l;
}
inline bool side_fork(SideLane* l, cudaStream_t st) {
    return l && cudaEventRecord(l->ev_fork, st) == cudaSuccess && cudaStreamWaitEvent(l->side, l->ev_fork, 0) == cudaSuccess;
}
inline bool side_join(SideLane* l, cudaStream_t st) {
    return l && cudaEventRecord(l->ev_join, l->side) == cudaSuccess && cudaStreamWaitEvent(st, l->ev_join, 0) == cudaSuccess;
}

}  // namespace
