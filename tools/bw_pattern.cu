// Micro-benchmark (dev tool): achievable HBM read bandwidth of the encoder kernels' access patterns vs a sequential sweep.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_pattern bw_pattern.cu && ./bw_pattern
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

// pattern A: persistent CTA, tiles of 128 rows x 768 floats; per k-block (CHUNK floats of every row) 256 threads load float4s.
template <int CHUNK, int DEPTH>
__global__ void __launch_bounds__(256) k_tile(const float* __restrict__ x, float* __restrict__ out, int64_t n_rows, int num_tiles) {
    constexpr int F4 = CHUNK / 4;          // float4 per row per k-block
    constexpr int RPI = 256 / F4;          // rows covered per load instruction across the CTA
    constexpr int LPT = 128 / RPI;         // loads per thread per k-block
    constexpr int KB = 768 / CHUNK;
    const int c = threadIdx.x % F4, r0 = threadIdx.x / F4;
    float acc = 0.f;
    float4 buf[DEPTH][LPT];
    const int n_my = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_iter = n_my * KB;
    auto load = [&](float4 (&dst)[LPT], int it) {
        const int tile = blockIdx.x + (it / KB) * gridDim.x, kb = it % KB;
        const float* src = x + ((int64_t)tile * 128 + r0) * 768 + kb * CHUNK + c * 4;
#pragma unroll
        for (int i = 0; i < LPT; ++i) dst[i] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)i * RPI * 768));
    };
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) if (d < n_iter) load(buf[d], d);
    for (int it = 0; it < n_iter; it += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            if (it + d + DEPTH - 1 < n_iter) load(buf[(d + DEPTH - 1) % DEPTH], it + d + DEPTH - 1);
            if (it + d < n_iter) {
#pragma unroll
                for (int i = 0; i < LPT; ++i) acc += buf[d][i].x + buf[d][i].y + buf[d][i].z + buf[d][i].w;
            }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

// pattern B: grid-stride sequential float4 sweep
__global__ void __launch_bounds__(256) k_seq(const float4* __restrict__ x, float* __restrict__ out, int64_t n4) {
    float acc = 0.f;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256 * 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int64_t j = i + (int64_t)u * gridDim.x * 256; v[u] = j < n4 ? __ldg(x + j) : make_float4(0, 0, 0, 0); }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F> float time_ms(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}

int main() {
    const int64_t rows = 4096ll * 200; const int tiles = (int)(rows / 128);
    float *x, *out; cudaMalloc(&x, rows * 768 * 4); cudaMalloc(&out, 4); cudaMemset(x, 0, rows * 768 * 4);
    const double gb = rows * 768 * 4 / 1e9;
    float ms;
    ms = time_ms([&] { k_seq<<<148 * 8, 256>>>((const float4*)x, out, rows * 192); }); printf("sequential sweep            : %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = time_ms([&] { k_tile<64, 2><<<148, 256>>>(x, out, rows, tiles); });  printf("tile 128x(64 floats) depth 2: %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = time_ms([&] { k_tile<64, 3><<<148, 256>>>(x, out, rows, tiles); });  printf("tile 128x(64 floats) depth 3: %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = time_ms([&] { k_tile<64, 4><<<148, 256>>>(x, out, rows, tiles); });  printf("tile 128x(64 floats) depth 4: %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = time_ms([&] { k_tile<128, 2><<<148, 256>>>(x, out, rows, tiles); }); printf("tile 128x(128 floats) depth 2: %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = time_ms([&] { k_tile<256, 2><<<148, 256>>>(x, out, rows, tiles); }); printf("tile 128x(256 floats) depth 2: %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = time_ms([&] { k_tile<768, 2><<<148, 256>>>(x, out, rows, tiles); }); printf("tile 128x(768 floats) depth 2: %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = time_ms([&] { k_tile<64, 3><<<296, 256>>>(x, out, rows, tiles); });  printf("tile 128x64 depth 3, 2 CTA/SM: %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
