// Micro-benchmark: HBM throughput of the TNS tile pattern on B200.  A warp owns 32 rows of 4 KB;
// it walks the first 2944 bytes of every row in column blocks of C bytes (C = 128, 256, 512) --
// each block is read from all 32 rows, then written to the same place of a second buffer --
// against a plain sequential copy of the same number of bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stride_copy stride_copy.cu && ./stride_copy
#include <cstdio>
#include <cuda_runtime.h>

template <int C>  // bytes per row per block
__global__ void tile_copy(const float4 *__restrict__ x, float4 *__restrict__ y, int n_rows) {
    constexpr int Q = C / 16;                 // float4 per row per block
    constexpr int RPI = 32 / Q;               // rows per instruction
    constexpr int K = 32 / RPI;               // instructions per block
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const size_t row0 = (size_t)warp * 32;
    if (row0 >= (size_t)n_rows) return;
    const int cc = lane % Q, rr = lane / Q;
    constexpr int NB = 2944 / C;
    float4 cur[K], nxt[K];
#pragma unroll
    for (int k = 0; k < K; ++k) cur[k] = x[(row0 + rr + RPI * k) * 256 + cc];
    for (int b = 0; b < NB; ++b) {
        if (b + 1 < NB) {
#pragma unroll
            for (int k = 0; k < K; ++k) nxt[k] = x[(row0 + rr + RPI * k) * 256 + (b + 1) * Q + cc];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) y[(row0 + rr + RPI * k) * 256 + b * Q + cc] = cur[k];
#pragma unroll
        for (int k = 0; k < K; ++k) cur[k] = nxt[k];
    }
}

__global__ void seq_copy(const float4 *__restrict__ x, float4 *__restrict__ y, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = x[i];
}

template <class F>
float time_it(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    const int n_rows = 131072;
    float4 *x, *y;
    cudaMalloc(&x, (size_t)n_rows * 4096); cudaMalloc(&y, (size_t)n_rows * 4096);
    cudaMemset(x, 0, (size_t)n_rows * 4096); cudaMemset(y, 0, (size_t)n_rows * 4096);
    const double bytes = 2.0 * n_rows * 2944;
    for (int tpb : {128, 256}) {
        const int grid = n_rows / tpb;
        float ms = time_it([&] { tile_copy<128><<<grid, tpb>>>(x, y, n_rows); });
        printf("tile 128 B  tpb %d: %.3f ms  %.0f GB/s\n", tpb, ms, bytes / ms / 1e6);
        ms = time_it([&] { tile_copy<256><<<grid, tpb>>>(x, y, n_rows); });
        printf("tile 256 B  tpb %d: %.3f ms  %.0f GB/s\n", tpb, ms, bytes / ms / 1e6);
        ms = time_it([&] { tile_copy<512><<<grid, tpb>>>(x, y, n_rows); });
        printf("tile 512 B  tpb %d: %.3f ms  %.0f GB/s\n", tpb, ms, bytes / ms / 1e6);
    }
    const size_t n4 = (size_t)n_rows * 2944 / 16;
    float ms = time_it([&] { seq_copy<<<148 * 16, 256>>>(x, y, n4); });
    printf("sequential copy of the same bytes: %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
