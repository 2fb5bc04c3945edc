// Micro-benchmark: FFMA issue rate per SM sub-partition on sm_100a, as a function of the warps
// per scheduler, the number of independent chains per thread and the operand pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, int MODE>
__global__ void k(float *out, const float *in, int iters) {
    float a[CHAINS], b[CHAINS], c[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 7 + i]; c[i] = in[threadIdx.x + 3 * i + 1]; }
    const float y = in[threadIdx.x + 5];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                if (MODE == 0) a[i] = __fmaf_rn(b[i], c[i], a[i]);        // 3 distinct registers, dependent on a[i]
                else if (MODE == 1) a[i] = __fmaf_rn(y, c[i], a[i]);      // one operand shared by consecutive FFMAs
                else a[i] = __fmaf_rn(a[i], 1.0009765625f, c[i]);         // immediate form
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS, int MODE>
void run(int warps_per_sm, float *out, float *in) {
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CHAINS, MODE><<<148, warps_per_sm * 32>>>(out, in, 10);
    cudaEventRecord(e0);
    k<CHAINS, MODE><<<148, warps_per_sm * 32>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int mhz; cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * 1.92e9;  // assume ~1.92 GHz under load
    const double ffma_per_smsp = (double)iters * 16 * CHAINS * warps_per_sm / 4.0;
    printf("mode %d chains %2d warps/SM %2d: %.3f ms  -> %.2f cycles per warp-FFMA per SMSP\n", MODE, CHAINS, warps_per_sm, ms,
           cycles / ffma_per_smsp);
}

int main() {
    float *in, *out;
    cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20);
    cudaMalloc(&out, 1 << 20);
    run<1, 0>(4, out, in); run<1, 0>(8, out, in); run<1, 0>(16, out, in); run<1, 0>(32, out, in);
    run<4, 0>(4, out, in); run<4, 0>(8, out, in); run<4, 0>(16, out, in);
    run<8, 0>(4, out, in); run<8, 0>(8, out, in); run<8, 0>(16, out, in);
    run<8, 1>(4, out, in); run<8, 1>(8, out, in); run<8, 1>(16, out, in);
    run<8, 2>(4, out, in); run<8, 2>(8, out, in); run<8, 2>(16, out, in);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
