// Micro-benchmark: issue rate of the packed FP32 FMA of sm_100 (PTX fma.rn.f32x2, SASS FFMA2)
// against scalar FFMA, per SM sub-partition.  The synthesis kernel runs the same arithmetic on
// two chains with shared twiddles, which is exactly the shape FFMA2 serves (one instruction for
// both chains); this measures whether it also costs one issue slot.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma2_rate fma2_rate.cu && ./fma2_rate
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 pack(float x, float y) {
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y));
    return d;
}
// negation written on the scalars folds into FFMA2's operand modifier (-R.F32x2.HI_LO)
__device__ __forceinline__ u64 neg2(u64 v) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
    return pack(-x, -y);
}

// MODE 0: scalar, 2*CHAINS dependent chains of 3-register FFMA (what two chains cost today)
// MODE 1: packed, CHAINS dependent chains of FFMA2 with three distinct 64-bit operands
// MODE 2: packed, one operand (the "twiddle") shared by consecutive FFMA2
// MODE 3: scalar butterfly pattern: 6 FFMA per butterfly, 2 chains (as aacfb_core.cuh bfly)
// MODE 4: packed butterfly pattern: 6 FFMA2 per butterfly pair
template <int CHAINS, int MODE>
__global__ void k(float *out, const float *in, int iters) {
    float a[2 * CHAINS], b[2 * CHAINS], c[2 * CHAINS];
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; ++i) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 7 + i]; c[i] = in[threadIdx.x + 3 * i + 1]; }
    u64 A[CHAINS], B[CHAINS], Cc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { A[i] = pack(a[2 * i], a[2 * i + 1]); B[i] = pack(b[2 * i], b[2 * i + 1]); Cc[i] = pack(c[2 * i], c[2 * i + 1]); }
    const float y = in[threadIdx.x + 5];
    const float y2 = in[threadIdx.x + 9];
    const u64 Y = pack(y, y), Y2 = pack(y2, y2), TWO = pack(2.0f, 2.0f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 2 * CHAINS; ++i) a[i] = __fmaf_rn(b[i], c[i], a[i]);
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < CHAINS; ++i) A[i] = fma2(B[i], Cc[i], A[i]);
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < CHAINS; ++i) A[i] = fma2(Y, Cc[i], A[i]);
            } else if (MODE == 3) {   // (a[0],a[1]) +- (a[2],a[3]) * (y, b[0]) for each chain pair of 4 floats
#pragma unroll
                for (int i = 0; i + 3 < 2 * CHAINS; i += 4) {
                    const float lr = __fmaf_rn(a[i + 2], y, __fmaf_rn(-a[i + 3], y2, a[i]));
                    const float li = __fmaf_rn(a[i + 2], y2, __fmaf_rn(a[i + 3], y, a[i + 1]));
                    a[i + 2] = __fmaf_rn(2.0f, a[i], -lr);
                    a[i + 3] = __fmaf_rn(2.0f, a[i + 1], -li);
                    a[i] = lr; a[i + 1] = li;
                }
            } else {
#pragma unroll
                for (int i = 0; i + 3 < CHAINS; i += 4) {   // same butterfly on packed (chain0, chain1) values
                    const u64 lr = fma2(A[i + 2], Y, fma2(neg2(A[i + 3]), Y2, A[i]));
                    const u64 li = fma2(A[i + 2], Y2, fma2(A[i + 3], Y, A[i + 1]));
                    A[i + 2] = fma2(TWO, A[i], neg2(lr));
                    A[i + 3] = fma2(TWO, A[i + 1], neg2(li));
                    A[i] = lr; A[i + 1] = li;
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; ++i) s += a[i];
    u64 S = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) S ^= A[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(S & 0xffff);
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
    const double cycles = ms * 1e-3 * 1.92e9;  // assume ~1.92 GHz under load
    // scalar-equivalent FMAs per thread per unrolled round
    double fmas = (MODE == 0 || MODE == 1 || MODE == 2) ? 2.0 * CHAINS : (MODE == 3 ? 6.0 * (2 * CHAINS / 4) : 12.0 * (CHAINS / 4));
    const double per_smsp = (double)iters * 16 * fmas * warps_per_sm / 4.0;
    printf("mode %d chains %2d warps/SM %2d: %.3f ms -> %.2f cycles per scalar-equivalent warp-FMA per SMSP\n", MODE, CHAINS,
           warps_per_sm, ms, cycles / per_smsp);
}

int main() {
    float *in, *out;
    cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20);
    cudaMalloc(&out, 1 << 20);
    run<8, 0>(4, out, in); run<8, 0>(12, out, in); run<8, 0>(16, out, in);
    run<8, 1>(4, out, in); run<8, 1>(12, out, in); run<8, 1>(16, out, in);
    run<8, 2>(4, out, in); run<8, 2>(12, out, in); run<8, 2>(16, out, in);
    run<8, 3>(4, out, in); run<8, 3>(12, out, in); run<8, 3>(16, out, in);
    run<8, 4>(4, out, in); run<8, 4>(12, out, in); run<8, 4>(16, out, in);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
