// Micro-benchmark: cycles per output of the TNS all-pole chain (tns.js:156-162: 12 dependent rounded
// multiply-subtracts per coefficient, newest output first) for one lane-per-row warp,
//   MODE 0: scalar FFMA, one row per lane
//   MODE 1: packed fma.rn.f32x2 (FFMA2), two rows per lane (the two channels of a frame)
//   MODE 2: scalar FFMA, two independent rows per lane (instruction-level parallelism only)
// with 1..4 such warps per SM sub-partition.  Decides how many rows per lane the TNS worker of the
// fused kernel carries.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tns_chain tns_chain.cu && ./tns_chain
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
__device__ __forceinline__ void unpack(u64 v, float &x, float &y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }

constexpr int ORD = 12;
#ifndef NEG
#define NEG 0
#endif
__device__ __forceinline__ u64 neg2(u64 v) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); return pack(-x, -y); }

template <int MODE>
__global__ void k(float *out, const float *in, int n_out, long long *cycles) {
    __shared__ float4 xs[2][32][9];   // one 128-byte tile row per lane and channel, pitch 144 B
    const int lane = threadIdx.x & 31;
    for (int q = 0; q < 8; ++q) {
        xs[0][lane][q] = make_float4(in[lane + q], in[lane + q + 1], in[lane + q + 2], in[lane + q + 3]);
        xs[1][lane][q] = make_float4(in[lane + q + 4], in[lane + q + 5], in[lane + q + 6], in[lane + q + 7]);
    }
    float c[2][ORD], h[2][ORD];
    u64 C[ORD], H[ORD];
#pragma unroll
    for (int i = 0; i < ORD; ++i) {
        c[0][i] = in[64 + i + lane] * 1e-3f; c[1][i] = in[128 + i + 2 * lane] * 1e-3f; h[0][i] = h[1][i] = 0.f;
        C[i] = pack(c[0][i], c[1][i]); H[i] = pack(0.f, 0.f);
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int m = 0; m < n_out; m += 32) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 a = xs[0][lane][q], b = xs[1][lane][q];
            const float va[4] = {a.x, a.y, a.z, a.w}, vb[4] = {b.x, b.y, b.z, b.w};
            float ya[4], yb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (MODE == 1) {
                    u64 acc = pack(va[j], vb[j]);
#pragma unroll
                    for (int i = 0; i < ORD; ++i) acc = NEG ? fma2(H[i], neg2(C[i]), acc) : fma2(H[i], C[i], acc);
#pragma unroll
                    for (int i = ORD - 1; i > 0; --i) H[i] = H[i - 1];
                    H[0] = acc;
                    unpack(acc, ya[j], yb[j]);
                } else {
                    float acc = va[j], acc1 = vb[j];
#pragma unroll
                    for (int i = 0; i < ORD; ++i) {
                        acc = __fmaf_rn(h[0][i], c[0][i], acc);
                        if (MODE == 2) acc1 = __fmaf_rn(h[1][i], c[1][i], acc1);
                    }
#pragma unroll
                    for (int i = ORD - 1; i > 0; --i) { h[0][i] = h[0][i - 1]; if (MODE == 2) h[1][i] = h[1][i - 1]; }
                    h[0][0] = acc; if (MODE == 2) h[1][0] = acc1;
                    ya[j] = acc; yb[j] = acc1;
                }
            }
            xs[0][lane][q] = make_float4(ya[0], ya[1], ya[2], ya[3]);
            if (MODE != 0) xs[1][lane][q] = make_float4(yb[0], yb[1], yb[2], yb[3]);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = xs[0][lane][3].x + xs[1][lane][5].y;
}

template <int MODE>
void run(const char *name, int warps, float *out, float *in, long long *cyc) {
    const int n_out = 32 * 23 * 8;
    k<MODE><<<148, 32 * warps>>>(out, in, n_out, cyc);
    k<MODE><<<148, 32 * warps>>>(out, in, n_out, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    printf("%-28s warps/CTA %2d: %.2f cycles per output step (%.2f per FMA step)\n", name, warps, avg / n_out, avg / n_out / ORD);
}

int main() {
    float *out, *in; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096); cudaMalloc(&cyc, 148 * 8);
    float hin[1024];
    for (int i = 0; i < 1024; ++i) hin[i] = 0.001f * (i % 37) - 0.01f;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    for (int w : {1, 4, 8, 16}) {
        run<0>("scalar, 1 row/lane", w, out, in, cyc);
        run<1>("FFMA2, 2 rows/lane", w, out, in, cyc);
        run<2>("scalar, 2 rows/lane (ILP)", w, out, in, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
