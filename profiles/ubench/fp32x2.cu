// Micro-benchmark: scalar FFMA vs packed fma.rn.f32x2 (FFMA2) and MUFU.RCP throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, int iters) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float m = 1.0000001f, c = 1e-9f;
    for (int i = 0; i < iters; ++i) {
        a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
        a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__global__ void k_ffma2(float* out, int iters) {
    float2 f = make_float2(threadIdx.x, threadIdx.x + 0.5f);
    unsigned long long a0 = *reinterpret_cast<unsigned long long*>(&f), a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4,
                       a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float2 mm = make_float2(1.0000001f, 1.0000001f), cc = make_float2(1e-9f, 1e-9f);
    const unsigned long long m = *reinterpret_cast<unsigned long long*>(&mm), c = *reinterpret_cast<unsigned long long*>(&cc);
    for (int i = 0; i < iters; ++i) {
        a0 = ffma2(a0, m, c); a1 = ffma2(a1, m, c); a2 = ffma2(a2, m, c); a3 = ffma2(a3, m, c);
        a4 = ffma2(a4, m, c); a5 = ffma2(a5, m, c); a6 = ffma2(a6, m, c); a7 = ffma2(a7, m, c);
    }
    unsigned long long s = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
__global__ void k_rcp(float* out, int iters) {
    float a0 = 1.1f + threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    for (int i = 0; i < iters; ++i) {
        a0 = __frcp_rn(a0) + 1.f; a1 = __fdividef(1.f, a1); a2 = __fdividef(1.f, a2); a3 = __fdividef(1.f, a3);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
int main() {
    float* out; cudaMalloc(&out, 1 << 24);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    const int it = 1 << 15, blocks = 148 * 8, thr = 256;
    for (int r = 0; r < 2; ++r) { cudaEventRecord(e0); k_ffma<<<blocks, thr>>>(out, it); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    printf("FFMA  : %.1f TFLOP/s\n", 2.0 * 8 * it * blocks * thr / (ms * 1e-3) / 1e12);
    for (int r = 0; r < 2; ++r) { cudaEventRecord(e0); k_ffma2<<<blocks, thr>>>(out, it); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    printf("FFMA2 : %.1f TFLOP/s\n", 4.0 * 8 * it * blocks * thr / (ms * 1e-3) / 1e12);
    for (int r = 0; r < 2; ++r) { cudaEventRecord(e0); k_rcp<<<blocks, thr>>>(out, it / 4); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    printf("MUFU.RCP (3 approx + 1 rn per iter): %.2f G rcp/s\n", 4.0 * (it / 4) * blocks * thr / (ms * 1e-3) / 1e9);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
