// Micro-benchmark: dependent-issue latency and throughput of the FP64 instructions the MM solver is
// built from (DFMA, DADD, MUFU.RCP64H + correction), and the cost of one attempted RK45 step of a
// single lane.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../python-based-sequential-monte-carlo-method-with-likelihood-tempering_b200/csrc/mm_solver.cuh"

__global__ void lat_dfma(double* out, long long* cyc, int iters, double m, double c) {
    double a = threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) a = fma(a, m, c);
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_dadd(double* out, long long* cyc, int iters, double c) {
    double a = threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) a = a + c;
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_rcpseed(double* out, long long* cyc, int iters) {
    double a = 1.5 + threadIdx.x * 1e-3;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) a = mmsolve::rcp_seed(a);
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_mmrate(double* out, long long* cyc, int iters, double c, double Km) {
    double a = 1.5 + threadIdx.x * 1e-3;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) a = mmsolve::mm_rate(c, Km, a) + 2.0;
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_rootm5(double* out, long long* cyc, int iters) {
    double a = 0.5 + threadIdx.x * 1e-3;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) a = mmsolve::rootm5(a);
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput: 8 independent chains per thread, many warps
__global__ void thr_dfma(double* out, int iters, double m, double c) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void thr_rcp(double* out, int iters) {
    double a0 = 1.1 + threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    for (int i = 0; i < iters; ++i) {
        a0 = mmsolve::rcp_seed(a0); a1 = mmsolve::rcp_seed(a1); a2 = mmsolve::rcp_seed(a2); a3 = mmsolve::rcp_seed(a3);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
// one stiff solve, `lanes` active lanes of one warp all integrating the same problem
__global__ void solve_steps(double* out, long long* cyc, unsigned* natt, double Vmax, double Km, double S0, int lanes,
                            double cut = INFINITY) {
    __shared__ mmsolve::ObsPair obs[40];
    for (int i = threadIdx.x; i < 40; i += blockDim.x) { obs[i].P = 0.05; obs[i].t_next = (i < 39) ? 10.0 * (i + 1) / 39.0 : INFINITY; }
    __syncthreads();
    if ((int)threadIdx.x >= lanes) return;
    mmsolve::Solve s;
    s.nVmax = -Vmax; s.Km = Km; s.S0 = S0; s.cut_lim = cut;
    unsigned a = 0, r = 0;
    long long t0 = clock64();
    int st = mmsolve::setup(s, 0.0, 10.0) ? mmsolve::RUNNING : mmsolve::FAILED;
    while (st == mmsolve::RUNNING) st = mmsolve::attempt<false>(s, obs, nullptr, a, r);
    long long t1 = clock64();
    out[threadIdx.x] = s.ssr;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; natt[0] = a + r; }
}

// the same solve in every lane of `grid` one-warp blocks: inter-warp contention without divergence
__global__ void solve_many(double* out, long long* cyc, unsigned* natt, double Vmax, double Km, double S0, int active_lanes) {
    __shared__ mmsolve::ObsPair obs[40];
    for (int i = threadIdx.x; i < 40; i += blockDim.x) { obs[i].P = 0.05; obs[i].t_next = (i < 39) ? 10.0 * (i + 1) / 39.0 : INFINITY; }
    __syncthreads();
    if ((int)threadIdx.x >= active_lanes) return;
    mmsolve::Solve s;
    // lanes other than 0 integrate a neighbouring problem (slightly different Km): same length class, own path
    s.nVmax = -Vmax; s.Km = Km * (1.0 + 0.01 * threadIdx.x); s.S0 = S0; s.cut_lim = INFINITY;
    unsigned a = 0, r = 0;
    long long t0 = clock64();
    int st = mmsolve::setup(s, 0.0, 10.0) ? mmsolve::RUNNING : mmsolve::FAILED;
    while (st == mmsolve::RUNNING) st = mmsolve::attempt<false>(s, obs, nullptr, a, r);
    long long t1 = clock64();
    out[blockIdx.x * 32 + threadIdx.x] = s.ssr;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; natt[0] = a + r; }
}

// the tail kernel's spelling (mmsolve::solve_lat): lanes other than 0 integrate neighbouring problems
template <bool HOIST>
__global__ void solve_lat_many(double* out, long long* cyc, unsigned* natt, double Vmax, double Km, double S0, int active_lanes,
                               double spread) {
    __shared__ mmsolve::ObsPair obs[40];
    __shared__ double s_coef[mmsolve::LAT_NCOEF];
    if (threadIdx.x == 0) mmsolve::lat_coef_fill(s_coef);
    for (int i = threadIdx.x; i < 40; i += blockDim.x) { obs[i].P = 0.05; obs[i].t_next = (i < 39) ? 10.0 * (i + 1) / 39.0 : INFINITY; }
    __syncthreads();
    if ((int)threadIdx.x >= active_lanes) return;
    mmsolve::Solve s;
    s.nVmax = -Vmax; s.Km = Km * (1.0 + spread * threadIdx.x); s.S0 = S0; s.cut_lim = INFINITY;
    unsigned a = 0, r = 0;
    long long t0 = clock64();
    int st = mmsolve::setup(s, 0.0, 10.0) ? mmsolve::RUNNING : mmsolve::FAILED;
    if (st == mmsolve::RUNNING) st = mmsolve::solve_lat<HOIST>(s, obs, a, r, s_coef);
    long long t1 = clock64();
    out[blockIdx.x * 32 + threadIdx.x] = s.ssr;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; natt[0] = a + r; cyc[1] = (long long)__double_as_longlong(s.ssr); }
}

int main() {
    double* out; long long* cyc; unsigned* natt;
    cudaMalloc(&out, 1 << 24); cudaMallocManaged(&cyc, 64); cudaMallocManaged(&natt, 64);
    const int it = 1 << 16;
    lat_dfma<<<1, 32>>>(out, cyc, it, 1.0000001, 1e-9); cudaDeviceSynchronize();
    printf("dependent DFMA            : %.2f cycles\n", (double)cyc[0] / it);
    lat_dadd<<<1, 32>>>(out, cyc, it, 1e-9); cudaDeviceSynchronize();
    printf("dependent DADD            : %.2f cycles\n", (double)cyc[0] / it);
    lat_rcpseed<<<1, 32>>>(out, cyc, it); cudaDeviceSynchronize();
    printf("dependent MUFU.RCP64H     : %.2f cycles\n", (double)cyc[0] / it);
    lat_mmrate<<<1, 32>>>(out, cyc, it, -1.3, 0.7); cudaDeviceSynchronize();
    printf("dependent mm_rate()+DADD  : %.2f cycles\n", (double)cyc[0] / it);
    lat_rootm5<<<1, 32>>>(out, cyc, it); cudaDeviceSynchronize();
    printf("dependent rootm5()        : %.2f cycles\n", (double)cyc[0] / it);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); thr_dfma<<<148 * 8, 256>>>(out, 1 << 14, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    printf("DFMA throughput           : %.2f TFLOP/s (%.1f lanes/clk/SM at 1.965 GHz)\n", 2.0 * 8 * (1 << 14) * 148 * 8 * 256 / (ms * 1e-3) / 1e12,
           8.0 * (1 << 14) * 148 * 8 * 256 / (ms * 1e-3) / 148 / 1.965e9);
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); thr_rcp<<<148 * 8, 256>>>(out, 1 << 14); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    printf("MUFU.RCP64H throughput    : %.1f lanes/clk/SM at 1.965 GHz\n", 4.0 * (1 << 14) * 148 * 8 * 256 / (ms * 1e-3) / 148 / 1.965e9);
    for (int lanes = 1; lanes <= 32; lanes *= 32) {
        solve_steps<<<1, 32>>>(out, cyc, natt, 7.30644405, 4.35543917e-04, 0.1, lanes); cudaDeviceSynchronize();
        printf("stiff solve, %2d lane(s)    : %u attempts, %.1f cycles per attempt (%.3f us at 1.965 GHz)\n", lanes, natt[0],
               (double)cyc[0] / natt[0], (double)cyc[0] / natt[0] / 1965.0);
    }
    solve_steps<<<1, 32>>>(out, cyc, natt, 9.88869508e+00, 4.28267643e-04, 0.1, 1); cudaDeviceSynchronize();
    printf("worst solve of the 2^20 bench prior, 1 lane: %u attempts, %.1f cycles per attempt, %.2f ms\n", natt[0],
           (double)cyc[0] / natt[0], (double)cyc[0] / 1.965e6);
    solve_steps<<<1, 32>>>(out, cyc, natt, 9.88869508e+00, 4.28267643e-04, 0.1, 1, 1e300); cudaDeviceSynchronize();
    printf("worst solve, 1 lane, run-time residual limit: %.1f cycles per attempt\n", (double)cyc[0] / natt[0]);
    for (int wps = 1; wps <= 12; wps *= 2) {
        if (wps == 8) wps = 12;
        solve_many<<<148 * wps, 32>>>(out, cyc, natt, 9.88869508e+00, 4.28267643e-04, 0.1, 1); cudaDeviceSynchronize();
        printf("same solve, lane 0 of %2d warps per SM : %.1f cycles per attempt\n", wps, (double)cyc[0] / natt[0]);
    }
    for (int lanes = 2; lanes <= 32; lanes *= 4) {
        solve_many<<<148 * 4, 32>>>(out, cyc, natt, 9.88869508e+00, 4.28267643e-04, 0.1, lanes); cudaDeviceSynchronize();
        printf("4 warps per SM, %2d lanes with neighbouring stiff problems: %.1f cycles per attempt of lane 0 (%u attempts)\n", lanes, (double)cyc[0] / natt[0], natt[0]);
    }
    // ---- round 2: the latency spelling of the tail kernel (solve_lat), registers-resident coefficients or not
    solve_steps<<<1, 32>>>(out, cyc, natt, 9.88869508e+00, 4.28267643e-04, 0.1, 1); cudaDeviceSynchronize();
    printf("attempt() loop, worst solve, 1 lane            : %u attempts, %.1f cycles per attempt, ssr bits %016llx\n", natt[0], (double)cyc[0] / natt[0],
           (unsigned long long)0);
    for (int lanes = 1; lanes <= 32; lanes *= 8) {
        if (lanes == 64) break;
        solve_lat_many<false><<<1, 32>>>(out, cyc, natt, 9.88869508e+00, 4.28267643e-04, 0.1, lanes, 0.01); cudaDeviceSynchronize();
        printf("solve_lat<false>, %2d lane(s), 1 warp             : %u attempts, %.1f cycles per attempt\n", lanes, natt[0], (double)cyc[0] / natt[0]);
        solve_lat_many<true><<<1, 32>>>(out, cyc, natt, 9.88869508e+00, 4.28267643e-04, 0.1, lanes, 0.01); cudaDeviceSynchronize();
        printf("solve_lat<true>,  %2d lane(s), 1 warp             : %u attempts, %.1f cycles per attempt\n", lanes, natt[0], (double)cyc[0] / natt[0]);
    }
    // lanes of very different stiffness in one warp (Km spread 10x per lane: the other lanes finish early or wait at the loop's exit)
    solve_lat_many<false><<<1, 32>>>(out, cyc, natt, 9.88869508e+00, 4.28267643e-04, 0.1, 8, 3.0); cudaDeviceSynchronize();
    printf("solve_lat<false>, 8 lanes of mixed stiffness      : %u attempts of lane 0, %.1f cycles per attempt\n", natt[0], (double)cyc[0] / natt[0]);
    solve_lat_many<true><<<1, 32>>>(out, cyc, natt, 9.88869508e+00, 4.28267643e-04, 0.1, 8, 3.0); cudaDeviceSynchronize();
    printf("solve_lat<true>,  8 lanes of mixed stiffness      : %u attempts of lane 0, %.1f cycles per attempt\n", natt[0], (double)cyc[0] / natt[0]);
    for (int wps = 4; wps <= 16; wps *= 2) {
        solve_lat_many<false><<<148 * wps, 32>>>(out, cyc, natt, 9.88869508e+00, 4.28267643e-04, 0.1, 1, 0.01); cudaDeviceSynchronize();
        printf("solve_lat<false>, lane 0 of %2d warps per SM       : %.1f cycles per attempt\n", wps, (double)cyc[0] / natt[0]);
        solve_lat_many<true><<<148 * wps, 32>>>(out, cyc, natt, 9.88869508e+00, 4.28267643e-04, 0.1, 1, 0.01); cudaDeviceSynchronize();
        printf("solve_lat<true>,  lane 0 of %2d warps per SM       : %.1f cycles per attempt\n", wps, (double)cyc[0] / natt[0]);
    }
    solve_steps<<<1, 32>>>(out, cyc, natt, 1.2, 0.5, 2.0, 1); cudaDeviceSynchronize();
    printf("posterior solve, 1 lane   : %u attempts, %.1f cycles per attempt\n", natt[0], (double)cyc[0] / natt[0]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
