// exp(x) for the Arrhenius factors of the kinetic model (4 per channel and right-hand side: a third to a half of
// the FP64 instructions of the reactor march with the library exp()).
//   x = (64 e + j) ln2/64 + r, |r| <= ln2/128;   exp(x) = 2^e * 2^(j/64) * (1 + expm1(r))
// 2^(j/64) from a 64-entry table (shared memory on the device), expm1(r) by a degree-5 polynomial (truncation
// 3.5e-17), 2^e by an integer add to the exponent field: 10 FP64-pipe operations against ~22.  Error < 1.1 ulp
// (table rounding + last fma; the library's is 1 ulp).  Branch-free, so that the four factors of a channel overlap,
// and the range test is integer work on the bits of x: |x| >= 708 gives 0 or +inf (the library returns subnormals
// between -745 and -708 and finite values up to 709.78; a rate constant built from either is zero or overflows all
// the same), NaN stays NaN.
// Coefficients live in constant memory on the device: as literals each one costs two UMOVs per use.
// The header also compiles with plain g++ (tests/host_exp.cpp) so that the arithmetic can be checked without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define EXPT_HD __device__ __forceinline__
#define EXPT_CONST static __constant__ double
#else
#define EXPT_HD inline
#define EXPT_CONST static const double
#endif

namespace expt {

EXPT_CONST C[8] = {92.33248261689366,        // 64/ln2
                   6755399441055744.0,        // 1.5 * 2^52: adding it leaves rint(x*64/ln2) in the low mantissa bits
                   -0x1.62e42fefa39efp-7,     // -(ln2/64), high part
                   -0x1.abc9e3b39803fp-62,    // -(ln2/64), low part
                   1.0 / 120, 1.0 / 24, 1.0 / 6, 0.5};
constexpr int TAB_N = 64;
EXPT_CONST TAB[TAB_N] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};

#if defined(__CUDACC__)
__device__ __forceinline__ void load_table(double* tab) {   // once per block, then __syncthreads()
    for (int j = threadIdx.x; j < TAB_N; j += blockDim.x) tab[j] = TAB[j];
}
__device__ __forceinline__ int lo_bits(double t) { return __double2loint(t); }
__device__ __forceinline__ int hi_bits(double t) { return __double2hiint(t); }
__device__ __forceinline__ double from_bits(int hi, int lo) { return __hiloint2double(hi, lo); }
#else
inline int lo_bits(double t) {
    uint64_t u;
    memcpy(&u, &t, 8);
    return (int)(uint32_t)u;
}
inline int hi_bits(double t) {
    uint64_t u;
    memcpy(&u, &t, 8);
    return (int)(uint32_t)(u >> 32);
}
inline double from_bits(int hi, int lo) {
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double y;
    memcpy(&y, &u, 8);
    return y;
}
#endif

// exp(x) for |x| < 708 (the caller has checked): no range handling at all
EXPT_HD double exp_core(double x, const double* tab) {
    const double t = fma(x, C[0], C[1]);
    const int k = lo_bits(t);
    const double kf = t - C[1];
    double r = fma(kf, C[2], x);
    r = fma(kf, C[3], r);
    double q = fma(r, C[4], C[5]);
    q = fma(q, r, C[6]);
    q = fma(q, r, C[7]);
    const double p = fma(q, r * r, r);
    const double tj = tab[k & 63];
    const double y = fma(tj, p, tj);
    return from_bits(hi_bits(y) + ((k >> 6) << 20), lo_bits(y));
}

// exp(x) for any x: out-of-range and non-finite arguments replaced by integer selects on the bits of x (written out
// rather than wrapped around exp_core: through a repacked double the compiler turns the selects into branches)
EXPT_HD double exp_fast(double x, const double* tab) {
    const double t = fma(x, C[0], C[1]);
    const int k = lo_bits(t);
    const double kf = t - C[1];
    double r = fma(kf, C[2], x);
    r = fma(kf, C[3], r);
    double q = fma(r, C[4], C[5]);
    q = fma(q, r, C[6]);
    q = fma(q, r, C[7]);
    const double p = fma(q, r * r, r);
    const double tj = tab[k & 63];
    const double y = fma(tj, p, tj);
    const int hx = hi_bits(x), lx = lo_bits(x);
    const unsigned ax = (unsigned)hx & 0x7fffffffu;
    int hi = hi_bits(y) + ((k >> 6) << 20), lo = lo_bits(y);
    const bool big = ax >= 0x40862000u;                                   // |x| >= 708, inf, NaN
    const bool nan = ax > 0x7ff00000u || (ax == 0x7ff00000u && lx != 0);
    hi = big ? (nan ? hx : (hx < 0 ? 0 : 0x7ff00000)) : hi;
    lo = big ? (nan ? lx : 0) : lo;
    return from_bits(hi, lo);
}

}  // namespace expt
