// One Michaelis-Menten progress-curve solve, as the reference defines it.
//
// The reference likelihood (SMC_example/Micmem_likelihood.py:14-33, 59-71) integrates
//     dS/dt = -Vmax*S/(Km+S),  S(0) = S0
// with scipy.integrate.solve_ivp(method="RK45") at the default tolerances (rtol 1e-3, atol 1e-6) and
// reads S at the 40 observation times through the quartic dense output.  At rtol = 1e-3 the result
// depends on the controller (up to 2.9e-3 relative away from a converged solve), so the
// likelihood is *defined* by scipy's sequence of accepted and rejected steps.  This header takes the
// same steps:
//     tableau C/A/B/E/P                scipy/integrate/_ivp/rk.py:538-567
//     select_initial_step              common.py:110-134
//     step loop, accept/reject, factor rk.py:111-176  (SAFETY .9, MIN_FACTOR .2, MAX_FACTOR 10, exponent -1/5)
//     stages, FSAL, error estimate     rk.py:61-109
//     dense output at t_eval           rk.py:178-180,723-737 and ivp.py:712-728
//
// Arithmetic.  Same quantities, FP64 throughout, but spelt for the FP64 pipe of sm_100a instead of in
// NumPy's operation order:
//   * the stage derivatives are carried pre-multiplied by the step, K_j = h*k_j, so a stage argument
//     is a pure FMA chain y + sum a_sj K_j, the error estimate is sum E_j K_j and the dense-output
//     coefficients need no further scaling (92 FP64 operations per attempted step against ~130
//     for the unscaled form);
//   * S/(Km+S) by reciprocal: MUFU.RCP64H seed (rcp.approx.ftz.f64, 20 bits) and one cubically
//     convergent correction r0*(1+e+e^2), e = 1-den*r0, folded into the product (error < 1 ulp);
//   * err^(-1/5) from an FP32 lg2/ex2 seed and one FP64 correction instead of pow().
// Every operation is within an ulp or two of the one scipy performs; over the reference's own 34
// sweeps (prior cloud to posterior) the log-likelihoods agree with scipy's to better than 1e-9
// relative (tests/test_gpu_kernels.py, tests/test_host_twin.py), far inside the 1e-5 bar.
//
// The header also compiles with plain g++ (tests/host_twin.cpp; the MUFU seeds are replaced by truncated
// host values, which the corrections make irrelevant) so that tests can compare this very arithmetic
// with scipy without a GPU.  The product never runs the host build.
#pragma once
#include <math.h>
#include <stdint.h>

// Device build: functions are __device__ and every coefficient lives in constant memory, so an FP64
// instruction takes it straight from the constant bank (as a literal each one costs two UMOVs per use:
// ncu showed those at 21% of all issued instructions).  Host build (g++, tests only): plain inline / constexpr.
#if defined(__CUDACC__)
#define MM_HD __device__ __forceinline__
#define MM_COEF static __constant__ double
#else
#define MM_HD inline
#define MM_COEF static constexpr double
#endif

namespace mmsolve {

MM_COEF RTOL = 1e-3, ATOL = 1e-6, SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0;
// 0.9*err^(-1/5) reaches MAX_FACTOR for err <= 0.09^5 and MIN_FACTOR for err >= 4.5^5
MM_COEF ERR_LO = 5.9049e-6, ERR_HI = 1845.28125, ERR_ONE = 0.59049;   // 0.9*err^(-1/5) = 10, 0.2, 1

// Dormand-Prince coefficients spelt as scipy spells them; the quotients are evaluated in FP64.
MM_COEF A21 = 1.0 / 5;
MM_COEF A31 = 3.0 / 40, A32 = 9.0 / 40;
MM_COEF A41 = 44.0 / 45, A42 = -56.0 / 15, A43 = 32.0 / 9;
MM_COEF A51 = 19372.0 / 6561, A52 = -25360.0 / 2187, A53 = 64448.0 / 6561, A54 = -212.0 / 729;
MM_COEF A61 = 9017.0 / 3168, A62 = -355.0 / 33, A63 = 46732.0 / 5247, A64 = 49.0 / 176,
                 A65 = -5103.0 / 18656;
MM_COEF B1 = 35.0 / 384, B3 = 500.0 / 1113, B4 = 125.0 / 192, B5 = -2187.0 / 6784, B6 = 11.0 / 84;
MM_COEF E1 = -71.0 / 57600, E3 = 71.0 / 16695, E4 = -71.0 / 1920, E5 = 17253.0 / 339200,
                 E6 = -22.0 / 525, E7 = 1.0 / 40;
// dense-output matrix P (7 x 4); row 2 is zero, P11 = 1.
MM_COEF P12 = -8048581381.0 / 2820520608, P13 = 8663915743.0 / 2820520608,
                 P14 = -12715105075.0 / 11282082432;
MM_COEF P32 = 131558114200.0 / 32700410799, P33 = -68118460800.0 / 10900136933,
                 P34 = 87487479700.0 / 32700410799;
MM_COEF P42 = -1754552775.0 / 470086768, P43 = 14199869525.0 / 1410260304,
                 P44 = -10690763975.0 / 1880347072;
MM_COEF P52 = 127303824393.0 / 49829197408, P53 = -318862633887.0 / 49829197408,
                 P54 = 701980252875.0 / 199316789632;
MM_COEF P62 = -282668133.0 / 205662961, P63 = 2019193451.0 / 616988883,
                 P64 = -1453857185.0 / 822651844;
MM_COEF P72 = 40617522.0 / 29380423, P73 = -110615467.0 / 29380423, P74 = 69997945.0 / 29380423;

MM_COEF C_MINSTEP_REL = 4e-15, C_MINSTEP_ABS = 1e-290, C_ROOT_A = 0.12, C_ROOT_B = 0.2;

// ---- seeds (the only lines that differ between the device and the host build) ------------------
MM_HD double rcp_seed(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#else
    union { double d; uint64_t u; } v;
    v.d = 1.0 / x;
    v.u &= 0xFFFFFFFF00000000ull;   // what MUFU.RCP64H keeps
    return v.d;
#endif
}
MM_HD double rootm5_seed(double x) {
#if defined(__CUDA_ARCH__)
    // x (normal, inside the FP32 range) is truncated to FP32 and the FP32 result widened by moving bits:
    // the F2F conversions cost ~40 cycles of latency each, these integer operations ~5.
    const unsigned hi = (unsigned)__double2hiint(x);
    const float xf = __uint_as_float(((hi - 0x38000000u) << 3) | ((unsigned)__double2loint(x) >> 29));
    float l, rf;   // plain MUFU.LG2 / MUFU.EX2: the range fix-ups of log2f()/exp2f() would sit on the critical chain
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(xf));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(-0.2f * l));
    const unsigned rb = __float_as_uint(rf);
    return __hiloint2double((int)((rb >> 3) + 0x38000000u), (int)(rb << 29));
#else
    return (double)powf((float)x, -0.2f);
#endif
}

// 1/x, error below one ulp for normal x: r0*(1+e+e^2) with e = 1 - x*r0  (|e| < 2^-20)
MM_HD double rcp64(double x) {
    const double r0 = rcp_seed(x);
    const double e = fma(-x, r0, 1.0);
    return fma(r0, fma(e, e, e), r0);
}
// c*S/(Km+S): the reciprocal correction is folded into the product so the chain after the seed is
// three FMAs deep.
MM_HD double mm_rate(double c, double Km, double S) {
    const double den = Km + S;
    const double r0 = rcp_seed(den);
    const double q0 = (c * S) * r0;
    const double e = fma(-den, r0, 1.0);
    return fma(q0, fma(e, e, e), q0);
}
// x^(-1/5) for x in [1e-30, 1e30]: seed r (relative error ~2^-20), then with e = 1 - x*r^5
// x^(-1/5) = r*(1-e)^(-1/5) = r*(1 + e/5 + 3e^2/25 + O(e^3)).
MM_HD double rootm5(double x) {
    const double r = rootm5_seed(x);
    const double r2 = r * r, r4 = r2 * r2;
    const double e = fma(-x * r, r4, 1.0);
    return fma(r, e * fma(C_ROOT_A, e, C_ROOT_B), r);
}

// SAFETY * x^(-1/5), the safety factor folded into the seed and the correction evaluated two operations deep
// (r9*e and the polynomial in parallel).  Meaningful for x inside the FP32 range; callers only use it there.
MM_HD double safety_rootm5(double x) {
    const double r = rootm5_seed(x);
    const double r9 = SAFETY * r;
    const double r2 = r * r, r4 = r2 * r2;
    const double e = fma(-x * r, r4, 1.0);
    return fma(r9 * e, fma(C_ROOT_A, e, C_ROOT_B), r9);
}

MM_HD double ulp10(double t) {
    // 10 * |nextafter(t, +inf) - t|   (rk.py:113, direction = +1)
    union { double d; int64_t i; } v;
    v.d = t;
    v.i += (t >= 0.0) ? 1 : -1;
    if (t == 0.0) v.i = 1;
    return 10.0 * fabs(v.d - t);
}

enum Status { RUNNING = 0, DONE = 1, FAILED = 2, CUT = 3 };

// Observation grid of one experiment as the solver reads it: obs[i] = (P_obs[i], t[i+1]) with
// t[n_t] = +inf, so emitting observation i and looking at the next time is one 16-byte load and the
// end of the grid needs no index test.
struct alignas(16) ObsPair {
    double P, t_next;
};
MM_HD void fill_pairs(ObsPair* obs, const double* t, const double* P, int n_t, int i) {
    obs[i].P = P[i];
    obs[i].t_next = (i + 1 < n_t) ? t[i + 1] : INFINITY;
}

struct Solve {
    double nVmax, Km;   // -Vmax, Km of the particle
    double S0;          // initial substrate of the experiment
    double t, y, f;     // current time, state, f(t, y) (first-same-as-last)
    double t_next;      // t[i_eval], the next observation time (+inf after the last)
    double t_bound;     // t[n_t-1]
    double h_abs;       // next step size
    double ssr;         // residual sum of squares so far
    double cut_lim;     // stop (CUT) as soon as ssr exceeds this; +inf = never
    int i_eval;         // next observation time to emit
    int rejected;       // the last attempt was rejected: this one retries the same scipy step (rk.py:150-170)
};

// select_initial_step (common.py:110-134) for n = 1, direction = +1, order 4.  Returns false when no
// positive step size results (Km + S0 = 0 gives NaN, with which scipy would loop forever).
MM_HD bool setup(Solve& s, double t0, double t_bound) {
    s.t = t0;
    s.t_bound = t_bound;
    s.y = s.S0;
    s.ssr = 0.0;
    s.i_eval = 0;
    s.rejected = 0;
    s.t_next = t0;
    const double y = s.S0;
    const double f = mm_rate(s.nVmax, s.Km, y);
    s.f = f;
    const double interval = fabs(t_bound - t0);
    const double iscale = rcp64(fma(fabs(y), RTOL, ATOL));
    const double d0 = fabs(y * iscale);
    const double d1 = fabs(f * iscale);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (0.01 * d0) * rcp64(d1);
    h0 = (interval < h0) ? interval : h0;
    const double y1 = fma(h0, f, y);
    const double f1 = mm_rate(s.nVmax, s.Km, y1);
    const double d2 = fabs((f1 - f) * iscale) * rcp64(h0);
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) {
        h1 = h0 * 1e-3;
        h1 = (h1 > 1e-6) ? h1 : 1e-6;
    } else {
        const double dm = ((d2 > d1) ? d2 : d1) * 100.0;   // (0.01/max(d1,d2))^(1/5) = (100 max)^(-1/5)
        h1 = (dm > 1e-30 && dm < 1e30) ? rootm5(dm) : pow(dm, -0.2);
    }
    double hh = 100 * h0;
    hh = (h1 < hh) ? h1 : hh;
    hh = (interval < hh) ? interval : hh;
    s.h_abs = hh;
    return hh > 0.0;   // false for NaN
}

// One attempted step (rk.py:111-176).  obs: the experiment's observation grid (ObsPair).  PRED: write
// P_model = S0 - S(t_eval) to pred[i] instead of accumulating residuals.  n_acc / n_rej count accepted /
// rejected attempts.
template <bool PRED>
MM_HD int attempt(Solve& s, const ObsPair* obs, double* pred, unsigned& n_acc, unsigned& n_rej) {
    const double t = s.t, y = s.y;
    const double t_bound = s.t_bound;
    double ha = s.h_abs;
    const bool rejected = s.rejected != 0;
    // t_new = t + h_abs, clipped to t_bound; h = t_new - t   (rk.py:128-134), the two cases side by side
    const double rem = t_bound - t;
    double t_new = t + ha;
    bool last = t_new - t_bound > 0;
    double h = last ? rem : t_new - t;
    if (ha < fma(fabs(t), C_MINSTEP_REL, C_MINSTEP_ABS)) {   // only then can min_step = 10*ulp(t) matter (rare)
        const double min_step = ulp10(t);
        // scipy clamps h_abs up to min_step when it enters a step; an attempt shrunk below min_step by a
        // rejection fails instead (TOO_SMALL_STEP: short solution, the reference would raise).
        if (!rejected && ha < min_step) ha = min_step;
        if (ha < min_step) return FAILED;
        t_new = t + ha;
        last = t_new - t_bound > 0;
        h = last ? rem : t_new - t;
    }
    if (last) t_new = t_bound;
    ha = fabs(h);
    const double hn = h * s.nVmax, Km = s.Km;
    const double K1 = h * s.f;
    const double K2 = mm_rate(hn, Km, fma(A21, K1, y));
    const double K3 = mm_rate(hn, Km, fma(A32, K2, fma(A31, K1, y)));
    const double K4 = mm_rate(hn, Km, fma(A43, K3, fma(A42, K2, fma(A41, K1, y))));
    const double K5 = mm_rate(hn, Km, fma(A54, K4, fma(A53, K3, fma(A52, K2, fma(A51, K1, y)))));
    const double K6 = mm_rate(hn, Km, fma(A65, K5, fma(A64, K4, fma(A63, K3, fma(A62, K2, fma(A61, K1, y))))));
    const double y_new = fma(B6, K6, fma(B5, K5, fma(B4, K4, fma(B3, K3, fma(B1, K1, y)))));
    const double k7 = mm_rate(s.nVmax, Km, y_new);
    // scale = atol + max(|y|, |y_new|)*rtol.  fmax drops a NaN where np.maximum keeps it, but a NaN y_new
    // makes the error estimate NaN on its own, so the step is rejected either way.
    const double iscale = rcp64(fma(fmax(fabs(y), fabs(y_new)), RTOL, ATOL));
    // error estimate sum_j E_j K_j; the last term as (E7*h)*k7 so that k7 enters by a single FMA
    const double ee = fma(E7 * h, k7, fma(E6, K6, fma(E5, K5, fma(E4, K4, fma(E3, K3, E1 * K1)))));
    const double err = fabs(ee * iscale);
    // 0.9*err^(-1/5): used only for ERR_LO < err < ERR_HI, where it lies strictly inside (MIN_FACTOR, MAX_FACTOR);
    // the clamps of rk.py:156-170 become selects on err, which are ready long before the root is
    const double fr = safety_rootm5(err);
    if (err < 1.0) {
        double factor = (err <= ERR_LO) ? MAX_FACTOR : fr;
        if (rejected && err < ERR_ONE) factor = 1.0;   // after a rejection the step may not grow (rk.py:158-159)
        s.h_abs = ha * factor;
        s.rejected = 0;
        n_acc++;
        // the step ends the solve iff t_new reached t_bound: clipped (last), or t + h_abs hit it exactly
        const bool done = last || t_new == t_bound;
        s.t = t_new;
        s.y = y_new;
        s.f = k7;
        // dense output for every t_eval in (t_old, t_new] (ivp.py:712-728; t_eval[0] = t0 is emitted by
        // the first step with x = 0)
        if (s.t_next <= t_new) {
            int i = s.i_eval;
            const double K7 = h * k7;
            const double q2 = fma(K7, P72, fma(K6, P62, fma(K5, P52, fma(K4, P42, fma(K3, P32, K1 * P12)))));
            const double q3 = fma(K7, P73, fma(K6, P63, fma(K5, P53, fma(K4, P43, fma(K3, P33, K1 * P13)))));
            const double q4 = fma(K7, P74, fma(K6, P64, fma(K5, P54, fma(K4, P44, fma(K3, P34, K1 * P14)))));
            const double ih = rcp64(h);
            double ssr = s.ssr, te = s.t_next;
            do {
                const double x = (te - t) * ih;
                const ObsPair o = obs[i];
                const double S = fma(x, fma(x, fma(x, fma(x, q4, q3), q2), K1), y);
                const double Pm = s.S0 - S;   // Micmem_likelihood.py:32
                if (PRED) {
                    pred[i] = Pm;
                } else {
                    const double r = o.P - Pm;   // :68
                    ssr = fma(r, r, ssr);
                }
                ++i;
                te = o.t_next;
            } while (te <= t_new);
            s.i_eval = i;
            s.t_next = te;
            s.ssr = ssr;
            if (!PRED && ssr > s.cut_lim) return CUT;   // residuals only change here
        }
        return done ? DONE : RUNNING;
    }
    const double factor = (err < ERR_HI) ? fr : MIN_FACTOR;   // NaN error: MIN_FACTOR, as Python's max()
    s.h_abs = ha * factor;
    s.rejected = 1;
    n_rej++;
    return RUNNING;
}

}  // namespace mmsolve
