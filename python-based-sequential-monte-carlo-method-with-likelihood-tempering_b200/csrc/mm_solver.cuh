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
// c ? a : b as a single select that the optimiser cannot re-distribute
MM_HD double select_late(bool c, double a, double b) {
#if defined(__CUDA_ARCH__)
    double o;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(o) : "d"(a), "d"(b), "r"((int)c));
    return o;
#else
    return c ? a : b;
#endif
}

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


// ---------------------------------------------------------------------------------------------------------
// solve_lat: a whole solve, spelt for LATENCY (mm_tail_kernel: one long solve per lane, 1e3 .. 1e5 strictly
// sequential attempts and nothing to overlap them with; attempt() above is spelt for instruction count, which is
// what bounds mm_bulk_kernel).  Steps that need anything beyond the plain RK45 attempt - a step that reaches an
// observation time or t_bound, a step at the min_step floor, a step that exploded - are handed to attempt() as
// they are.  All others (all but ~50 of the 1e3 .. 1e5 attempts of a stiff solve) run in a loop whose dependent chain
// is as short as the arithmetic allows.  Same tableau, same controller; what differs from attempt() is where roundings
// fall, never by more than the ulp or two that already separate either spelling from scipy's operation order:
//   * a stage's rate is K = q + q*p with q = (hn*S)*r0 and p = e + e^2 the reciprocal correction.  The NEXT stage's
//     denominator Km + y + sum a_j K_j is formed as fma(a*q, p, fma(a, q, Km + partial sum)): it needs only p from
//     the chain - not K, not the stage argument, no separate Km + S - so a stage costs MUFU.RCP64H -> e -> p -> den
//     (~40 cycles) instead of ... -> K -> S -> den (~56);
//   * the error estimate takes k7 the same way (one FMA after p7);
//   * err^(-1/5): the FP32 seed is 2^(-0.2*(lg2|ee| - lg2 scale)), so the logarithm starts as soon as ee is known
//     and 1/scale (needed only by the FP64 correction) is off the chain;
//   * the new step size is fma((0.9*ha*r)*e, poly(e), 0.9*ha*r) instead of ha*(0.9*root), chosen against the
//     clamped alternatives by ONE select;
//   * accept / reject is a select on the state, and the loop is rotated: t + h, the first stage and the test
//     "is the next step a plain one" are issued at the end of the previous step, so the only branch of a step is
//     the loop's own and its predicate is ready before it is needed.
// A solve takes the same accepted / rejected steps as with attempt() alone (tests/test_host_twin.py compares both
// spellings with the C twin of scipy's RK45 solve by solve, the stiffest solves of the bench's prior cloud included);
// residual sums differ by rounding only.  Which spelling performs a step is a function of the solve's own state, so
// results are reproducible and independent of sharding, grid and scheduling.
struct QP {
    double q, p;
};
// q = (cS)/den to 20 bits and the correction p = e + e^2, e = 1 - den*r0:  cS/den = q + q*p  (error < 1 ulp)
MM_HD QP rate_qp(double cS, double den) {
    const double r0 = rcp_seed(den);
    QP o;
    o.q = cS * r0;
    const double e = fma(-den, r0, 1.0);
    o.p = fma(e, e, e);
    return o;
}
// lg2 of |x| truncated to FP32 by moving bits (the sign leaves through the shift); garbage outside the FP32 range
MM_HD float lg2_trunc(double x) {
#if defined(__CUDA_ARCH__)
    const unsigned hi = (unsigned)__double2hiint(x);
    const float xf = __uint_as_float(((hi - 0x38000000u) << 3) | ((unsigned)__double2loint(x) >> 29));
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(xf));
    return l;
#else
    return log2f((float)fabs(x));
#endif
}
MM_HD double ex2_widen(float a) {
#if defined(__CUDA_ARCH__)
    float rf;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(a));
    const unsigned rb = __float_as_uint(rf);
    return __hiloint2double((int)((rb >> 3) + 0x38000000u), (int)(rb << 29));
#else
    return (double)exp2f(a);
#endif
}

MM_HD double mul_early(double a, double b) {
#if defined(__CUDA_ARCH__)
    double o;
    asm("mul.f64 %0, %1, %2;" : "=d"(o) : "d"(a), "d"(b));
    return o;
#else
    return a * b;
#endif
}

// The coefficients of solve_lat's loop, in the order the loop names them.  HOIST: the caller has copied them to
// `tab` (any memory that is not constant memory, e.g. shared) with lat_coef_fill and the loop keeps them in registers:
// read through a volatile pointer so that the loads can neither be folded back into constant operands nor be repeated
// inside the loop.
constexpr int LAT_NCOEF = 36;
MM_HD void lat_coef_fill(double* tab) {
    tab[0] = A21;
    tab[1] = A31;
    tab[2] = A32;
    tab[3] = A41;
    tab[4] = A42;
    tab[5] = A43;
    tab[6] = A51;
    tab[7] = A52;
    tab[8] = A53;
    tab[9] = A54;
    tab[10] = A61;
    tab[11] = A62;
    tab[12] = A63;
    tab[13] = A64;
    tab[14] = A65;
    tab[15] = B1;
    tab[16] = B3;
    tab[17] = B4;
    tab[18] = B5;
    tab[19] = B6;
    tab[20] = E1;
    tab[21] = E3;
    tab[22] = E4;
    tab[23] = E5;
    tab[24] = E6;
    tab[25] = E7;
    tab[26] = RTOL;
    tab[27] = ATOL;
    tab[28] = SAFETY;
    tab[29] = MAX_FACTOR;
    tab[30] = MIN_FACTOR;
    tab[31] = ERR_LO;
    tab[32] = ERR_ONE;
    tab[33] = ERR_HI;
    tab[34] = C_ROOT_A;
    tab[35] = C_ROOT_B;
}
template <bool HOIST>
MM_HD double coef(double c, const double* tab, int i) {
    return HOIST ? *reinterpret_cast<const volatile double*>(tab + i) : c;
}

// Returns DONE, FAILED or CUT.  s: a solve after setup() (or after any number of attempts).
// HOIST: keep the 36 coefficients of the loop in registers (the tail kernel has registers to spare) instead of
// re-reading them from constant memory in every step; tab: see lat_coef_fill.
template <bool HOIST>
MM_HD int solve_lat(Solve& s, const ObsPair* obs, unsigned& n_acc, unsigned& n_rej, const double* tab = nullptr) {
    const double Km = s.Km, nVmax = s.nVmax, t_bound = s.t_bound;
    // a step size at or above this is above attempt()'s min_step test at every t of the solve
    const double h_floor = fma(fmax(fabs(s.t), fabs(t_bound)), C_MINSTEP_REL, C_MINSTEP_ABS);
    // loop coefficients (registers if HOIST)
    const double c_a21 = coef<HOIST>(A21, tab, 0), c_a31 = coef<HOIST>(A31, tab, 1), c_a32 = coef<HOIST>(A32, tab, 2), c_a41 = coef<HOIST>(A41, tab, 3);
    const double c_a42 = coef<HOIST>(A42, tab, 4), c_a43 = coef<HOIST>(A43, tab, 5), c_a51 = coef<HOIST>(A51, tab, 6), c_a52 = coef<HOIST>(A52, tab, 7);
    const double c_a53 = coef<HOIST>(A53, tab, 8), c_a54 = coef<HOIST>(A54, tab, 9), c_a61 = coef<HOIST>(A61, tab, 10), c_a62 = coef<HOIST>(A62, tab, 11);
    const double c_a63 = coef<HOIST>(A63, tab, 12), c_a64 = coef<HOIST>(A64, tab, 13), c_a65 = coef<HOIST>(A65, tab, 14), c_b1 = coef<HOIST>(B1, tab, 15);
    const double c_b3 = coef<HOIST>(B3, tab, 16), c_b4 = coef<HOIST>(B4, tab, 17), c_b5 = coef<HOIST>(B5, tab, 18), c_b6 = coef<HOIST>(B6, tab, 19);
    const double c_e1 = coef<HOIST>(E1, tab, 20), c_e3 = coef<HOIST>(E3, tab, 21), c_e4 = coef<HOIST>(E4, tab, 22), c_e5 = coef<HOIST>(E5, tab, 23);
    const double c_e6 = coef<HOIST>(E6, tab, 24), c_e7 = coef<HOIST>(E7, tab, 25), c_rtol = coef<HOIST>(RTOL, tab, 26), c_atol = coef<HOIST>(ATOL, tab, 27);
    const double c_safety = coef<HOIST>(SAFETY, tab, 28), c_max_factor = coef<HOIST>(MAX_FACTOR, tab, 29), c_min_factor = coef<HOIST>(MIN_FACTOR, tab, 30), c_err_lo = coef<HOIST>(ERR_LO, tab, 31);
    const double c_err_one = coef<HOIST>(ERR_ONE, tab, 32), c_err_hi = coef<HOIST>(ERR_HI, tab, 33), c_c_root_a = coef<HOIST>(C_ROOT_A, tab, 34), c_c_root_b = coef<HOIST>(C_ROOT_B, tab, 35);
    for (;;) {
        // plain step: not tiny, and t + h_abs lies before the next observation time and before t_bound
        // (then attempt() would neither clip the step, nor finish, nor emit an observation)
        const double lim = fmin(s.t_next, t_bound);
        double t = s.t, y = s.y, f = s.f, ha = s.h_abs;
        double tn = t + ha;
        if (tn < lim) {
            bool rej = s.rejected != 0, plain;
            unsigned nt = 0, na = 0;
            double h = tn - t;
            double K1 = h * f, hn = h * nVmax;
            double d2 = fma(c_a21, K1, Km + y), cS2 = hn * fma(c_a21, K1, y);
            do {
                // ---- stages 2..6, y_new, f(y_new)
                // the candidates of the next step size that need no root (opaque products: the optimiser would fold
                // them into one late multiplication by a conditionally loaded constant)
                const double ha_in = ha;
                const double haS = mul_early(h, c_safety), ha_max = mul_early(h, c_max_factor), ha_min = mul_early(h, c_min_factor);
                const QP s2 = rate_qp(cS2, d2);
                const double y3 = fma(c_a31, K1, y);
                const double d3 = fma(c_a32 * s2.q, s2.p, fma(c_a32, s2.q, Km + y3));
                const double K2 = fma(s2.q, s2.p, s2.q);
                const QP s3 = rate_qp(hn * fma(c_a32, K2, y3), d3);
                const double y4 = fma(c_a42, K2, fma(c_a41, K1, y));
                const double d4 = fma(c_a43 * s3.q, s3.p, fma(c_a43, s3.q, Km + y4));
                const double K3 = fma(s3.q, s3.p, s3.q);
                const QP s4 = rate_qp(hn * fma(c_a43, K3, y4), d4);
                const double y5 = fma(c_a53, K3, fma(c_a52, K2, fma(c_a51, K1, y)));
                const double d5 = fma(c_a54 * s4.q, s4.p, fma(c_a54, s4.q, Km + y5));
                const double K4 = fma(s4.q, s4.p, s4.q);
                const QP s5 = rate_qp(hn * fma(c_a54, K4, y5), d5);
                const double y6 = fma(c_a64, K4, fma(c_a63, K3, fma(c_a62, K2, fma(c_a61, K1, y))));
                const double d6 = fma(c_a65 * s5.q, s5.p, fma(c_a65, s5.q, Km + y6));
                const double K5 = fma(s5.q, s5.p, s5.q);
                const QP s6 = rate_qp(hn * fma(c_a65, K5, y6), d6);
                const double yn = fma(c_b5, K5, fma(c_b4, K4, fma(c_b3, K3, fma(c_b1, K1, y))));
                const double d7 = fma(c_b6 * s6.q, s6.p, fma(c_b6, s6.q, Km + yn));
                const double K6 = fma(s6.q, s6.p, s6.q);
                const double y_new = fma(c_b6, K6, yn);
                const QP s7 = rate_qp(nVmax * y_new, d7);
                const double k7 = fma(s7.q, s7.p, s7.q);
                // ---- error estimate sum_j E_j K_j, the last term (c_e7*h)*k7 entering through p7
                const double pe = fma(c_e6, K6, fma(c_e5, K5, fma(c_e4, K4, fma(c_e3, K3, c_e1 * K1))));
                const double E7h = c_e7 * h;
                const double ee = fma(E7h * s7.q, s7.p, fma(E7h, s7.q, pe));
                const double scale = fma(fmax(fabs(y), fabs(y_new)), c_rtol, c_atol);
                const float l_scale = lg2_trunc(scale);
                const double err = fabs(ee * rcp64(scale));
                // ---- h * 0.9 * err^(-1/5)
                const double r = ex2_widen(fmaf(lg2_trunc(ee), -0.2f, 0.2f * l_scale));
                const double r2 = r * r, r4 = r2 * r2;
                const double e = fma(-err * r, r4, 1.0);
                const double har = haS * r;
                const double h_fr = fma(har * e, fma(c_c_root_a, e, c_c_root_b), har);
                // ---- rk.py:156-170.  accepted: min(c_max_factor, .), no growth right after a rejection;
                //      rejected: max(c_min_factor, .), a NaN error estimate included (Python's max()).
                // bad: the step exploded out of the range in which the FP32 seed means anything, or it was entered with
                // a step size at attempt()'s min_step floor (tested here, in the shadow of the stages, not between two
                // steps); nothing of such a step is kept and attempt() repeats it.
                const bool bad = !(scale < 1e30) || ha_in < h_floor;
                const bool acc = err < 1.0 && !bad;
                const bool use_max = acc && err <= c_err_lo;
                const bool use_one = acc && rej && err < c_err_one;
                const bool use_min = !acc && !(err < c_err_hi);
                double h_other = use_max ? ha_max : ha_min;
                h_other = use_one ? h : h_other;
                h_other = bad ? ha_in : h_other;
                // ONE select after the root (the optimiser otherwise spreads the conditions over a cascade of five)
                ha = select_late(use_max || use_one || use_min || bad, h_other, h_fr);
                nt += bad ? 0u : 1u;
                na += acc ? 1u : 0u;
                rej = bad ? rej : !acc;
                t = acc ? tn : t;
                y = acc ? y_new : y;
                f = acc ? k7 : f;
                // ---- head of the next step
                tn = t + ha;
                h = tn - t;
                K1 = h * f;
                hn = h * nVmax;
                d2 = fma(c_a21, K1, Km + y);
                cS2 = hn * fma(c_a21, K1, y);
                plain = tn < lim && !bad;
            } while (plain);
            s.t = t;
            s.y = y;
            s.f = f;
            s.h_abs = ha;
            s.rejected = rej ? 1 : 0;
            n_acc += na;
            n_rej += nt - na;
        }
        const int st = attempt<false>(s, obs, nullptr, n_acc, n_rej);
        if (st != RUNNING) return st;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Handing a running solve from one kernel to another (mm_bulk_kernel parks a solve that exceeds its budget,
// mm_tail_kernel resumes it): everything setup() and the attempts so far have put into the Solve, in six words.
constexpr int PARK_WORDS = 6;       // t, y, f, h_abs, ssr, (i_eval | rejected << 16 | attempts << 32)

MM_HD double bits_to_double(uint64_t w) {
    union { double d; uint64_t u; } v;
    v.u = w;
    return v.d;
}
MM_HD uint64_t double_to_bits(double d) {
    union { double d; uint64_t u; } v;
    v.d = d;
    return v.u;
}
MM_HD void park_store(double* rec, const Solve& s, unsigned n_att) {
    rec[0] = s.t;
    rec[1] = s.y;
    rec[2] = s.f;
    rec[3] = s.h_abs;
    rec[4] = s.ssr;
    rec[5] = bits_to_double((uint64_t)(unsigned)s.i_eval | ((uint64_t)(s.rejected != 0) << 16) | ((uint64_t)n_att << 32));
}
// restores what park_store saved (nVmax, Km, S0, cut_lim are the caller's, as before setup()); returns the attempts
// already made
MM_HD unsigned park_load(const double* rec, Solve& s, const ObsPair* obs, double t_bound) {
    s.t = rec[0];
    s.y = rec[1];
    s.f = rec[2];
    s.h_abs = rec[3];
    s.ssr = rec[4];
    const uint64_t w = double_to_bits(rec[5]);
    s.i_eval = (int)(w & 0xffffu);
    s.rejected = (int)((w >> 16) & 1u);
    s.t_bound = t_bound;
    // next observation time: the pair before it carries it (obs[i].t_next = t[i+1]); nothing emitted yet: t[0] = t0 <= t
    s.t_next = (s.i_eval > 0) ? obs[s.i_eval - 1].t_next : s.t;
    return (unsigned)(w >> 32);
}

}  // namespace mmsolve
