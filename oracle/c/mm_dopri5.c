/*
 * mm_dopri5.c - plain-C restatement of the reference's Michaelis-Menten progress-curve likelihood.
 * TEST INFRASTRUCTURE ONLY (oracle): never linked into the product library.
 *
 * Follows, operation for operation, what the reference executes for one particle:
 *   log_likelihood_mm_multi   /root/reference/SMC_example/Micmem_likelihood.py:35-77
 *   simulate_mm_on_grid       :17-33  (scipy.integrate.solve_ivp, method="RK45", rtol=1e-3, atol=1e-6, t_eval)
 *   mm_ode                    :14-15  (dS/dt = -Vmax*S/(Km+S))
 * scipy's RK45 (scipy 1.18.1, scipy/integrate/_ivp): tableau rk.py:538-567, first step common.py:110-134,
 * step loop rk.py:111-176, stages rk.py:61-71, error rk.py:105-109, dense output rk.py:178-180,723-737 and
 * ivp.py:712-728.  It is the C twin of oracle/dopri5.py (same order of floating-point operations;
 * compile with -ffp-contract=off so no FMA is formed) and is pinned against scipy itself by
 * tests/test_oracle_c.py on the golden fixture.
 *
 * Build:  make -C oracle/c      (gcc -O2 -ffp-contract=off -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>

#define RTOL 1e-3
#define ATOL 1e-6
#define SAFETY 0.9
#define MIN_FACTOR 0.2
#define MAX_FACTOR 10.0

static const double A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
static const double P[7][4] = {
    {1.0, -8048581381.0 / 2820520608, 8663915743.0 / 2820520608, -12715105075.0 / 11282082432},
    {0, 0, 0, 0},
    {0, 131558114200.0 / 32700410799, -68118460800.0 / 10900136933, 87487479700.0 / 32700410799},
    {0, -1754552775.0 / 470086768, 14199869525.0 / 1410260304, -10690763975.0 / 1880347072},
    {0, 127303824393.0 / 49829197408, -318862633887.0 / 49829197408, 701980252875.0 / 199316789632},
    {0, -282668133.0 / 205662961, 2019193451.0 / 616988883, -1453857185.0 / 822651844},
    {0, 40617522.0 / 29380423, -110615467.0 / 29380423, 69997945.0 / 29380423}};

static inline double rhs(double Vmax, double Km, double S) { return -Vmax * S / (Km + S); }

/* Integrates one experiment; writes S(t_eval) into out[n_t]; returns number of values written
 * (n_t on success, fewer when the step size underflows, like a short sol.y).  counters: [0]+=nfev,
 * [1]+=accepted, [2]+=rejected. */
static int solve_on_grid(double Vmax, double Km, double y0, const double* t_eval, int n_t, double* out,
                         int64_t* counters) {
    double t = t_eval[0];
    const double t_bound = t_eval[n_t - 1];
    double y = y0;
    double f = rhs(Vmax, Km, y);
    int64_t nfev = 1, nacc = 0, nrej = 0;
    int i_eval = 0;
    double h_abs;
    {   /* select_initial_step, n=1, direction=+1, order=4 */
        const double interval = fabs(t_bound - t);
        if (interval == 0.0) {
            h_abs = 0.0;
        } else {
            const double scale = ATOL + fabs(y) * RTOL;
            const double d0 = fabs(y / scale), d1 = fabs(f / scale);
            double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            if (interval < h0) h0 = interval;
            const double y1 = y + h0 * f;
            const double f1 = rhs(Vmax, Km, y1);
            nfev++;
            const double d2 = fabs((f1 - f) / scale) / h0;
            double h1;
            if (d1 <= 1e-15 && d2 <= 1e-15) {
                h1 = h0 * 1e-3;
                if (h1 < 1e-6) h1 = 1e-6;
            } else {
                h1 = pow(0.01 / (d1 > d2 ? d1 : d2), 1.0 / 5.0);
            }
            h_abs = 100 * h0;
            if (h1 < h_abs) h_abs = h1;
            if (interval < h_abs) h_abs = interval;
        }
    }
    int finished = (t == t_bound);
    if (finished) {
        while (i_eval < n_t && t_eval[i_eval] <= t) out[i_eval++] = y;
    }
    double K[7];
    while (!finished) {
        const double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        if (h_abs < min_step) h_abs = min_step;
        int accepted = 0, rejected = 0;
        double t_new = t, y_new = y, f_new = f, h = 0;
        while (!accepted) {
            if (h_abs < min_step) {
                counters[0] += nfev; counters[1] += nacc; counters[2] += nrej;
                return i_eval;
            }
            h = h_abs;
            t_new = t + h;
            if (t_new - t_bound > 0) t_new = t_bound;
            h = t_new - t;
            h_abs = fabs(h);
            K[0] = f;
            for (int s = 1; s < 6; ++s) {
                double acc = 0.0;
                for (int j = 0; j < s; ++j) acc += K[j] * A[s][j];
                K[s] = rhs(Vmax, Km, y + acc * h);
            }
            double acc = 0.0;
            for (int j = 0; j < 6; ++j) acc += K[j] * B[j];
            y_new = y + h * acc;
            f_new = rhs(Vmax, Km, y_new);
            K[6] = f_new;
            nfev += 6;
            const double ay = fabs(y), ayn = fabs(y_new);
            const double scale = ATOL + ((ayn > ay || ayn != ayn) ? ayn : ay) * RTOL;   /* np.maximum propagates NaN */
            acc = 0.0;
            for (int j = 0; j < 7; ++j) acc += K[j] * E[j];
            const double err = fabs((acc * h) / scale);
            if (err < 1) {
                double factor;
                if (err == 0) factor = MAX_FACTOR;
                else {
                    factor = SAFETY * pow(err, -0.2);
                    if (!(factor < MAX_FACTOR)) factor = MAX_FACTOR;
                }
                if (rejected && !(factor < 1.0)) factor = 1.0;
                h_abs *= factor;
                accepted = 1;
            } else {
                double factor = SAFETY * pow(err, -0.2);
                if (!(factor > MIN_FACTOR)) factor = MIN_FACTOR;
                h_abs *= factor;
                rejected = 1;
                nrej++;
            }
        }
        nacc++;
        const double t_old = t, y_old = y;
        t = t_new; y = y_new; f = f_new;
        finished = (t - t_bound >= 0);
        if (i_eval < n_t && t_eval[i_eval] <= t) {
            const double hh = t - t_old;
            double Q[4];
            for (int m = 0; m < 4; ++m) {
                double acc = 0.0;
                for (int j = 0; j < 7; ++j) acc += K[j] * P[j][m];
                Q[m] = acc;
            }
            while (i_eval < n_t && t_eval[i_eval] <= t) {
                const double x = (t_eval[i_eval] - t_old) / hh;
                const double p1 = x, p2 = p1 * x, p3 = p2 * x, p4 = p3 * x;
                out[i_eval++] = hh * (Q[0] * p1 + Q[1] * p2 + Q[2] * p3 + Q[3] * p4) + y_old;
            }
        }
    }
    counters[0] += nfev; counters[1] += nacc; counters[2] += nrej;
    return i_eval;
}

/* np.sum of a contiguous float64 vector of fewer than 128 elements: NumPy's pairwise routine keeps eight
 * partial sums and combines them as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then adds the remainder in order. */
static double np_sum(const double* a, int n) {
    if (n < 8) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += a[i];
        return s;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

/* theta: [n][3] row-major (Vmax, Km, sigma); t, Pobs: [n_ex][n_t]; S0: [n_ex].
 * n_t must be below 128 for np_sum to mirror NumPy (larger vectors fall back to the same 8-lane order).
 * lk[n] out; counters (may be NULL): int64[4] += {nfev, accepted, rejected, failed solves};
 * steps_per_solve (may be NULL): int32[n][n_ex] step attempts of every solve;
 * pred (may be NULL): [n][n_ex][n_t] model predictions P_model = S0 - S. */
void mm_progress_loglik(const double* theta, int64_t n, const double* t, const double* Pobs, const double* S0,
                        int n_ex, int n_t, double* lk, int64_t* counters, int32_t* steps_per_solve, double* pred) {
    double buf[4096];
    int64_t local[4] = {0, 0, 0, 0};
    for (int64_t p = 0; p < n; ++p) {
        const double Vmax = theta[3 * p], Km = theta[3 * p + 1], sigma = theta[3 * p + 2];
        if (sigma <= 0) {          /* Micmem_likelihood.py:53-54 */
            lk[p] = -INFINITY;
            if (steps_per_solve) for (int e = 0; e < n_ex; ++e) steps_per_solve[p * n_ex + e] = 0;
            continue;
        }
        double total = 0.0;
        int ok = 1;
        for (int e = 0; e < n_ex; ++e) {
            int64_t c[3] = {0, 0, 0};
            const int got = solve_on_grid(Vmax, Km, S0[e], t + (int64_t)e * n_t, n_t, buf, c);
            local[0] += c[0]; local[1] += c[1]; local[2] += c[2];
            if (steps_per_solve) steps_per_solve[p * n_ex + e] = (int32_t)(c[1] + c[2]);
            if (got != n_t) { ok = 0; local[3]++; continue; }
            for (int i = 0; i < n_t; ++i) {
                const double Pm = S0[e] - buf[i];
                if (pred) pred[((int64_t)p * n_ex + e) * n_t + i] = Pm;
                const double r = Pobs[(int64_t)e * n_t + i] - Pm;
                buf[i] = r * r;
            }
            const double ssr = np_sum(buf, n_t);
            total += -0.5 * n_t * log(2 * M_PI * (sigma * sigma)) - ssr / (2 * (sigma * sigma));
        }
        lk[p] = ok ? total : -INFINITY;
    }
    if (counters) for (int k = 0; k < 4; ++k) counters[k] += local[k];
}
