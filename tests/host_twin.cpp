// Host build of the DEVICE solver arithmetic (csrc/mm_solver.cuh) for tests only.
//
// The kernels of loglik_mm.cu call mmsolve::setup / mmsolve::attempt; this file compiles the very same
// header with g++ (FMA instructions on, no contraction of anything that is not spelt fma()) so that the
// arithmetic the GPU performs can be compared with scipy without a GPU (tests/test_host_twin.py).  It is
// test infrastructure: nothing in the product loads it.
#include <math.h>
#include <stdint.h>

#include <vector>

#include "../python-based-sequential-monte-carlo-method-with-likelihood-tempering_b200/csrc/mm_solver.cuh"

// form 0: attempt() alone (the bulk kernel), form 1: solve_lat() (the tail kernel)
template <int FORM>
static void loglik_form(const double* theta, int64_t n, const double* t, const double* P, const double* S0,
                        int n_ex, int n_t, double* lk, int64_t* counters, int32_t* steps) {
    using namespace mmsolve;
    std::vector<ObsPair> obs((size_t)n_ex * n_t);
    for (int e = 0; e < n_ex; ++e)
        for (int i = 0; i < n_t; ++i) fill_pairs(obs.data() + (size_t)e * n_t, t + e * n_t, P + e * n_t, n_t, i);
    for (int64_t p = 0; p < n; ++p) {
        const double Vmax = theta[3 * p], Km = theta[3 * p + 1], sigma = theta[3 * p + 2];
        if (sigma <= 0) {
            lk[p] = -INFINITY;
            continue;
        }
        const double s2 = sigma * sigma;
        const double c0 = -0.5 * n_t * log(2 * M_PI * s2);
        const double inv_den = 1.0 / (2 * s2);
        double total = 0.0;
        for (int e = 0; e < n_ex; ++e) {
            Solve s;
            s.nVmax = -Vmax;
            s.Km = Km;
            s.S0 = S0[e];
            s.cut_lim = INFINITY;
            unsigned n_acc = 0, n_rej = 0;
            int st = setup(s, t[e * n_t], t[e * n_t + n_t - 1]) ? RUNNING : FAILED;
            if (FORM) {
                if (st == RUNNING) st = solve_lat<false>(s, obs.data() + (size_t)e * n_t, n_acc, n_rej);
            } else {
                while (st == RUNNING) st = attempt<false>(s, obs.data() + (size_t)e * n_t, nullptr, n_acc, n_rej);
            }
            if (counters) {
                counters[0] += 2 + 6 * (int64_t)(n_acc + n_rej);
                counters[1] += n_acc;
                counters[2] += n_rej;
                counters[3] += (st == FAILED);
            }
            if (steps) steps[p * n_ex + e] = (int32_t)(n_acc + n_rej);
            total += (st == DONE) ? c0 - s.ssr * inv_den : -INFINITY;
        }
        lk[p] = total;
    }
}

extern "C" void twin_loglik(const double* theta, int64_t n, const double* t, const double* P, const double* S0,
                            int n_ex, int n_t, double* lk, int64_t* counters, int32_t* steps) {
    loglik_form<0>(theta, n, t, P, S0, n_ex, n_t, lk, counters, steps);
}
extern "C" void twin_loglik_lat(const double* theta, int64_t n, const double* t, const double* P, const double* S0,
                                int n_ex, int n_t, double* lk, int64_t* counters, int32_t* steps) {
    loglik_form<1>(theta, n, t, P, S0, n_ex, n_t, lk, counters, steps);
}

// The hand-over between the two kernels: attempt() for at most `budget` attempts (mm_bulk_kernel), then the solve is
// parked (park_store), restored into a fresh Solve (park_load) and finished by solve_lat (mm_tail_kernel).
extern "C" void twin_loglik_parked(const double* theta, int64_t n, const double* t, const double* P, const double* S0,
                                   int n_ex, int n_t, int budget, double* lk, int64_t* counters, int32_t* steps) {
    using namespace mmsolve;
    std::vector<ObsPair> obs((size_t)n_ex * n_t);
    for (int e = 0; e < n_ex; ++e)
        for (int i = 0; i < n_t; ++i) fill_pairs(obs.data() + (size_t)e * n_t, t + e * n_t, P + e * n_t, n_t, i);
    for (int64_t p = 0; p < n; ++p) {
        const double Vmax = theta[3 * p], Km = theta[3 * p + 1], sigma = theta[3 * p + 2];
        const double s2 = sigma * sigma;
        const double c0 = -0.5 * n_t * log(2 * M_PI * s2);
        const double inv_den = 1.0 / (2 * s2);
        double total = 0.0;
        for (int e = 0; e < n_ex; ++e) {
            const ObsPair* ob = obs.data() + (size_t)e * n_t;
            Solve s;
            s.nVmax = -Vmax;
            s.Km = Km;
            s.S0 = S0[e];
            s.cut_lim = INFINITY;
            unsigned n_acc = 0, n_rej = 0, n_att = 0;
            int st = setup(s, t[e * n_t], t[e * n_t + n_t - 1]) ? RUNNING : FAILED;
            while (st == RUNNING && n_att < (unsigned)budget) {
                st = attempt<false>(s, ob, nullptr, n_acc, n_rej);
                ++n_att;
            }
            if (st == RUNNING) {
                double rec[PARK_WORDS];
                park_store(rec, s, n_att);
                Solve r;
                r.nVmax = -Vmax;
                r.Km = Km;
                r.S0 = S0[e];
                r.cut_lim = INFINITY;
                const unsigned parked = park_load(rec, r, ob, t[e * n_t + n_t - 1]);
                if (parked != n_att) counters[3] += 1000;      // the record lost the attempt count
                st = solve_lat<false>(r, ob, n_acc, n_rej);
                s = r;
            }
            counters[1] += n_acc;
            counters[2] += n_rej;
            counters[3] += (st == FAILED);
            steps[p * n_ex + e] = (int32_t)(n_acc + n_rej);
            total += (st == DONE) ? c0 - s.ssr * inv_den : -INFINITY;
        }
        lk[p] = total;
    }
}

extern "C" void twin_predict(const double* theta, int64_t n, const double* t, const double* S0, int n_ex,
                             int n_t, double* pred) {
    using namespace mmsolve;
    std::vector<ObsPair> obs((size_t)n_ex * n_t);
    std::vector<double> zero((size_t)n_t, 0.0);
    for (int e = 0; e < n_ex; ++e)
        for (int i = 0; i < n_t; ++i) fill_pairs(obs.data() + (size_t)e * n_t, t + e * n_t, zero.data(), n_t, i);
    for (int64_t p = 0; p < n; ++p)
        for (int e = 0; e < n_ex; ++e) {
            Solve s;
            s.nVmax = -theta[3 * p];
            s.Km = theta[3 * p + 1];
            s.S0 = S0[e];
            s.cut_lim = INFINITY;
            unsigned a = 0, r = 0;
            int st = setup(s, t[e * n_t], t[e * n_t + n_t - 1]) ? RUNNING : FAILED;
            while (st == RUNNING) st = attempt<true>(s, obs.data() + (size_t)e * n_t, pred + (p * n_ex + e) * n_t, a, r);
        }
}
