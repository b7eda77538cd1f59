// Host build of csrc/exp_table.cuh (the device's Arrhenius exp) for tests/test_host_twin.py: no GPU needed.
#include "../python-based-sequential-monte-carlo-method-with-likelihood-tempering_b200/csrc/exp_table.cuh"

extern "C" void exp_fast_host(const double* x, double* y, long n) {
    for (long i = 0; i < n; ++i) y[i] = expt::exp_fast(x[i], expt::TAB);
}
