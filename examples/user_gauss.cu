// A likelihood the library was not compiled with: y_j ~ N(mu + slope * x_j, sigma^2), parameters (mu, slope, sigma).
// Built by __graft_entry__.build() (and by examples/user_likelihood.py) into examples/libuser_gauss.so; the
// sampler reaches it through smcb_set_user_likelihood / smcb200.UserKernelLikelihood.  This is the counterpart of
// writing a new `sim_particle` for the reference (SMC_example/Micmem_likelihood.py:79-92).
#include <math.h>

#include "smcb_user.cuh"

namespace {

struct LineModel {
    const double* x;   // device
    const double* y;   // device
    int m;
    __device__ double operator()(const smcb_user::Particle& p) const {
        const double mu = p[0], slope = p[1], sigma = p[2];
        if (!(sigma > 0)) return -INFINITY;
        double ssr = 0.0;
        for (int j = 0; j < m; ++j) {
            const double r = y[j] - (mu + slope * x[j]);
            ssr = fma(r, r, ssr);
        }
        return -0.5 * m * log(2 * M_PI * sigma * sigma) - ssr / (2 * sigma * sigma);
    }
};

LineModel g_model = {nullptr, nullptr, 0};

}  // namespace

// data upload (host arrays); synchronous
extern "C" int user_gauss_set_data(const double* x_host, const double* y_host, int m) {
    double *x = nullptr, *y = nullptr;
    if (cudaMalloc((void**)&x, sizeof(double) * m) != cudaSuccess) return 1;
    if (cudaMalloc((void**)&y, sizeof(double) * m) != cudaSuccess) return 1;
    cudaMemcpy(x, x_host, sizeof(double) * m, cudaMemcpyHostToDevice);
    cudaMemcpy(y, y_host, sizeof(double) * m, cudaMemcpyHostToDevice);
    if (g_model.x) cudaFree(const_cast<double*>(g_model.x));
    if (g_model.y) cudaFree(const_cast<double*>(g_model.y));
    g_model = LineModel{x, y, m};
    return 0;
}

// smcb_user_loglik_fn
extern "C" int user_gauss_loglik(void* /*user_data*/, const double* theta_dev, int64_t ld, int64_t n, int d,
                                 const uint8_t* active_dev, double* lk_dev, void* stream) {
    if (d != 3 || g_model.m == 0) return 2;
    return smcb_user::launch(g_model, theta_dev, ld, n, d, active_dev, lk_dev, stream);
}
