/*
 * smcb_user.cuh - the few lines a user-written likelihood kernel needs (SMCB_MODEL_USER).
 *
 * The reference's plug-in surface is a user-written `sim_particle(particle) -> llk`
 * (SMC_example/Micmem_likelihood.py:79-92, SMC_methanation/methanation_functions.py:70-92): one Python function
 * evaluated per particle by a ray task.  Here the user writes a device functor
 *
 *     struct MyModel {
 *         ...data pointers (device)...
 *         __device__ double operator()(const smcb_user::Particle& p) const { return log-likelihood of p[0], p[1], ...; }
 *     };
 *
 * and exports one C function of type smcb_user_loglik_fn that calls smcb_user::launch(...) with it.  The library
 * calls that function wherever it would have called one of its own likelihood kernels (first sweep, MH sweeps with
 * their active mask); tempering, resampling and the MH mutation are unchanged.  Compile with
 *     nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -I<repo>/include my_model.cu -o libmy_model.so
 * and hand the library + symbol to `smcb200.UserKernelLikelihood` (or the function pointer to smcb_set_user_likelihood).
 */
#ifndef SMCB_USER_CUH
#define SMCB_USER_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "smcb200.h"

namespace smcb_user {

/* One particle of the SoA state: p[k] is parameter k. */
struct Particle {
    const double* theta;
    int64_t ld, i;
    int d;
    __device__ __forceinline__ double operator[](int k) const { return theta[(int64_t)k * ld + i]; }
};

template <class Model>
__global__ void __launch_bounds__(256)
loglik_kernel(const double* __restrict__ theta, int64_t ld, int64_t n, int d, const uint8_t* __restrict__ active,
              double* __restrict__ lk, const Model model) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (active != nullptr && !active[i])) return;      /* masked particles keep their old value */
    lk[i] = model(Particle{theta, ld, i, d});
}

/* One thread per particle on `stream`; returns 0, or 1 if the launch failed (-> SMCB_ERR_USER). */
template <class Model>
inline int launch(const Model& model, const double* theta_dev, int64_t ld, int64_t n, int d,
                  const uint8_t* active_dev, double* lk_dev, void* stream) {
    if (n <= 0) return 0;
    loglik_kernel<Model><<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        theta_dev, ld, n, d, active_dev, lk_dev, model);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace smcb_user
#endif
