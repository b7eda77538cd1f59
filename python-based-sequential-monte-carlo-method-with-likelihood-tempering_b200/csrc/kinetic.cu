// K1': kinetic reactor log-likelihood (one thread per (particle, condition)) and the fused
// multi-sweep Metropolis-Hastings kernel (one thread per particle, state in registers).
// Replaces cal_parallel_new / my_model / my_loglike (SMC_methanation/methanation_functions.py:44-65,
// methanation_set_likelihood.py:144-300) with the fixed-step RK4 reactor defined in kinetic.cuh.
#include "common.cuh"
#include "kinetic.cuh"
#include "philox.cuh"

namespace {

constexpr int KB = 128;

// assemble the full parameter vector of particle p: estimated positions from theta, others from base
template <int M>
__device__ __forceinline__ void load_kin(const double* __restrict__ theta, int64_t ld, int64_t p,
                                         const double* __restrict__ base, const int* __restrict__ inv_pos,
                                         kin::Kin<M>& K, double* sigma) {
#pragma unroll
    for (int j = 0; j < 4 * M; ++j) {
        const int ia = inv_pos[2 * j], ie = inv_pos[2 * j + 1];
        K.A[j] = (ia >= 0) ? theta[(int64_t)ia * ld + p] : base[2 * j];
        const double E = (ie >= 0) ? theta[(int64_t)ie * ld + p] : base[2 * j + 1];
        K.nEoR[j] = -E / kin::R_GAS;
    }
    const int is = inv_pos[8 * M];
    *sigma = (is >= 0) ? theta[(int64_t)is * ld + p] : base[8 * M];
}

template <int M>
__global__ void __launch_bounds__(KB)
kinetic_ssr_kernel(const double* __restrict__ theta, int64_t ld, int64_t n, const uint8_t* __restrict__ active,
                   const double* __restrict__ cond, const double* __restrict__ obs, int n_cond, int n_steps,
                   const double* __restrict__ base, const int* __restrict__ inv_pos, double* __restrict__ ssr) {
    const int64_t p = (int64_t)blockIdx.x * KB + threadIdx.x;
    const int c = blockIdx.y;
    if (p >= n) return;
    if (active != nullptr && !active[p]) return;
    kin::Kin<M> K;
    double sigma;
    load_kin<M>(theta, ld, p, base, inv_pos, K, &sigma);
    ssr[(int64_t)c * n + p] =
        kin::condition_ssr<M>(K, cond + (int64_t)c * SMCB_KIN_NCOND_FIELDS, n_steps, obs, n_cond, c);
}

// lk = sum_k [ -(0.5/sigma^2) * sum_c r_kc^2 - n_cond*log(sigma) ]   (set_likelihood.py:289-298)
__global__ void kinetic_finalize_kernel(const double* __restrict__ theta, int64_t ld, int64_t n,
                                        const uint8_t* __restrict__ active, const double* __restrict__ ssr,
                                        int n_cond, const double* __restrict__ base, const int* __restrict__ inv_pos,
                                        int sigma_pos, double* __restrict__ lk) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (active != nullptr && !active[p]) return;
    const int is = inv_pos[sigma_pos];
    const double sigma = (is >= 0) ? theta[(int64_t)is * ld + p] : base[sigma_pos];
    double tot = 0.0;
    for (int c = 0; c < n_cond; ++c) tot += ssr[(int64_t)c * n + p];
    lk[p] = -(0.5 / (sigma * sigma)) * tot - 5.0 * n_cond * log(sigma);
}

// ---- fused multi-sweep MH ------------------------------------------------------------------------
struct FusedParams {
    double F[SMCB_MAX_DIM * SMCB_MAX_DIM];
    double low[SMCB_MAX_DIM];
    double high[SMCB_MAX_DIM];
};

template <int M>
__device__ __forceinline__ double kinetic_loglik_regs(const double* th, int d, const double* __restrict__ base,
                                                      const int* __restrict__ inv_pos,
                                                      const double* __restrict__ cond,
                                                      const double* __restrict__ obs, int n_cond, int n_steps) {
    kin::Kin<M> K;
#pragma unroll
    for (int j = 0; j < 4 * M; ++j) {
        const int ia = inv_pos[2 * j], ie = inv_pos[2 * j + 1];
        K.A[j] = (ia >= 0) ? th[ia] : base[2 * j];
        const double E = (ie >= 0) ? th[ie] : base[2 * j + 1];
        K.nEoR[j] = -E / kin::R_GAS;
    }
    const int is = inv_pos[8 * M];
    const double sigma = (is >= 0) ? th[is] : base[8 * M];
    double tot = 0.0;
    for (int c = 0; c < n_cond; ++c)
        tot += kin::condition_ssr<M>(K, cond + (int64_t)c * SMCB_KIN_NCOND_FIELDS, n_steps, obs, n_cond, c);
    return -(0.5 / (sigma * sigma)) * tot - 5.0 * n_cond * log(sigma);
}

template <int M>
__global__ void __launch_bounds__(KB)
kinetic_mh_fused_kernel(double* __restrict__ theta, int64_t ld, double* __restrict__ lk, int64_t n, int d,
                        const __grid_constant__ FusedParams prm, double ratio, double gamma, int n_sweeps,
                        uint64_t seed, uint64_t id_offset, uint32_t stage, uint32_t sweep0,
                        const double* __restrict__ cond, const double* __restrict__ obs, int n_cond, int n_steps,
                        const double* __restrict__ base, const int* __restrict__ inv_pos,
                        uint8_t* __restrict__ moved, unsigned long long* __restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * KB + threadIdx.x;
    long long n_acc = 0, n_new = 0, n_eval = 0;
    if (i < n) {
        double th[SMCB_MAX_DIM], pr[SMCB_MAX_DIM];
        for (int k = 0; k < d; ++k) th[k] = theta[(int64_t)k * ld + i];
        double l1 = lk[i];
        bool mv = moved[i] != 0;
        const bool mv0 = mv;
        for (int s = 0; s < n_sweeps; ++s) {
            const uint32_t sweep = sweep0 + (uint32_t)s;
            for (int k = 0; k < d; ++k) pr[k] = 0.0;
            for (int j = 0; j < d; j += 2) {
                double z0, z1;
                philox_normal2(seed, id_offset + (uint64_t)i, stage, sweep, (uint32_t)(j >> 1), &z0, &z1);
                for (int k = 0; k < d; ++k) pr[k] += z0 * prm.F[j * d + k];
                if (j + 1 < d)
                    for (int k = 0; k < d; ++k) pr[k] += z1 * prm.F[(j + 1) * d + k];
            }
            bool ok = true;
            for (int k = 0; k < d; ++k) {
                pr[k] = th[k] + pr[k] * ratio;
                ok = ok && pr[k] >= prm.low[k] && pr[k] <= prm.high[k];
            }
            const double u = philox_uniform(seed, id_offset + (uint64_t)i, stage, sweep, SMCB_SLOT_UNIFORM);
            double pp = 0.0, l2 = 0.0;
            if (ok) {
                l2 = kinetic_loglik_regs<M>(pr, d, base, inv_pos, cond, obs, n_cond, n_steps);
                ++n_eval;
                pp = exp((l2 - l1) * gamma);
            }
            if (pp >= u) {
                ++n_acc;
                if (ok) {
                    for (int k = 0; k < d; ++k) th[k] = pr[k];
                    l1 = l2;
                }
                mv = true;
            }
        }
        for (int k = 0; k < d; ++k) theta[(int64_t)k * ld + i] = th[k];
        lk[i] = l1;
        if (mv && !mv0) {
            moved[i] = 1;
            n_new = 1;
        }
    }
    n_acc = warp_sum_ll(n_acc);
    n_new = warp_sum_ll(n_new);
    n_eval = warp_sum_ll(n_eval);
    if ((threadIdx.x & 31) == 0) {
        if (n_acc) atomicAdd(&counts[0], (unsigned long long)n_acc);
        if (n_new) atomicAdd(&counts[1], (unsigned long long)n_new);
        if (n_eval) atomicAdd(&counts[2], (unsigned long long)n_eval);
    }
}

}  // namespace

extern "C" int smcb_set_data_kinetic(smcb_handle* h, const double* cond_host, const double* obs_host, int n_cond,
                                     const double* base_host, int n_pairs, const int* est_pos_host, int d,
                                     int n_steps) {
    REQUIRE(h, h && cond_host && obs_host && base_host && est_pos_host, SMCB_ERR_INVALID, "null pointer");
    REQUIRE(h, n_cond >= 1 && n_steps >= 1, SMCB_ERR_INVALID, "need n_cond>=1, n_steps>=1");
    REQUIRE(h, n_pairs == 4 || n_pairs == 16, SMCB_ERR_UNSUPPORTED, "n_pairs must be 4 (reference) or 16");
    REQUIRE(h, d >= 1 && d <= SMCB_MAX_DIM && d <= 2 * n_pairs + 1, SMCB_ERR_INVALID, "bad d");
    const int full = 2 * n_pairs + 1;
    int inv[2 * kin::MAX_PAIRS + 1];
    for (int j = 0; j < full; ++j) inv[j] = -1;
    for (int k = 0; k < d; ++k) {
        REQUIRE(h, est_pos_host[k] >= 0 && est_pos_host[k] < full && inv[est_pos_host[k]] < 0, SMCB_ERR_INVALID,
                "est_pos entries must be distinct positions of the full parameter vector");
        inv[est_pos_host[k]] = k;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    KineticData& D = h->kin;
    if (D.cond) cudaFree(D.cond);
    if (D.obs) cudaFree(D.obs);
    if (D.base) cudaFree(D.base);
    if (D.est_pos) cudaFree(D.est_pos);
    D.cond = D.obs = D.base = nullptr;
    D.est_pos = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&D.cond, sizeof(double) * n_cond * SMCB_KIN_NCOND_FIELDS));
    CUDA_TRY(h, cudaMalloc((void**)&D.obs, sizeof(double) * 5 * n_cond));
    CUDA_TRY(h, cudaMalloc((void**)&D.base, sizeof(double) * full));
    CUDA_TRY(h, cudaMalloc((void**)&D.est_pos, sizeof(int) * full));
    CUDA_TRY(h, cudaMemcpy(D.cond, cond_host, sizeof(double) * n_cond * SMCB_KIN_NCOND_FIELDS,
                           cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(D.obs, obs_host, sizeof(double) * 5 * n_cond, cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(D.base, base_host, sizeof(double) * full, cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(D.est_pos, inv, sizeof(int) * full, cudaMemcpyHostToDevice));
    D.n_cond = n_cond;
    D.n_pairs = n_pairs;
    D.d = d;
    D.n_steps = n_steps;
    if (h->n_max > 0 && n_cond > h->ssr_rows) return smcb_reserve(h, h->n_max, h->d_max);
    return SMCB_OK;
}

int launch_loglik_kinetic(smcb_handle* h, const double* theta, int64_t ld, int64_t n, int d, const uint8_t* active,
                          double* lk, cudaStream_t st) {
    const KineticData& D = h->kin;
    REQUIRE(h, D.cond != nullptr, SMCB_ERR_STATE, "smcb_set_data_kinetic has not been called");
    REQUIRE(h, d == D.d, SMCB_ERR_INVALID, "d differs from the d given to smcb_set_data_kinetic");
    if (n == 0) return SMCB_OK;
    REQUIRE(h, h->ssr != nullptr && n <= h->n_max && D.n_cond <= h->ssr_rows, SMCB_ERR_STATE,
            "smcb_reserve too small for this sweep");
    const dim3 grid((unsigned)((n + KB - 1) / KB), (unsigned)D.n_cond);
    if (D.n_pairs == 4)
        kinetic_ssr_kernel<1><<<grid, KB, 0, st>>>(theta, ld, n, active, D.cond, D.obs, D.n_cond, D.n_steps, D.base,
                                                  D.est_pos, h->ssr);
    else
        kinetic_ssr_kernel<4><<<grid, KB, 0, st>>>(theta, ld, n, active, D.cond, D.obs, D.n_cond, D.n_steps, D.base,
                                                  D.est_pos, h->ssr);
    LAUNCH_CHECK(h);
    kinetic_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(theta, ld, n, active, h->ssr, D.n_cond,
                                                                       D.base, D.est_pos, 2 * D.n_pairs, lk);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_mh_fused(smcb_handle* h, int model, double* theta_dev, int64_t ld, double* lk_dev, int64_t n,
                             int d, const double* F_host, double ratio, const double* low_host,
                             const double* high_host, double gamma, int n_sweeps, uint64_t seed, uint64_t id_offset,
                             uint32_t stage, uint32_t sweep0, uint8_t* moved_dev, int64_t* counts_dev, void* stream) {
    REQUIRE(h, h && theta_dev && lk_dev && F_host && low_host && high_host && moved_dev && counts_dev,
            SMCB_ERR_INVALID, "null pointer");
    REQUIRE(h, n > 0 && d >= 1 && d <= SMCB_MAX_DIM && ld >= n && n_sweeps >= 1, SMCB_ERR_INVALID, "bad size");
    REQUIRE(h, model == SMCB_MODEL_KINETIC_RK, SMCB_ERR_UNSUPPORTED,
            "fused sweeps are implemented for KINETIC_RK only");
    const KineticData& D = h->kin;
    REQUIRE(h, D.cond != nullptr && d == D.d, SMCB_ERR_STATE, "kinetic data not set or d mismatch");
    FusedParams prm;
    memset(&prm, 0, sizeof(prm));
    memcpy(prm.F, F_host, sizeof(double) * d * d);
    memcpy(prm.low, low_host, sizeof(double) * d);
    memcpy(prm.high, high_host, sizeof(double) * d);
    const unsigned grid = (unsigned)((n + KB - 1) / KB);
    cudaStream_t st = as_stream(stream);
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(counts_dev);
    if (D.n_pairs == 4)
        kinetic_mh_fused_kernel<1><<<grid, KB, 0, st>>>(theta_dev, ld, lk_dev, n, d, prm, ratio, gamma, n_sweeps, seed,
                                                      id_offset, stage, sweep0, D.cond, D.obs, D.n_cond, D.n_steps,
                                                      D.base, D.est_pos, moved_dev, cnt);
    else
        kinetic_mh_fused_kernel<4><<<grid, KB, 0, st>>>(theta_dev, ld, lk_dev, n, d, prm, ratio, gamma, n_sweeps, seed,
                                                      id_offset, stage, sweep0, D.cond, D.obs, D.n_cond, D.n_steps,
                                                      D.base, D.est_pos, moved_dev, cnt);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}
