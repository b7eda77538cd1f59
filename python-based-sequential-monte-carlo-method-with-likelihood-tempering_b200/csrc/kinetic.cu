// K1': kinetic reactor log-likelihood (one thread per (particle, condition); masked sweeps are packed into full
// warps first) and the multi-sweep Metropolis-Hastings call (propose / march / accept kernels per sweep, no host
// round trip in between).
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
    K.set_limit();
    const int is = inv_pos[8 * M];
    *sigma = (is >= 0) ? theta[(int64_t)is * ld + p] : base[8 * M];
}

// Work list of a masked sweep (MH proposals that passed the prior box): the RK march costs the same for every
// particle, so a warp with one active lane costs as much as a full one.  The list packs the active particles into
// full warps; its order is whatever the warp-aggregated appends give, which no result depends on.
__global__ void kinetic_compact_kernel(const uint8_t* __restrict__ active, int64_t n, unsigned* __restrict__ list,
                                       unsigned* __restrict__ count) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = p < n && active[p];
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (m == 0) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (on) list[base + __popc(m & ((1u << lane) - 1))] = (unsigned)p;
}

// One thread per (particle, condition).  Three ways of naming the particles of a sweep: all n of them (list and
// count null), the entries of a work list (masked sweep), or the first *count columns of theta (fused sweeps: theta
// is then the packed list of surviving proposals).  Grid-stride in x so that a launch sized before the count is
// known does not pay for empty blocks.
template <int M, bool PACKED>
__global__ void __launch_bounds__(KB, M == 1 ? (PACKED ? 4 : 5) : 3)
kinetic_ssr_kernel(const double* __restrict__ theta, int64_t ld, int64_t n, const unsigned* __restrict__ list,
                   const unsigned* __restrict__ count, const double* __restrict__ cond,
                   const double* __restrict__ obs, int n_cond, int n_steps, const double* __restrict__ base,
                   const int* __restrict__ inv_pos, double* __restrict__ ssr) {
    const int c = blockIdx.y;
    __shared__ double etab[expt::TAB_N];
    expt::load_table(etab);
    __syncthreads();
    if (!PACKED) {
        const int64_t p = (int64_t)blockIdx.x * KB + threadIdx.x;
        if (p >= n) return;
        kin::Kin<M> K;
        double sigma;
        load_kin<M>(theta, ld, p, base, inv_pos, K, &sigma);
        ssr[(int64_t)c * n + p] =
            kin::condition_ssr<M>(K, cond + (int64_t)c * SMCB_KIN_NCOND_FIELDS, n_steps, obs, n_cond, c, etab);
    } else {
        const unsigned m = *count, stride = gridDim.x * KB;
#pragma unroll 1
        for (unsigned q = blockIdx.x * KB + threadIdx.x; q < m; q += stride) {
            const int64_t p = (list != nullptr) ? (int64_t)list[q] : (int64_t)q;
            kin::Kin<M> K;
            double sigma;
            load_kin<M>(theta, ld, p, base, inv_pos, K, &sigma);
            ssr[(int64_t)c * n + p] =
                kin::condition_ssr<M>(K, cond + (int64_t)c * SMCB_KIN_NCOND_FIELDS, n_steps, obs, n_cond, c, etab);
        }
    }
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

// Several sweeps per call, no host round trip in between.  In a d-dimensional box few proposals survive the prior
// test (0.5% of a 32-parameter prior cloud per sweep) and every survivor costs n_cond RK marches of identical
// length, so a sweep is three launches that pool the survivors of the whole shard:
//   propose  every particle: Philox normals, factor mat-vec (factor in shared memory, proposal in registers), box
//            test; survivors are appended to a packed SoA list
//   march    kinetic_ssr_kernel over the list: full warps of (proposal, condition) marches
//   accept   the owner adds its conditions up in the order of kinetic_finalize_kernel and accepts or rejects
constexpr int FUSED_MAX_SWEEPS = 64;

template <int DMAX>
__global__ void __launch_bounds__(KB)
fused_propose_kernel(const double* __restrict__ theta, int64_t ld, int64_t n, int d,
                     const __grid_constant__ FusedParams prm, double ratio, uint64_t seed, uint64_t id_offset,
                     uint32_t stage, uint32_t sweep, uint8_t* __restrict__ moved,
                     unsigned long long* __restrict__ counts, unsigned* __restrict__ count,
                     double* __restrict__ plist, unsigned* __restrict__ owner) {
    constexpr int DS = (DMAX + 1) & ~1;   // even row stride: 16-byte aligned rows
    __shared__ __align__(16) double sF[DMAX * DS];
    for (int e = threadIdx.x; e < d * d; e += KB) sF[(e / d) * DS + (e % d)] = prm.F[e];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * KB + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool ok = false;
    long long n_acc = 0, n_new = 0;
    double pr[DMAX];   // compile-time loop bounds below keep the proposal in registers
    if (i < n) {
        // the particle is only needed after the mat-vec: start its rows on their way now, without holding registers
#pragma unroll
        for (int k = 0; k < DMAX; ++k)
            if (k < d) asm volatile("prefetch.global.L2 [%0];" ::"l"(theta + (int64_t)k * ld + i));
#pragma unroll
        for (int k = 0; k < DMAX; ++k) pr[k] = 0.0;
#pragma unroll
        for (int j = 0; j < DMAX; j += 2) {
            if (j < d) {
                double z0, z1;
                philox_normal2(seed, id_offset + (uint64_t)i, stage, sweep, (uint32_t)(j >> 1), &z0, &z1);
                if (j + 1 >= d) z1 = 0.0;
                const double* f0 = sF + j * DS;
                const double* f1 = sF + (j + 1 < DMAX ? j + 1 : j) * DS;
#pragma unroll
                for (int k = 0; k < DMAX; ++k)
                    if (k < d) pr[k] = fma(z1, f1[k], fma(z0, f0[k], pr[k]));
            }
        }
        ok = true;
#pragma unroll
        for (int k = 0; k < DMAX; ++k)
            if (k < d) {
                pr[k] = theta[(int64_t)k * ld + i] + pr[k] * ratio;
                ok = ok && pr[k] >= prm.low[k] && pr[k] <= prm.high[k];
            }
        // outside the box the reference's acceptance probability is 0, and 0 >= u holds for u == 0
        if (!ok && philox_uniform(seed, id_offset + (uint64_t)i, stage, sweep, SMCB_SLOT_UNIFORM) <= 0.0) {
            n_acc = 1;
            if (!moved[i]) {
                moved[i] = 1;
                n_new = 1;
            }
        }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, ok);
    if (mask) {
        const int leader = __ffs(mask) - 1;
        unsigned slot0 = 0;
        if (lane == leader) slot0 = atomicAdd(count, (unsigned)__popc(mask));
        slot0 = __shfl_sync(0xffffffffu, slot0, leader);
        if (ok) {
            const unsigned q = slot0 + __popc(mask & ((1u << lane) - 1));
            owner[q] = (unsigned)i;
#pragma unroll
            for (int k = 0; k < DMAX; ++k)
                if (k < d) plist[(int64_t)k * n + q] = pr[k];
        }
    }
    if (__any_sync(0xffffffffu, n_acc != 0)) {   // u == 0: once in 2^53 draws
        n_acc = warp_sum_ll(n_acc);
        n_new = warp_sum_ll(n_new);
        if (lane == 0) {
            atomicAdd(&counts[0], (unsigned long long)n_acc);
            if (n_new) atomicAdd(&counts[1], (unsigned long long)n_new);
        }
    }
}

__global__ void fused_accept_kernel(double* __restrict__ theta, int64_t ld, double* __restrict__ lk, int64_t n, int d,
                                    double gamma, uint64_t seed, uint64_t id_offset, uint32_t stage, uint32_t sweep,
                                    int n_cond, const double* __restrict__ base, const int* __restrict__ inv_pos,
                                    int sigma_pos, uint8_t* __restrict__ moved,
                                    unsigned long long* __restrict__ counts, const unsigned* __restrict__ count,
                                    const double* __restrict__ plist, const unsigned* __restrict__ owner,
                                    const double* __restrict__ ssr) {
    const int64_t m = *count;
    long long n_acc = 0, n_new = 0;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = owner[q];
        const int is = inv_pos[sigma_pos];
        const double sigma = (is >= 0) ? plist[(int64_t)is * n + q] : base[sigma_pos];
        double tot = 0.0;
        for (int c = 0; c < n_cond; ++c) tot += ssr[(int64_t)c * n + q];
        const double l2 = -(0.5 / (sigma * sigma)) * tot - 5.0 * n_cond * log(sigma);
        const double u = philox_uniform(seed, id_offset + (uint64_t)i, stage, sweep, SMCB_SLOT_UNIFORM);
        if (exp((l2 - lk[i]) * gamma) >= u) {
            ++n_acc;
            for (int k = 0; k < d; ++k) theta[(int64_t)k * ld + i] = plist[(int64_t)k * n + q];
            lk[i] = l2;
            if (!moved[i]) {
                moved[i] = 1;
                ++n_new;
            }
        }
    }
    n_acc = warp_sum_ll(n_acc);
    n_new = warp_sum_ll(n_new);
    if ((threadIdx.x & 31) == 0) {
        if (n_acc) atomicAdd(&counts[0], (unsigned long long)n_acc);
        if (n_new) atomicAdd(&counts[1], (unsigned long long)n_new);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && m) atomicAdd(&counts[2], (unsigned long long)m);
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

// work list of a masked sweep (null pointers when there is no mask)
int kinetic_pack_active(smcb_handle* h, const uint8_t* active, int64_t n, cudaStream_t st, unsigned** list,
                        unsigned** count) {
    *list = nullptr;
    *count = nullptr;
    if (active == nullptr) return SMCB_OK;
    *list = h->mm_perm;
    *count = reinterpret_cast<unsigned*>(h->mm_ctl);
    CUDA_TRY(h, cudaMemsetAsync(*count, 0, sizeof(unsigned), st));
    kinetic_compact_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(active, n, *list, *count);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

// lk from the per-condition residual sums in h->ssr
int kinetic_finalize(smcb_handle* h, const double* theta, int64_t ld, int64_t n, const uint8_t* active, double* lk,
                     cudaStream_t st) {
    const KineticData& D = h->kin;
    kinetic_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(theta, ld, n, active, h->ssr, D.n_cond,
                                                                       D.base, D.est_pos, 2 * D.n_pairs, lk);
    LAUNCH_CHECK(h);
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
    dim3 grid((unsigned)((n + KB - 1) / KB), (unsigned)D.n_cond);
    unsigned* list = nullptr;
    unsigned* count = nullptr;
    const int rc = kinetic_pack_active(h, active, n, st, &list, &count);
    if (rc != SMCB_OK) return rc;
    if (active != nullptr && grid.x > (unsigned)h->sm_count * 4) grid.x = (unsigned)h->sm_count * 4;
#define SSR_ARGS theta, ld, n, list, count, D.cond, D.obs, D.n_cond, D.n_steps, D.base, D.est_pos, h->ssr
    if (D.n_pairs == 4 && !active) kinetic_ssr_kernel<1, false><<<grid, KB, 0, st>>>(SSR_ARGS);
    else if (D.n_pairs == 4) kinetic_ssr_kernel<1, true><<<grid, KB, 0, st>>>(SSR_ARGS);
    else if (!active) kinetic_ssr_kernel<4, false><<<grid, KB, 0, st>>>(SSR_ARGS);
    else kinetic_ssr_kernel<4, true><<<grid, KB, 0, st>>>(SSR_ARGS);
#undef SSR_ARGS
    LAUNCH_CHECK(h);
    return kinetic_finalize(h, theta, ld, n, active, lk, st);
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
    REQUIRE(h, n_sweeps <= FUSED_MAX_SWEEPS, SMCB_ERR_UNSUPPORTED, "at most 64 sweeps per launch");
    REQUIRE(h, h->ssr != nullptr && n <= h->n_max && D.n_cond <= h->ssr_rows, SMCB_ERR_STATE,
            "smcb_reserve too small for this sweep");
    CUDA_TRY(h, cudaSetDevice(h->device));
    // grow-only list of surviving proposals [d][n] and their owners [n], both sized for the handle's current n_max
    // (fused_cap = the n_max they were allocated for: smcb_reserve may have grown n_max since)
    if (h->fused_plist == nullptr || h->fused_cap < h->n_max) {
        if (h->fused_plist) cudaFree(h->fused_plist);
        if (h->fused_owner) cudaFree(h->fused_owner);
        h->fused_plist = nullptr;
        h->fused_owner = nullptr;
        h->fused_cap = 0;
        CUDA_TRY(h, cudaMalloc((void**)&h->fused_plist, sizeof(double) * h->n_max * SMCB_MAX_DIM));
        CUDA_TRY(h, cudaMalloc((void**)&h->fused_owner, sizeof(unsigned) * h->n_max));
        h->fused_cap = h->n_max;
    }
    if (!h->fused_ctl) {
        CUDA_TRY(h, cudaMalloc((void**)&h->fused_ctl, sizeof(unsigned) * FUSED_MAX_SWEEPS));
    }
    cudaStream_t st = as_stream(stream);
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(counts_dev);
    unsigned* count = reinterpret_cast<unsigned*>(h->fused_ctl);
    CUDA_TRY(h, cudaMemsetAsync(count, 0, sizeof(unsigned) * FUSED_MAX_SWEEPS, st));
    const unsigned blocks = (unsigned)((n + KB - 1) / KB);
    dim3 mgrid(blocks, (unsigned)D.n_cond);
    if (mgrid.x > (unsigned)h->sm_count * 4) mgrid.x = (unsigned)h->sm_count * 4;
    unsigned agrid = (unsigned)((n + 255) / 256);
    if (agrid > (unsigned)h->sm_count * 8) agrid = (unsigned)h->sm_count * 8;
    for (int s = 0; s < n_sweeps; ++s) {
        const uint32_t sweep = sweep0 + (uint32_t)s;
        if (d <= 9)
            fused_propose_kernel<9><<<blocks, KB, 0, st>>>(theta_dev, ld, n, d, prm, ratio, seed, id_offset, stage, sweep,
                                                          moved_dev, cnt, count + s, h->fused_plist, h->fused_owner);
        else
            fused_propose_kernel<SMCB_MAX_DIM><<<blocks, KB, 0, st>>>(theta_dev, ld, n, d, prm, ratio, seed, id_offset,
                                                                     stage, sweep, moved_dev, cnt, count + s,
                                                                     h->fused_plist, h->fused_owner);
        LAUNCH_CHECK(h);
        if (D.n_pairs == 4)
            kinetic_ssr_kernel<1, true><<<mgrid, KB, 0, st>>>(h->fused_plist, n, n, nullptr, count + s, D.cond, D.obs,
                                                       D.n_cond, D.n_steps, D.base, D.est_pos, h->ssr);
        else
            kinetic_ssr_kernel<4, true><<<mgrid, KB, 0, st>>>(h->fused_plist, n, n, nullptr, count + s, D.cond, D.obs,
                                                       D.n_cond, D.n_steps, D.base, D.est_pos, h->ssr);
        LAUNCH_CHECK(h);
        fused_accept_kernel<<<agrid, 256, 0, st>>>(theta_dev, ld, lk_dev, n, d, gamma, seed, id_offset, stage, sweep,
                                                   D.n_cond, D.base, D.est_pos, 2 * D.n_pairs, moved_dev, cnt,
                                                   count + s, h->fused_plist, h->fused_owner, h->ssr);
        LAUNCH_CHECK(h);
    }
    return SMCB_OK;
}
