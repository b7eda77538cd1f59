// C-ABI entry points that are not kernels themselves: lifetime, scratch, data upload, dispatch.
#include <stdarg.h>
#include <stdlib.h>
#include <vector>

#include "common.cuh"

char g_create_err[512] = {0};

int smcb_fail(smcb_handle* h, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    char* dst = h ? h->err : g_create_err;
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

namespace {

template <typename T>
int dev_alloc(smcb_handle* h, T** p, size_t count) {
    if (*p) {
        cudaFree(*p);
        *p = nullptr;
    }
    if (count == 0) return SMCB_OK;
    CUDA_TRY(h, cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
    return SMCB_OK;
}

template <typename T>
void dev_free(T** p) {
    if (*p) cudaFree(*p);
    *p = nullptr;
}

// ---- FMA peak micro-benchmarks ---------------------------------------------------------------
__global__ void fma64_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void fma32_kernel(float* out, int iters) {
    float a0 = threadIdx.x * 1e-9f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
          a7 = a0 + 7;
    const float m = 1.0000001f, c = 1e-9f;
    for (int i = 0; i < iters; ++i) {
        a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
        a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace

extern "C" int smcb_version(void) { return SMCB_VERSION; }

extern "C" const char* smcb_last_error(const smcb_handle* h) { return h ? h->err : g_create_err; }

extern "C" int64_t smcb_launch_count(const smcb_handle* h) { return h ? h->launches : 0; }

extern "C" int smcb_create(int device, smcb_handle** out) {
    if (out == nullptr) return smcb_fail(nullptr, SMCB_ERR_INVALID, "smcb_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return smcb_fail(nullptr, SMCB_ERR_CUDA,
                         "smcb_create: no CUDA device (%s); this library has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count)
        return smcb_fail(nullptr, SMCB_ERR_INVALID, "smcb_create: device %d out of range [0,%d)", device, count);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return smcb_fail(nullptr, SMCB_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess)
        return smcb_fail(nullptr, SMCB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return smcb_fail(nullptr, SMCB_ERR_UNSUPPORTED,
                         "smcb_create: device is sm_%d%d; this library is built for sm_100a (B200) only",
                         prop.major, prop.minor);
    smcb_handle* h = new smcb_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    e = cudaMalloc(reinterpret_cast<void**>(&h->stats), SMCB_N_STATS * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(h->stats, 0, SMCB_N_STATS * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->seq_carry), 4 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->mm_ctl), 8 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(h->mm_ctl, 0, 8 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->mm_hist), 2 * 512 * sizeof(unsigned));
    if (e != cudaSuccess) {
        smcb_fail(nullptr, SMCB_ERR_CUDA, "smcb_create: cudaMalloc: %s", cudaGetErrorString(e));
        delete h;
        return SMCB_ERR_CUDA;
    }
    *out = h;
    return SMCB_OK;
}

extern "C" int smcb_destroy(smcb_handle* h) {
    if (!h) return SMCB_OK;
    cudaSetDevice(h->device);
    dev_free(&h->ssr); dev_free(&h->partial); dev_free(&h->stats); dev_free(&h->mm_ctl); dev_free(&h->mm_defer); dev_free(&h->mm_cutlim); dev_free(&h->mm_park);
    dev_free(&h->mm_bins); dev_free(&h->mm_perm); dev_free(&h->mm_hist); dev_free(&h->mm_tailrec);
    dev_free(&h->fused_plist); dev_free(&h->fused_owner);
    if (h->fused_ctl) cudaFree(h->fused_ctl);
    dev_free(&h->dae_dts);
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    h->prof_ev.clear();
    dev_free(&h->floor_cnt); dev_free(&h->resid_q); dev_free(&h->resid_f); dev_free(&h->tile_tot);
    dev_free(&h->tile_tot2); dev_free(&h->mark); dev_free(&h->seq_carry); dev_free(&h->rs_desc);
    dev_free(&h->mmp.t); dev_free(&h->mmp.P); dev_free(&h->mmp.S0);
    dev_free(&h->mmr.S); dev_free(&h->mmr.v); dev_free(&h->mmr.Sv32); dev_free(&h->mmr.suff);
    dev_free(&h->kin.cond); dev_free(&h->kin.obs); dev_free(&h->kin.base); dev_free(&h->kin.est_pos);
    smcb_comm_destroy(h);
    dev_free(&h->comm_send); dev_free(&h->comm_recv);
    delete h;
    return SMCB_OK;
}

extern "C" int smcb_reserve(smcb_handle* h, int64_t n_max, int d_max) {
    REQUIRE(h, h != nullptr, SMCB_ERR_INVALID, "null handle");
    REQUIRE(h, n_max > 0 && d_max >= 1 && d_max <= SMCB_MAX_DIM, SMCB_ERR_INVALID, "bad sizes");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rows = 1;
    if (h->mmp.n_ex > rows) rows = h->mmp.n_ex;
    if (h->kin.n_cond > rows) rows = h->kin.n_cond;
    // grow-only: scratch that is already large enough is kept (cudaMalloc/cudaFree of these buffers cost
    // more than a whole sampler run at 2^20 particles)
    if (n_max <= h->n_max && rows <= h->ssr_rows) {
        if (d_max > h->d_max) h->d_max = d_max;
        return SMCB_OK;
    }
    if (n_max < h->n_max) n_max = h->n_max;
    if (rows < h->ssr_rows) rows = (int)h->ssr_rows;
    if (d_max < h->d_max) d_max = h->d_max;
    int rc;
    if ((rc = dev_alloc(h, &h->ssr, (size_t)rows * n_max))) return rc;
    h->ssr_rows = rows;
    {   // worst case over the reductions that use it: [blocks][columns] partials followed by one row of results
        //   smcb_centered_moments, d > 6: 2*sm_count blocks x d(d+1)/2 pairs (528 at d = 32)
        //   smcb_temper_sums: 8*sm_count blocks x 2*SMCB_MAX_CAND sums;  smcb_colsum: 4*sm_count blocks x d
        const int64_t npair_max = (int64_t)SMCB_MAX_DIM * (SMCB_MAX_DIM + 1) / 2;
        int64_t need = ((int64_t)h->sm_count * 2 + 1) * npair_max;
        const int64_t t = ((int64_t)h->sm_count * 8 + 1) * 2 * SMCB_MAX_CAND;
        if (t > need) need = t;
        h->partial_len = need + 8192;
    }
    if ((rc = dev_alloc(h, &h->partial, (size_t)h->partial_len))) return rc;
    if ((rc = dev_alloc(h, &h->floor_cnt, (size_t)n_max))) return rc;
    if ((rc = dev_alloc(h, &h->resid_q, (size_t)n_max))) return rc;
    if ((rc = dev_alloc(h, &h->resid_f, (size_t)n_max))) return rc;
    if ((rc = dev_alloc(h, &h->mark, (size_t)n_max))) return rc;
    if ((rc = dev_alloc(h, &h->mm_defer, (size_t)(rows + 1) * n_max))) return rc;   // deferred solves + particles
    if ((rc = dev_alloc(h, &h->mm_cutlim, (size_t)n_max))) return rc;
    {   // parked solves: 48 bytes each; a sweep that hands over more than this restarts the excess from t0
        size_t cap = (size_t)(h->mmp.n_ex > 0 ? h->mmp.n_ex : 1) * n_max / 16;   // (progress-curve model only)
        if (cap < 65536) cap = 65536;
        if ((rc = dev_alloc(h, &h->mm_park, cap * 6))) return rc;
        h->mm_park_cap = (unsigned)cap;
    }
    if ((rc = dev_alloc(h, &h->mm_bins, (size_t)n_max))) return rc;
    if ((rc = dev_alloc(h, &h->mm_perm, (size_t)n_max))) return rc;
    {   // tail kernel: one 4-word record per thread of max(32 warps per SM, one lane per 32 solves) + the loop launch
        size_t lanes = (size_t)h->sm_count * 32 * 32;
        if ((size_t)rows * n_max / 32 > lanes) lanes = (size_t)rows * n_max / 32 + 32;
        lanes += (size_t)h->sm_count * 4 * 32;
        if ((rc = dev_alloc(h, &h->mm_tailrec, lanes * 4))) return rc;
    }
    const size_t tiles = (size_t)(n_max + 2047) / 2048 + 8;
    if ((rc = dev_alloc(h, &h->tile_tot, 2 * tiles))) return rc;
    if ((rc = dev_alloc(h, &h->tile_tot2, 2 * tiles))) return rc;
    if ((rc = dev_alloc(h, &h->rs_desc, 4 * tiles + 8))) return rc;
    h->n_max = n_max;
    h->d_max = d_max;
    return SMCB_OK;
}

extern "C" int smcb_set_data_mm_progress(smcb_handle* h, const double* t_host, const double* P_host,
                                         const double* S0_host, int n_ex, int n_t) {
    REQUIRE(h, h && t_host && P_host && S0_host, SMCB_ERR_INVALID, "null pointer");
    REQUIRE(h, n_ex >= 1 && n_t >= 2, SMCB_ERR_INVALID, "need n_ex>=1 and n_t>=2");
    for (int e = 0; e < n_ex; ++e)
        for (int i = 1; i < n_t; ++i)
            REQUIRE(h, t_host[e * n_t + i] > t_host[e * n_t + i - 1], SMCB_ERR_INVALID,
                    "t must be strictly increasing within an experiment");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc;
    const size_t m = (size_t)n_ex * n_t;
    if ((rc = dev_alloc(h, &h->mmp.t, m))) return rc;
    if ((rc = dev_alloc(h, &h->mmp.P, m))) return rc;
    if ((rc = dev_alloc(h, &h->mmp.S0, (size_t)n_ex))) return rc;
    CUDA_TRY(h, cudaMemcpy(h->mmp.t, t_host, m * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->mmp.P, P_host, m * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->mmp.S0, S0_host, n_ex * sizeof(double), cudaMemcpyHostToDevice));
    h->mmp.n_ex = n_ex;
    h->mmp.n_t = n_t;
    h->mm_bulk_blocks_per_sm = 0;   // shared-memory footprint changed: query the occupancy again
    h->mm_tail_blocks_per_sm = 0;
    h->mm_smem_set = false;
    if (h->n_max > 0 && n_ex > h->ssr_rows) return smcb_reserve(h, h->n_max, h->d_max);
    return SMCB_OK;
}

extern "C" int smcb_set_data_mm_rate(smcb_handle* h, const double* S_host, const double* v_host, int64_t n_obs,
                                     int precision) {
    REQUIRE(h, h && S_host && v_host && n_obs > 0, SMCB_ERR_INVALID, "bad argument");
    REQUIRE(h, precision == 32 || precision == 64, SMCB_ERR_INVALID, "precision must be 32 or 64");
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = dev_alloc(h, &h->mmr.S, (size_t)n_obs))) return rc;
    if ((rc = dev_alloc(h, &h->mmr.v, (size_t)n_obs))) return rc;
    if ((rc = dev_alloc(h, &h->mmr.Sv32, (size_t)n_obs))) return rc;
    CUDA_TRY(h, cudaMemcpy(h->mmr.S, S_host, n_obs * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->mmr.v, v_host, n_obs * sizeof(double), cudaMemcpyHostToDevice));
    std::vector<float2> packed((size_t)n_obs);
    for (int64_t i = 0; i < n_obs; ++i) packed[i] = make_float2((float)S_host[i], (float)v_host[i]);
    CUDA_TRY(h, cudaMemcpy(h->mmr.Sv32, packed.data(), n_obs * sizeof(float2), cudaMemcpyHostToDevice));
    h->mmr.n_obs = n_obs;
    h->mmr.precision = precision;
    return SMCB_OK;
}

// Sufficient-statistic form of the rate-law likelihood (SURVEY.md 8(d) "sanity of the headline target", H6).  The data
// enter  sum_i (v_i - Vmax S_i/(Km+S_i))^2 = sum v^2 - 2 Vmax A(Km) + Vmax^2 B(Km)  only through sum v^2 and the two
// one-dimensional functions A(Km) = sum_i v_i S_i/(Km+S_i), B(Km) = sum_i S_i^2/(Km+S_i)^2.  Both are analytic for
// Km > -min(S); on SMCB_SUFF_INT geometric intervals of u = Km + min(S) a degree-13 Chebyshev interpolant has its nearest
// pole 24 half-widths from the interval centre, i.e. a truncation error ~ (2*24)^-14 < 1e-23 relative: the tables carry
// A and B to the last bits of FP64 (they are summed in long double here), and a likelihood then costs ~60 flop instead
// of 7 flop per observation.
extern "C" int smcb_set_data_mm_rate_sufficient(smcb_handle* h, const double* S_host, const double* v_host, int64_t n_obs,
                                                double km_lo, double km_hi) {
    REQUIRE(h, h && S_host && v_host && n_obs > 0, SMCB_ERR_INVALID, "bad argument");
    REQUIRE(h, km_lo >= 0.0 && km_hi > km_lo && km_hi < 1e300, SMCB_ERR_INVALID, "need 0 <= km_lo < km_hi");
    double s0 = S_host[0];
    for (int64_t i = 0; i < n_obs; ++i) {
        REQUIRE(h, S_host[i] > 0.0, SMCB_ERR_INVALID, "substrate levels must be positive");
        if (S_host[i] < s0) s0 = S_host[i];
    }
    int rc = smcb_set_data_mm_rate(h, S_host, v_host, n_obs, 64);   // the direct FP64 sum serves Km outside the tables
    if (rc) return rc;
    const int NI = SMCB_SUFF_INT, M = SMCB_SUFF_M;
    const double u_lo = km_lo + s0, u_hi = km_hi + s0;
    const double log2rho = log2(u_hi / u_lo) / NI;
    std::vector<double> tab((size_t)NI * 2 * M + 2 * NI);
    std::vector<long double> fa(M), fb(M);
    long double sv2 = 0.0L;
    for (int64_t i = 0; i < n_obs; ++i) sv2 += (long double)v_host[i] * v_host[i];
    for (int j = 0; j < NI; ++j) {
        const double a = u_lo * exp2(log2rho * j), b = (j + 1 == NI) ? u_hi : u_lo * exp2(log2rho * (j + 1));
        const double ctr = 0.5 * (a + b), hw = 0.5 * (b - a);
        tab[(size_t)NI * 2 * M + j] = ctr;
        tab[(size_t)NI * 2 * M + NI + j] = 1.0 / hw;
        for (int k = 0; k < M; ++k) {
            const long double t = cosl(M_PIl * (k + 0.5L) / M);
            const long double km = (long double)ctr + (long double)hw * t - (long double)s0;
            long double A = 0.0L, B = 0.0L;
            for (int64_t i = 0; i < n_obs; ++i) {
                const long double g = (long double)S_host[i] / (km + (long double)S_host[i]);
                A += (long double)v_host[i] * g;
                B += g * g;
            }
            fa[k] = A;
            fb[k] = B;
        }
        for (int n = 0; n < M; ++n) {
            long double ca = 0.0L, cb = 0.0L;
            for (int k = 0; k < M; ++k) {
                const long double c = cosl(M_PIl * n * (k + 0.5L) / M);
                ca += fa[k] * c;
                cb += fb[k] * c;
            }
            const long double sc = (n == 0 ? 1.0L : 2.0L) / M;
            tab[((size_t)j * 2 + 0) * M + n] = (double)(ca * sc);
            tab[((size_t)j * 2 + 1) * M + n] = (double)(cb * sc);
        }
    }
    if ((rc = dev_alloc(h, &h->mmr.suff, tab.size()))) return rc;
    CUDA_TRY(h, cudaMemcpy(h->mmr.suff, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->mmr.sum_v2 = (double)sv2;
    h->mmr.suff_s0 = s0;
    h->mmr.suff_ulo = u_lo;
    h->mmr.suff_uhi = u_hi;
    h->mmr.suff_inv_log2rho = 1.0 / log2rho;
    h->mmr.precision = 0;
    return SMCB_OK;
}

extern "C" int smcb_loglik(smcb_handle* h, int model, const double* theta_dev, int64_t ld, int64_t n, int d,
                           const uint8_t* active_dev, double* lk_dev, void* stream) {
    return smcb_loglik_bounded(h, model, theta_dev, ld, n, d, active_dev, nullptr, lk_dev, stream);
}

extern "C" int smcb_loglik_bounded(smcb_handle* h, int model, const double* theta_dev, int64_t ld, int64_t n, int d,
                                   const uint8_t* active_dev, const double* lkmin_dev, double* lk_dev,
                                   void* stream) {
    REQUIRE(h, h && theta_dev && lk_dev, SMCB_ERR_INVALID, "null pointer");
    REQUIRE(h, n >= 0 && ld >= n, SMCB_ERR_INVALID, "need 0<=n<=ld");
    cudaStream_t st = as_stream(stream);
    switch (model) {
        case SMCB_MODEL_MM_PROGRESS:
            REQUIRE(h, d == 3, SMCB_ERR_INVALID, "MM_PROGRESS expects d=3 (Vmax, Km, sigma)");
            return launch_loglik_mm_progress(h, theta_dev, ld, n, active_dev, lkmin_dev, lk_dev, nullptr, st);
        case SMCB_MODEL_MM_RATE:
            REQUIRE(h, d == 3, SMCB_ERR_INVALID, "MM_RATE expects d=3 (Vmax, Km, sigma)");
            return launch_loglik_mm_rate(h, theta_dev, ld, n, active_dev, lk_dev, st);
        case SMCB_MODEL_KINETIC_RK:
            return launch_loglik_kinetic(h, theta_dev, ld, n, d, active_dev, lk_dev, st);
        case SMCB_MODEL_KINETIC_DAE:
            return launch_loglik_dae(h, theta_dev, ld, n, d, active_dev, lk_dev, st);
        case SMCB_MODEL_USER: {
            REQUIRE(h, h->user_fn != nullptr, SMCB_ERR_STATE, "smcb_set_user_likelihood has not been called");
            REQUIRE(h, d >= 1 && d <= SMCB_MAX_DIM, SMCB_ERR_INVALID, "bad d");
            if (n == 0) return SMCB_OK;
            // the callback enqueues the user's kernels on `stream`; lkmin is advisory and not passed on
            const int urc = h->user_fn(h->user_data, theta_dev, ld, n, d, active_dev, lk_dev, stream);
            if (urc != 0) return smcb_fail(h, SMCB_ERR_USER, "user likelihood returned %d", urc);
            h->launches++;
            const cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess)
                return smcb_fail(h, SMCB_ERR_CUDA, "user likelihood left a CUDA error: %s", cudaGetErrorString(e));
            return SMCB_OK;
        }
        default:
            return smcb_fail(h, SMCB_ERR_INVALID, "smcb_loglik: unknown model %d", model);
    }
}

extern "C" int smcb_zero(smcb_handle* h, void* ptr_dev, int64_t bytes, void* stream) {
    REQUIRE(h, h && ptr_dev && bytes >= 0, SMCB_ERR_INVALID, "bad argument");
    if (bytes > 0) CUDA_TRY(h, cudaMemsetAsync(ptr_dev, 0, (size_t)bytes, as_stream(stream)));
    return SMCB_OK;
}

extern "C" int smcb_copy_rows(smcb_handle* h, const double* src_dev, int64_t ld_src, double* dst_dev, int64_t ld_dst,
                              int64_t width, int rows, void* stream) {
    REQUIRE(h, h && src_dev && dst_dev && width >= 0 && rows >= 0 && ld_src >= width && ld_dst >= width,
            SMCB_ERR_INVALID, "bad argument");
    if (width > 0 && rows > 0)
        CUDA_TRY(h, cudaMemcpy2DAsync(dst_dev, sizeof(double) * (size_t)ld_dst, src_dev, sizeof(double) * (size_t)ld_src,
                                      sizeof(double) * (size_t)width, (size_t)rows, cudaMemcpyDeviceToDevice,
                                      as_stream(stream)));
    return SMCB_OK;
}

extern "C" int smcb_set_user_likelihood(smcb_handle* h, smcb_user_loglik_fn fn, void* user_data) {
    REQUIRE(h, h != nullptr, SMCB_ERR_INVALID, "null handle");
    h->user_fn = fn;
    h->user_data = user_data;
    return SMCB_OK;
}

extern "C" int smcb_set_param(smcb_handle* h, int key, double value) {
    REQUIRE(h, h != nullptr, SMCB_ERR_INVALID, "null handle");
    switch (key) {
        case SMCB_PARAM_MM_BUDGET:
            REQUIRE(h, value >= 1 && value <= 1e9, SMCB_ERR_INVALID, "MM_BUDGET must be in [1, 1e9]");
            h->mm_budget = (int)value;
            return SMCB_OK;
        case SMCB_PARAM_MM_REFILL_MIN:
            REQUIRE(h, value >= 1 && value <= 32, SMCB_ERR_INVALID, "MM_REFILL_MIN must be in [1, 32]");
            h->mm_refill_min = (int)value;
            return SMCB_OK;
        case SMCB_PARAM_MM_CHUNK:
            REQUIRE(h, value >= 1 && value <= 65536, SMCB_ERR_INVALID, "MM_CHUNK must be in [1, 65536]");
            h->mm_chunk = (int)value;
            return SMCB_OK;
        case SMCB_PARAM_MM_TAIL_WARPS:
            REQUIRE(h, value >= 1 && value <= 32, SMCB_ERR_INVALID, "MM_TAIL_WARPS must be in [1, 32]");
            h->mm_tail_warps = (int)value;
            return SMCB_OK;
        case SMCB_PARAM_PROFILE:
            // switching the recording on starts a new measurement; switching it off keeps what was recorded
            // until smcb_profile_read
            if (value != 0) {
                h->prof_sweeps = 0;
                h->prof_acc[0] = h->prof_acc[1] = 0.0;
                h->prof_acc_sweeps = 0;
            }
            h->prof_on = value != 0;
            return SMCB_OK;
        case SMCB_PARAM_MM_INTEGRATOR:
            REQUIRE(h, value == SMCB_MM_RK45_SCIPY || value == SMCB_MM_EXACT, SMCB_ERR_INVALID,
                    "MM_INTEGRATOR must be SMCB_MM_RK45_SCIPY or SMCB_MM_EXACT");
            h->mm_integrator = (int)value;
            return SMCB_OK;
        case SMCB_PARAM_MM_PATIENCE:
            REQUIRE(h, value >= 0 && value <= 1e6, SMCB_ERR_INVALID, "MM_PATIENCE must be in [0, 1e6]");
            h->mm_patience = (int)value;
            return SMCB_OK;
        default:
            return smcb_fail(h, SMCB_ERR_INVALID, "smcb_set_param: unknown key %d", key);
    }
}

extern "C" int smcb_predict_mm_progress(smcb_handle* h, const double* theta_dev, int64_t ld, int64_t n,
                                        double* pred_dev, void* stream) {
    REQUIRE(h, h && theta_dev && pred_dev && n > 0 && ld >= n, SMCB_ERR_INVALID, "bad argument");
    return launch_loglik_mm_progress(h, theta_dev, ld, n, nullptr, nullptr, nullptr, pred_dev, as_stream(stream));
}

extern "C" int smcb_loglik_stats(smcb_handle* h, int64_t* out_host) {
    REQUIRE(h, h && out_host, SMCB_ERR_INVALID, "null pointer");
    CUDA_TRY(h, cudaMemcpy(out_host, h->stats, SMCB_N_STATS * sizeof(int64_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(h, cudaMemset(h->stats + 16, 0, sizeof(unsigned long long)));   // [16] is "since the last read"
    return SMCB_OK;
}

extern "C" int smcb_profile_read(smcb_handle* h, double* out_host) {
    REQUIRE(h, h && out_host, SMCB_ERR_INVALID, "null pointer");
    out_host[0] = out_host[1] = out_host[2] = 0.0;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = prof_drain(h);
    if (rc) return rc;
    out_host[0] = h->prof_acc[0];
    out_host[1] = h->prof_acc[1];
    out_host[2] = (double)h->prof_acc_sweeps;
    h->prof_acc[0] = h->prof_acc[1] = 0.0;
    h->prof_acc_sweeps = 0;
    return SMCB_OK;
}

// Every recorded sweep is read exactly once: the times of the sweeps in the event list are added to the
// accumulators and the list starts over (its events are kept and reused).
int prof_drain(smcb_handle* h) {
    if (h->prof_sweeps == 0) return SMCB_OK;
    const size_t last = (size_t)(h->prof_sweeps - 1) * 4 + 3;
    REQUIRE(h, last < h->prof_ev.size(), SMCB_ERR_STATE, "profiling event list shorter than the recorded sweeps");
    CUDA_TRY(h, cudaEventSynchronize(h->prof_ev[last]));
    for (int k = 0; k < h->prof_sweeps; ++k) {
        float a = 0.f, b = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&a, h->prof_ev[(size_t)k * 4 + 0], h->prof_ev[(size_t)k * 4 + 1]));
        CUDA_TRY(h, cudaEventElapsedTime(&b, h->prof_ev[(size_t)k * 4 + 2], h->prof_ev[(size_t)k * 4 + 3]));
        h->prof_acc[0] += a;
        h->prof_acc[1] += b;
    }
    h->prof_acc_sweeps += h->prof_sweeps;
    h->prof_sweeps = 0;
    return SMCB_OK;
}

// Called at the start of a profiled sweep: the list grows SMCB_PROF_RING sweeps at a time (no synchronisation);
// once it holds SMCB_PROF_CAP sweeps it is drained, which costs one host sync every SMCB_PROF_CAP sweeps.
int prof_begin_sweep(smcb_handle* h) {
    if (!h->prof_on) return SMCB_OK;
    if (h->prof_sweeps >= SMCB_PROF_CAP) {
        int rc = prof_drain(h);
        if (rc) return rc;
    }
    const size_t need = (size_t)(h->prof_sweeps + 1) * 4;
    if (h->prof_ev.size() < need) {
        const size_t grown = h->prof_ev.size() + (size_t)SMCB_PROF_RING * 4;
        h->prof_ev.reserve(grown);
        while (h->prof_ev.size() < grown) {
            cudaEvent_t e;
            CUDA_TRY(h, cudaEventCreate(&e));
            h->prof_ev.push_back(e);
        }
    }
    return SMCB_OK;
}

extern "C" int smcb_measure_fma_peak(smcb_handle* h, double* out_host) {
    REQUIRE(h, h && out_host, SMCB_ERR_INVALID, "null pointer");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int threads = 256, blocks = h->sm_count * 8, iters = 1 << 15;
    double* buf = nullptr;
    CUDA_TRY(h, cudaMalloc(reinterpret_cast<void**>(&buf), (size_t)threads * blocks * sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0.f;
    for (int rep = 0; rep < 2; ++rep) {   // first repetition warms up
        cudaEventRecord(e0);
        fma64_kernel<<<blocks, threads>>>(buf, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    h->launches += 2;
    out_host[0] = 2.0 * 8 * (double)iters * threads * blocks / (ms * 1e-3);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        fma32_kernel<<<blocks, threads>>>(reinterpret_cast<float*>(buf), iters * 4);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    h->launches += 2;
    out_host[1] = 2.0 * 8 * (double)iters * 4 * threads * blocks / (ms * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    CUDA_TRY(h, cudaGetLastError());
    return SMCB_OK;
}
