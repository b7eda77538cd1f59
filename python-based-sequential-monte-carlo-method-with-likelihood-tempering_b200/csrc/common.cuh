// Shared declarations for libsmcb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <vector>

#include "../../include/smcb200.h"

#define SMCB_VERSION SMCB_ABI_VERSION
#define SMCB_N_STATS 24
#define SMCB_PROF_RING 512      // sweeps per growth step of the profiling event list
#define SMCB_PROF_CAP 16384     // sweeps after which the list is drained into the accumulators (one host sync)
#define SMCB_COMM_ROW 1280      // doubles per rank in the small all-gathers of the stage loop (d = 32 moments: 1093)
#define FULL_MASK 0xffffffffu

struct MmProgressData {
    double* t = nullptr;   // [n_ex][n_t]
    double* P = nullptr;   // [n_ex][n_t]
    double* S0 = nullptr;  // [n_ex]
    int n_ex = 0, n_t = 0;
};

struct MmRateData {
    double* S = nullptr;   // [n_obs] FP64
    double* v = nullptr;
    float2* Sv32 = nullptr;  // [n_obs] packed (S, v) FP32
    int64_t n_obs = 0;
    int precision = 64;    // 64 / 32: arithmetic of the direct sum; 0: sufficient-statistic form (tables below)
    double sum_v2 = 0.0;   // sum v^2 (host, FP64)
    // sufficient-statistic form (smcb_set_data_mm_rate_sufficient): Chebyshev tables of A(Km), B(Km)
    double* suff = nullptr;      // [SUFF_INT][2][SUFF_M] coefficients, then [SUFF_INT] centres, [SUFF_INT] 1/half-widths
    double suff_s0 = 0.0, suff_ulo = 0.0, suff_uhi = 0.0, suff_inv_log2rho = 0.0;
};
constexpr int SMCB_SUFF_INT = 64;   // geometric intervals in u = Km + min(S)
constexpr int SMCB_SUFF_M = 14;     // Chebyshev coefficients per interval (degree 13)

struct KineticData {
    double* cond = nullptr;   // [n_cond][SMCB_KIN_NCOND_FIELDS]
    double* obs = nullptr;    // [5][n_cond]
    double* base = nullptr;   // [n_pairs*2+1]
    int* est_pos = nullptr;   // [d]
    int n_cond = 0, n_pairs = 0, d = 0, n_steps = 0;
};

struct smcb_handle {
    int device = 0;
    int sm_count = 0;
    int64_t launches = 0;
    char err[512] = {0};
    // scratch (smcb_reserve)
    int64_t n_max = 0;
    int d_max = 0;
    double* ssr = nullptr;           // [n_ex][n_max]   per-(experiment,particle) residual sums
    int64_t ssr_rows = 0;
    double* partial = nullptr;       // block partials for reductions
    int64_t partial_len = 0;
    unsigned long long* stats = nullptr;  // SMCB_N_STATS counters (device)
    unsigned* mm_ctl = nullptr;      // [0] bulk queue head, [1] deferred solves, [2] deferred particles, [3] particles to evaluate, [4] parked solves
    unsigned* mm_defer = nullptr;    // [ssr_rows*n_max] deferred solves, then [n_max] their particles
    double* mm_cutlim = nullptr;     // [n_max] per-particle residual limit of a bounded sweep
    double* mm_park = nullptr;       // [mm_park_cap][6] state of the solves the bulk kernel handed over (the tail kernel resumes them)
    unsigned mm_park_cap = 0;
    unsigned short* mm_bins = nullptr;   // [n_max] cost bin of every particle (0xFFFF = no solve needed)
    unsigned* mm_perm = nullptr;     // [n_max] particles to evaluate, heaviest cost bin first
    unsigned long long* mm_tailrec = nullptr;   // [lanes of the tail launches][4] per-thread work record of the tail kernel (sized in smcb_reserve)
    unsigned* mm_hist = nullptr;     // [2*512] histogram and scatter cursors of the counting sort
    double* fused_plist = nullptr;   // smcb_mh_fused: surviving proposals of a sweep [d][n_max]
    unsigned* fused_owner = nullptr; // ... their particles
    void* fused_ctl = nullptr;       // ... per-sweep counts of surviving proposals
    int64_t fused_cap = 0;           // ... particle count (n_max) the two lists were allocated for
    double* dae_dts = nullptr;       // transient reactor model: time steps of the fixed grid
    int dae_n_dt = 0;
    bool prof_on = false;            // per-kernel CUDA-event timing of the MM_PROGRESS sweeps
    int prof_sweeps = 0;             // sweeps whose events sit in prof_ev (not yet folded into prof_acc)
    std::vector<cudaEvent_t> prof_ev;    // 4 events per sweep (bulk start/end, tail start/end); grows on demand
    double prof_acc[2] = {0.0, 0.0}; // bulk / tail milliseconds of the sweeps already drained
    long long prof_acc_sweeps = 0;
    int mm_integrator = 0;           // SMCB_MM_RK45_SCIPY (reference parity) | SMCB_MM_EXACT (closed form)
    int mm_budget = 512;             // attempted steps after which the bulk kernel defers a solve
    int mm_tail_warps = 32;          // one-warp blocks per SM of the tail kernel
    int mm_chunk = 32;               // particles per queue item of the bulk kernel
    int mm_patience = 3;             // ... for this many attempted steps (unless the whole warp is free)
    int mm_refill_min = 8;           // free lanes a warp of the bulk kernel waits for before setting up new solves
    int mm_bulk_blocks_per_sm = 0;   // occupancy of the bulk kernel (queried once)
    int mm_tail_blocks_per_sm = 0;   // occupancy of the tail kernel (queried once)
    bool mm_smem_set = false;
    int32_t* floor_cnt = nullptr;    // [n_max]
    uint64_t* resid_q = nullptr;     // [n_max] fixed-point residuals
    double* resid_f = nullptr;       // [n_max] FP64 residuals (sequential mode)
    int64_t* tile_tot = nullptr;     // [2*tiles] tile totals (floor, q)
    int64_t* tile_tot2 = nullptr;    // second tile array (ancestor expansion)
    int32_t* mark = nullptr;         // [n_max] head marks for the ancestor fill
    double* seq_carry = nullptr;     // [4] sequential-scan carry (device)
    unsigned long long* rs_desc = nullptr;   // single-pass resampling: [tile counter | aggregates 2*tiles | inclusive prefixes 2*tiles]
    MmProgressData mmp;
    MmRateData mmr;
    KineticData kin;
    // communicator (comm.cu): NCCL, one process per GPU; world == 1 without smcb_comm_init
    void* comm = nullptr;            // ncclComm_t
    int rank = 0, world = 1;
    int64_t collectives = 0;         // NCCL operations enqueued so far
    double* comm_send = nullptr;     // [SMCB_COMM_ROW] this rank's row of a small all-gather
    double* comm_recv = nullptr;     // [world][SMCB_COMM_ROW]
    int comm_recv_rows = 0;
    // user-supplied likelihood (SMCB_MODEL_USER): host callback that enqueues the user's kernels
    smcb_user_loglik_fn user_fn = nullptr;
    void* user_data = nullptr;
};

extern char g_create_err[512];

int smcb_fail(smcb_handle* h, int code, const char* fmt, ...);

#define CUDA_TRY(h, expr)                                                                  \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return smcb_fail((h), SMCB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,           \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                  \
    } while (0)

#define LAUNCH_CHECK(h)                                                                    \
    do {                                                                                   \
        (h)->launches++;                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess)                                                             \
            return smcb_fail((h), SMCB_ERR_CUDA, "kernel launch failed: %s (%s:%d)",       \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                  \
    } while (0)

#define REQUIRE(h, cond, code, msg)                                                        \
    do {                                                                                   \
        if (!(cond)) return smcb_fail((h), (code), "%s: %s", __func__, (msg));             \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------ warp / block reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// Block-wide sum of K doubles per thread; result valid in thread 0.  smem: K*32 doubles.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* smem) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) smem[k * 32 + wid] = v[k];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double x = (lane < nw) ? smem[k * 32 + lane] : 0.0;
            v[k] = warp_sum(x);
        }
    }
    __syncthreads();
}

// SMCB_PARAM_PROFILE bookkeeping (api.cu): make room for one more sweep's events / fold recorded sweeps into prof_acc
int prof_begin_sweep(smcb_handle* h);
int prof_drain(smcb_handle* h);

// comm.cu: recv_dev[r*count + i] = send_dev[i] of rank r (a device copy when world == 1)
int comm_all_gather_f64(smcb_handle* h, const double* send_dev, double* recv_dev, int64_t count, cudaStream_t st);
// small staging row / table for the all-gathers of the stage loop (allocated on first use, also without NCCL)
int comm_staging(smcb_handle* h);

// kernels implemented across translation units
int launch_loglik_mm_progress(smcb_handle* h, const double* theta, int64_t ld, int64_t n,
                              const uint8_t* active, const double* lkmin, double* lk, double* pred,
                              cudaStream_t st);
int launch_loglik_mm_rate(smcb_handle* h, const double* theta, int64_t ld, int64_t n,
                          const uint8_t* active, double* lk, cudaStream_t st);
int launch_loglik_kinetic(smcb_handle* h, const double* theta, int64_t ld, int64_t n, int d,
                          const uint8_t* active, double* lk, cudaStream_t st);
int launch_loglik_dae(smcb_handle* h, const double* theta, int64_t ld, int64_t n, int d,
                      const uint8_t* active, double* lk, cudaStream_t st);
int kinetic_pack_active(smcb_handle* h, const uint8_t* active, int64_t n, cudaStream_t st, unsigned** list,
                        unsigned** count);
int kinetic_finalize(smcb_handle* h, const double* theta, int64_t ld, int64_t n, const uint8_t* active, double* lk,
                     cudaStream_t st);
