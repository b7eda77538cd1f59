/*
 * smcb200.h - C-ABI of the B200-native likelihood-tempered SMC hot path.
 *
 * One shared library (libsmcb200.so, hand-written sm_100a CUDA) replaces the
 * body of the reference sampler loop.  The reference has no FFI; its de-facto
 * interface is a handful of Python call sites, cited per entry point below
 * (paths relative to the upstream repository):
 *
 *   EX/main = SMC_example/Micmem_SMC_main.py      EX/lik = SMC_example/Micmem_likelihood.py
 *   EX/set  = SMC_example/Micmem_settings.py      ME/lik = SMC_methanation/methanation_set_likelihood.py
 *   ME/fun  = SMC_methanation/methanation_functions.py
 *
 * Conventions
 *   - Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - Pointers named *_dev are device pointers owned by the caller (the Python
 *     host keeps them alive as torch tensors); *_host are host pointers.
 *   - Particle state is SoA: theta_dev[k*ld + i] is parameter k of local
 *     particle i (ld >= n, the leading dimension in elements).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Every call is asynchronous on that stream unless it says "synchronous".
 *   - Every function returns 0 on success or a negative SMCB_ERR_* code; the
 *     message is available from smcb_last_error().  A non-finite likelihood is
 *     a value (-inf), never an error.
 *   - The library allocates device scratch only in smcb_create / smcb_reserve /
 *     smcb_set_data_*; nothing is allocated on the hot path.
 *   - There is no CPU fallback: without a CUDA device smcb_create fails.
 */
#ifndef SMCB200_H
#define SMCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMCB_ABI_VERSION 200 /* returned by smcb_version(); the ctypes layer refuses a library that differs */

#define SMCB_OK 0
#define SMCB_ERR_INVALID (-1)     /* bad argument                               */
#define SMCB_ERR_CUDA (-2)        /* a CUDA runtime call failed                 */
#define SMCB_ERR_STATE (-3)       /* call order (e.g. data not set, no reserve) */
#define SMCB_ERR_UNSUPPORTED (-4) /* size / option outside what is compiled in  */
#define SMCB_ERR_COMM (-5)        /* an NCCL call failed                        */
#define SMCB_ERR_USER (-6)        /* the user-supplied likelihood reported an error */

/* likelihood models */
#define SMCB_MODEL_MM_PROGRESS 1 /* EX/lik:35-77: six progress curves, scipy-RK45 twin */
#define SMCB_MODEL_MM_RATE 2     /* synthetic rate-law observations (S_i, v_i)         */
#define SMCB_MODEL_KINETIC_RK 3  /* methanation-style plug-flow reactor, fixed-step RK4 */
#define SMCB_MODEL_KINETIC_DAE 4 /* ME/lik:69-277: the reference's transient 357-unknown reactor DAE, implicit Euler to
                                    75 s (data as for KINETIC_RK with the reference's 8 kinetic parameters; n_steps unused) */
#define SMCB_MODEL_USER 5        /* EX/lik:79-92, ME/fun:70-92: the reference's plug-in is a user-written sim_particle;
                                    here a user callback that enqueues its own kernels (smcb_set_user_likelihood) */

/* resampling prefix-sum arithmetic */
#define SMCB_SCAN_SEQUENTIAL 0 /* the reference's sequentially rounded FP64 sum (EX/main:165-174), bit-exact */
#define SMCB_SCAN_FIXED 1      /* exact 2^-62 fixed-point parallel scan, independent of sharding */

#define SMCB_MAX_DIM 32        /* max estimated parameters d                  */
#define SMCB_MAX_CAND 16       /* max tempering candidates per pass           */

typedef struct smcb_handle smcb_handle;

/* A user-supplied likelihood (SMCB_MODEL_USER).  The reference's whole plug-in surface is a user-written
 * `sim_particle(particle) -> llk` (EX/lik:79-92, ME/fun:70-92); its counterpart here is a host function that
 * ENQUEUES the user's own kernels on `stream` (it must not synchronise): for i < n with active_dev == NULL or
 * active_dev[i] != 0 it writes lk_dev[i] = log-likelihood of (theta_dev[0*ld+i], ..., theta_dev[(d-1)*ld+i]); other
 * entries of lk_dev are left untouched.  Returns 0, or non-zero to make the calling smcb_* entry fail with
 * SMCB_ERR_USER.  `smcb_user.cuh` (next to this header) has the few lines a user kernel needs. */
typedef int (*smcb_user_loglik_fn)(void* user_data, const double* theta_dev, int64_t ld, int64_t n, int d,
                                   const uint8_t* active_dev, double* lk_dev, void* stream);

/* ---- lifetime ------------------------------------------------------------------------- */
int smcb_version(void);
/* Replaces ray.init (EX/main:56): binds the handle to one CUDA device. */
int smcb_create(int device, smcb_handle** out);
int smcb_destroy(smcb_handle* h);
const char* smcb_last_error(const smcb_handle* h); /* h may be NULL: last create error */
/* Size the scratch for at most n_max local particles / slots and d_max parameters. */
int smcb_reserve(smcb_handle* h, int64_t n_max, int d_max);
/* Number of kernels this handle has launched so far (for bench.py's gpu_launches). */
int64_t smcb_launch_count(const smcb_handle* h);

/* ---- data (replaces the module globals `dataset`, `n_ex`, `datapoint`, EX/set:103-115) - */
/* t, P: [n_ex][n_t] row-major; S0: [n_ex].  Synchronous (copies to the device). */
int smcb_set_data_mm_progress(smcb_handle* h, const double* t_host, const double* P_host,
                              const double* S0_host, int n_ex, int n_t);
/* S, v: [n_obs].  precision: 64 = FP64 arithmetic, 32 = FP32 arithmetic with FP32 accumulation
 * in 64-observation tiles and FP64 across tiles. */
int smcb_set_data_mm_rate(smcb_handle* h, const double* S_host, const double* v_host,
                          int64_t n_obs, int precision);
/* The same likelihood in its sufficient-statistic form: the residual sum of squares is
 * sum v^2 - 2 Vmax A(Km) + Vmax^2 B(Km) with A = sum v_i S_i/(Km+S_i), B = sum S_i^2/(Km+S_i)^2; A and B are tabulated
 * once (piecewise Chebyshev in Km over [km_lo, km_hi], summed in long double on the host, accurate to the last bits
 * of FP64) and a likelihood then costs ~60 flop whatever n_obs is.  Agreement with the direct FP64 sum: better than
 * 1e-9 relative on the log-likelihood (tests/test_gpu_kernels.py; the cancellation in the three-term form is what
 * limits it, not the tables).  Particles with Km outside [km_lo, km_hi] take the direct FP64 sum.  Synchronous. */
int smcb_set_data_mm_rate_sufficient(smcb_handle* h, const double* S_host, const double* v_host, int64_t n_obs,
                                     double km_lo, double km_hi);
/* Methanation-style reactor (ME/lik:44-66,204-208,289-298; reactor definition in DESIGN.md).
 * cond: [n_cond][SMCB_KIN_NCOND_FIELDS] row-major operating conditions
 *       (Ca,Cb,Cc,Cd,Ce inlet [mol/m3], T_in [K], T_jacket [K], u_in [m/s], void, length [m]);
 * obs:  [5][n_cond] outlet flows [sccm] (the layout of ME's `data.csv`);
 * base: [n_pairs*2+1] full parameter vector (A_j,E_j pairs then sigma) used for positions that
 *       are not estimated; est_pos: [d] positions of the estimated parameters in that vector
 *       (ME/fun:80 `p_pred_bases[:, est_position] = particle`). */
#define SMCB_KIN_NCOND_FIELDS 10
int smcb_set_data_kinetic(smcb_handle* h, const double* cond_host, const double* obs_host, int n_cond,
                          const double* base_host, int n_pairs, const int* est_pos_host, int d,
                          int n_steps);

/* ---- K1: per-particle log-likelihood (replaces sim_particle, EX/lik:79-92, ME/fun:70-92) */
/* lk_dev[i] = log-likelihood of particle i for i<n.  active_dev (may be NULL) is a byte mask:
 * particles with active==0 are skipped and lk_dev[i] is left untouched. */
int smcb_loglik(smcb_handle* h, int model, const double* theta_dev, int64_t ld, int64_t n, int d,
                const uint8_t* active_dev, double* lk_dev, void* stream);
/* The same sweep with early rejection: lkmin_dev[i] (may be NULL = smcb_loglik) is a value below which
 * the caller does not need lk_dev[i] exactly (smcb_mh_threshold).  A particle whose log-likelihood is
 * PROVEN to lie below lkmin_dev[i] - from an upper bound that only uses the residuals accumulated so
 * far - may stop early and report -inf; every other particle reports its exact value.  MM_PROGRESS
 * uses it; the other models ignore lkmin_dev. */
int smcb_loglik_bounded(smcb_handle* h, int model, const double* theta_dev, int64_t ld, int64_t n, int d,
                        const uint8_t* active_dev, const double* lkmin_dev, double* lk_dev, void* stream);
/* Registers the callback smcb_loglik(..., SMCB_MODEL_USER, ...) dispatches to (NULL removes it). */
int smcb_set_user_likelihood(smcb_handle* h, smcb_user_loglik_fn fn, void* user_data);
/* Tunables.  SMCB_PARAM_MM_BUDGET: attempted RK steps after which the bulk MM_PROGRESS kernel hands a
 * solve, with its state, to the tail kernel (default 512; the Python engine picks 32 .. 512 by particle count).
 * A solve takes the same accepted / rejected steps whichever kernel finishes it; the two kernels spell the
 * arithmetic of a step differently (throughput / latency, csrc/mm_solver.cuh), so results depend on the budget
 * by rounding only (<= 1e-11 relative on a log-likelihood). */
#define SMCB_PARAM_MM_BUDGET 1
/* SMCB_PARAM_MM_REFILL_MIN: free lanes a warp of the bulk kernel waits for before it sets up new solves
 * (default 8, 1..32; results do not depend on it). */
#define SMCB_PARAM_MM_REFILL_MIN 2
/* SMCB_PARAM_MM_PATIENCE: ... and for how many attempted steps it waits (default 3); a warp whose lanes
 * are all free refills at once. */
#define SMCB_PARAM_MM_PATIENCE 3
/* SMCB_PARAM_PROFILE: non-zero = record CUDA events on the launching stream around the bulk and the tail
 * kernel of every MM_PROGRESS sweep (any number of sweeps between two reads: the event list grows on demand);
 * switching it on starts a new record; see smcb_profile_read. */
#define SMCB_PARAM_PROFILE 4
/* SMCB_PARAM_MM_CHUNK: particles per work-queue item of the bulk kernel (default 32). */
#define SMCB_PARAM_MM_CHUNK 5
/* SMCB_PARAM_MM_TAIL_WARPS: upper limit of the one-warp blocks per SM the tail kernel may use (default 32,
 * 1..32; it uses one per scheduler unless a sweep leaves it more than 128 solves per warp). */
#define SMCB_PARAM_MM_TAIL_WARPS 6
/* SMCB_PARAM_MM_INTEGRATOR: how MM_PROGRESS integrates dS/dt = -Vmax S/(Km+S) (EX/lik:14-33).
 *   SMCB_MM_RK45_SCIPY (default): scipy's adaptive RK45 taken step for step - the reference's likelihood (parity mode);
 *   SMCB_MM_EXACT: the closed form S(t) = Km*omega(ln(S0/Km) + (S0 - Vmax t)/Km), omega = Wright omega function -
 *   the converged solution of the same ODE (it differs from the reference's rtol-1e-3 result by up to
 *   2.9e-3 relative on the log-likelihood), cost independent of stiffness, early rejection per observation.  The mode for
 *   users who want the ODE's own likelihood; labelled wherever it is reported. */
#define SMCB_PARAM_MM_INTEGRATOR 7
#define SMCB_MM_RK45_SCIPY 0
#define SMCB_MM_EXACT 1
int smcb_set_param(smcb_handle* h, int key, double value);
/* Device time of the MM_PROGRESS kernels since the last read (SMCB_PARAM_PROFILE): out_host[0] = ms inside
 * mm_bulk_kernel, [1] = ms inside mm_tail_kernel, [2] = sweeps covered.  Synchronous; resets the record. */
int smcb_profile_read(smcb_handle* h, double* out_host);
/* Model predictions for a few particles (the `C_l_` the reference returns for its parity plots,
 * EX/lik:74-77): pred_dev[i][n_ex][n_t].  MM_PROGRESS only. */
int smcb_predict_mm_progress(smcb_handle* h, const double* theta_dev, int64_t ld, int64_t n,
                             double* pred_dev, void* stream);
/* Work counters of MM_PROGRESS sweeps, out_host int64[24]: [0]=RHS evaluations, [1]=accepted steps,
 * [2]=rejected steps, [3]=failed solves of the last sweep; [4..7] the same four accumulated over every
 * sweep since smcb_create; [8]/[9] particles reported -inf by early rejection (last sweep / accumulated);
 * [10] largest number of attempted steps of one solve in the last sweep; [11]/[12] solves deferred to
 * the tail kernel; [13]/[14] particles the tail kernel processed; [15] attempted steps taken inside the tail kernel
 * (accumulated); [16] longest solve the tail kernel ran since the previous call of this function, as
 * (attempts << 32) | device clock cycles per attempted step of that solve; [17..23] unused.  Synchronous. */
int smcb_loglik_stats(smcb_handle* h, int64_t* out_host);

/* ---- K2: tempering reductions (replaces EX/main:116-134) --------------------------------- */
/* out_dev[0] = max_i lk[i]  (NaN-free input assumed; -inf allowed). */
int smcb_lk_max(smcb_handle* h, const double* lk_dev, int64_t n, double* out_dev, void* stream);
/* For each candidate increment gm_k (k<n_cand<=SMCB_MAX_CAND, host array):
 *   out_dev[2k] = sum_i exp((lk_i-max)*gm_k),  out_dev[2k+1] = sum_i exp((lk_i-max)*gm_k)^2.
 * max_dev points to the (already all-reduced) maximum on the device. */
int smcb_temper_sums(smcb_handle* h, const double* lk_dev, int64_t n, const double* max_dev,
                     const double* gm_host, int n_cand, double* out_dev, void* stream);

/* Max and sums of one tempering round over ALL shards with a single exchange (north_star: "online logsumexp"):
 * every rank reduces its shard to (max_r, S1_r[k], S2_r[k]) relative to its OWN maximum, one all-gather moves the
 * rows, and every rank merges them in rank order with the logsumexp rescale
 *     max = max_r max_r,   S1[k] = sum_r S1_r[k]*exp((max_r-max)*gm_k),   S2[k] = sum_r S2_r[k]*exp(2(max_r-max)*gm_k).
 * n_cand <= 3*SMCB_MAX_CAND candidates (gm_host).  out_dev[0] = max, out_dev[1] is left alone (the engine keeps the
 * accepted sum there), out_dev[2+2k], out_dev[3+2k] = S1[k], S2[k].  Without a communicator (or world == 1) the
 * result is bit-identical to smcb_lk_max followed by smcb_temper_sums. */
int smcb_temper_eval(smcb_handle* h, const double* lk_dev, int64_t n, const double* gm_host, int n_cand,
                     double* out_dev, void* stream);

/* ---- K3: residual-systematic resampling (replaces EX/main:147-184) ----------------------- */
/* Normalised weights p_weight_i = exp((lk_i-max)*gm)/sum_w  (EX/main:124-130). */
int smcb_weights(smcb_handle* h, const double* lk_dev, int64_t n, const double* max_dev, double gm,
                 const double* sum_w_dev, double* w_dev, void* stream);
/* From normalised weights to copy counts.
 *   n_total  : global particle count N (inv_Np = 1/N), n: local particles,
 *   u0       : the single U[0,1) draw (EX/main:156),
 *   mode     : SMCB_SCAN_SEQUENTIAL or SMCB_SCAN_FIXED,
 *   carry_host[2] (SEQUENTIAL, may be NULL = {0, u0/N}): running sum and threshold entering this
 *              shard; on return (synchronous in that case) holds the values leaving it,
 *   carry_q  (FIXED): exclusive prefix of the fixed-point residual totals of lower ranks,
 *   id_offset: global index of this shard's first particle (the very first particle of the run
 *              starts with zero thresholds crossed, which matters only when u0 == 0),
 *   counts_dev int32[n]: copies per particle (floor + crossing),
 *   totals_dev: int64[2] = {sum of floor counts, sum of fixed-point residuals (FIXED) or crossings
 *              (SEQUENTIAL)} of this shard. */
int smcb_resample_counts(smcb_handle* h, const double* w_dev, int64_t n, int64_t n_total, double u0,
                         int mode, double* carry_host, uint64_t carry_q, int64_t id_offset,
                         int32_t* counts_dev, int64_t* totals_dev, void* stream);
/* First half of the FIXED mode for sharded runs: floor counts and fixed-point residuals only.
 * totals_dev int64[2] = {sum floor, sum q}.  Call smcb_resample_counts afterwards with carry_q. */
int smcb_resample_totals(smcb_handle* h, const double* w_dev, int64_t n, int64_t n_total,
                         int64_t* totals_dev, void* stream);
/* Expand counts into the non-decreasing ancestor vector: ancestors_dev[s] for s<m is the local
 * index of the particle copied into slot s.  If sum(counts) < m the tail is padded with the last
 * ancestor; copies beyond m are dropped.  filled_dev int64[1] = sum(counts). */
int smcb_ancestors(smcb_handle* h, const int32_t* counts_dev, int64_t n, int64_t m,
                   int32_t* ancestors_dev, int64_t* filled_dev, void* stream);
/* Vectorised gather of particle state: dst[k*ld_dst+s] = src[k*ld_src+anc[s]] for k<rows, s<m. */
int smcb_gather(smcb_handle* h, const double* src_dev, int64_t ld_src, const int32_t* ancestors_dev,
                int64_t m, int rows, double* dst_dev, int64_t ld_dst, void* stream);

/* The whole of K3 for one shard in ONE kernel (FIXED arithmetic): weights (from lk_dev, max_dev, gm, sum_w_dev exactly as
 * smcb_weights computes them, or explicit normalised weights w_dev with lk_dev = NULL), floor counts and fixed-point
 * residuals, copy counts, output offsets (decoupled look-back over 2048-particle tiles), ancestors and the gather
 * dst[k*ld_dst+s] = src[k*ld_src+anc[s]] of `rows` rows for the first m_out output slots of this shard (more copies
 * are dropped, fewer are padded with the last ancestor).  Same counts, ancestors and clamp / pad rule as
 * smcb_resample_counts(SMCB_SCAN_FIXED) + smcb_ancestors + smcb_gather, bit for bit.
 *   One GPU: n_total = n, carry_q = 0, id_offset = 0, m_out = n.
 *   Sharded: n_total = N, carry_q = exclusive prefix of the fixed-point residual totals of the lower ranks
 *   (smcb_resample_totals + all-gather), id_offset = global index of this shard's first particle, m_out = the slots
 *   this shard fills (the rows of dst then leave with smcb_comm_exchange_rows).
 * ancestors_dev / counts_dev may be NULL; filled_dev int64[1] = sum of this shard's counts. */
int smcb_resample_fused(smcb_handle* h, const double* lk_dev, const double* w_dev, int64_t n, int64_t n_total,
                        uint64_t carry_q, int64_t id_offset, int64_t m_out, const double* max_dev, double gm,
                        const double* sum_w_dev, double u0, const double* src_dev, int64_t ld_src, int rows,
                        double* dst_dev, int64_t ld_dst, int32_t* ancestors_dev, int32_t* counts_dev,
                        int64_t* filled_dev, void* stream);

/* ---- K4: Metropolis-Hastings mutation (replaces EX/main:209-249) ------------------------- */
/* Column sums: out_dev[k] = sum_i theta[k][i]. */
int smcb_colsum(smcb_handle* h, const double* theta_dev, int64_t ld, int64_t n, int d,
                double* out_dev, void* stream);
/* Centred second moments: out_dev[a*d+b] = sum_i (theta[a][i]-mean[a])*(theta[b][i]-mean[b]).
 * mean_dev: device pointer to d doubles (the all-reduced mean). */
int smcb_centered_moments(smcb_handle* h, const double* theta_dev, int64_t ld, int64_t n, int d,
                          const double* mean_dev, double* out_dev, void* stream);
/* Proposal: prop = theta + (z @ F) * ratio, z ~ N(0,I_d)  (EX/main:220), followed by the box test
 * of the uniform prior (EX/main:224-228): inbox[i] = all_k low_k <= prop_k <= high_k; out-of-box
 * proposals are replaced by the current particle.
 *   F_host[d*d] row-major factor (x = z @ F), low_host/high_host[d].
 *   z_dev: external normals [n][d] row-major (parity mode) or NULL = Philox4x32-10 keyed by
 *   (seed, global particle id = id_offset+i, stage, sweep). */
int smcb_mh_propose(smcb_handle* h, const double* theta_dev, int64_t ld, int64_t n, int d,
                    const double* F_host, double ratio, const double* low_host, const double* high_host,
                    const double* z_dev, uint64_t seed, uint64_t id_offset, uint32_t stage, uint32_t sweep,
                    double* prop_dev, int64_t ld_prop, uint8_t* inbox_dev, void* stream);
/* The same proposal with the factor read from device memory (F_dev[d*d] row-major, e.g. the one smcb_moments_merged
 * left there), so that no host round trip sits between the moment reduction and the proposal. */
int smcb_mh_propose_dev(smcb_handle* h, const double* theta_dev, int64_t ld, int64_t n, int d,
                        const double* F_dev, double ratio, const double* low_host, const double* high_host,
                        const double* z_dev, uint64_t seed, uint64_t id_offset, uint32_t stage, uint32_t sweep,
                        double* prop_dev, int64_t ld_prop, uint8_t* inbox_dev, void* stream);
/* Moments of ALL shards with a single exchange, plus the proposal factor, on the device (EX/main:212-215 and the
 * factor inside np.random.multivariate_normal at :220).  Every rank reduces its shard to (n_r, mean_r, M2_r) with
 * the two-pass arithmetic of np.cov (M2_r centred on the shard's own mean), appends the four MH counters of
 * smcb_mh_accept (counts_dev, may be NULL), one all-gather moves the rows and every rank merges them in rank order
 * (Chan et al.): mean = sum n_r mean_r / N,  M2 = sum [M2_r + n_r (mean_r-mean)(mean_r-mean)^T].
 *   out_dev layout (doubles): [0:4] counters summed over ranks, [4:4+d] mean, [4+d : 4+d+d*d] M2 (the centred
 *   second-moment SUM; cov = M2/n_total), [4+d+d*d : 4+d+2*d*d] factor F with x = z @ F.
 * The factor is built from cov (*) w_cov (Hadamard product, EX/main:215; w_cov_host[d*d], NULL = all ones) by a cyclic
 * Jacobi eigen-decomposition on the device, cov (*) w_cov = V diag(lambda) V^T:  F[j][:] = sqrt(|lambda_j|) * v_j with the
 * eigenvalues in descending order, every eigenvector signed so that its largest component is positive and
 * |lambda_j| <= 1e-13 max|lambda| treated as 0.  F^T F = |cov (*) w_cov|, which is what NumPy's SVD-based sampler
 * draws from; NumPy's own factor (LAPACK's sign and ordering choices) is reproduced only by the host path
 * (smcb_mh_propose with F_host), which the parity mode with external normals uses.
 * With world == 1 mean and M2 are bit-identical to smcb_colsum / n followed by smcb_centered_moments. */
int smcb_moments_merged(smcb_handle* h, const double* theta_dev, int64_t ld, int64_t n, int d, int64_t n_total,
                        const int64_t* counts_dev, const double* w_cov_host, double* out_dev, void* stream);
/* Accept step (EX/main:231-241): r = exp((lk2-lk1)*gamma)*inbox [* exp(dlp)] >= u; theta/lk updated in place;
 * moved_dev[i] |= r; counts_dev int64[4] += {accepted this sweep, newly moved, in-box proposals
 * (= likelihood evaluations this sweep requested), in-box proposals whose lk2 is -inf (= rejected early
 * by smcb_loglik_bounded before every observation was integrated)}.
 *   u_dev: external uniforms [n] or NULL = Philox (same key, separate stream id). */
int smcb_mh_accept(smcb_handle* h, double* theta_dev, int64_t ld, double* lk_dev, const double* prop_dev,
                   int64_t ld_prop, const double* lk2_dev, const uint8_t* inbox_dev, int64_t n, int d,
                   double gamma, const double* u_dev, const double* dlp_dev, uint64_t seed, uint64_t id_offset,
                   uint32_t stage, uint32_t sweep, uint8_t* moved_dev, int64_t* counts_dev, void* stream);
/* Early-rejection threshold for smcb_loglik_bounded: with u the uniform smcb_mh_accept will draw for
 * particle i (same u_dev / Philox key), the proposal is certainly rejected if
 *     lk2 < lkmin[i] = lk1[i] + log(u)/gamma - margin,   margin = 1e-9*(1 + |lk1| + |log u|/gamma),
 * because then exp((lk2-lk1)*gamma) < u even after rounding.  u == 0, gamma <= 0 or a non-finite lk1 give
 * -inf (never reject early).  Out-of-box particles (inbox == 0) get -inf as well; they are not evaluated. */
int smcb_mh_threshold(smcb_handle* h, const double* lk_dev, const uint8_t* inbox_dev, int64_t n, double gamma,
                      const double* u_dev, const double* dlp_dev, uint64_t seed, uint64_t id_offset, uint32_t stage,
                      uint32_t sweep, double* lkmin_dev, void* stream);
/* Mixed normal / uniform priors (the reference's `pp = exp(px*gamma) * (p0_2/p0_1)`, SMC_methanation_main.py:359-375,
 * cal_prior SMC_example/Micmem_SMC_main.py:60-90): out_dev[i] = log p(prop_i) - log p(theta_i) over the normally
 * distributed parameters, sum_k inv2var_k*((theta_k-mu_k)^2 - (prop_k-mu_k)^2) with inv2var_k = 1/(2 sigma_k^2), 0 for a
 * uniform parameter (those enter through the box test of smcb_mh_propose; give them low/high, and -inf/+inf to the
 * normal ones).  Pass the result as dlp_dev to smcb_mh_threshold and smcb_mh_accept (NULL = uniform priors only). */
int smcb_prior_logratio(smcb_handle* h, const double* theta_dev, int64_t ld, const double* prop_dev, int64_t ld_prop,
                        int64_t n, int d, const double* mu_host, const double* inv2var_host, const uint8_t* inbox_dev,
                        double* out_dev, void* stream);
/* Several MH sweeps in one call with a frozen proposal factor (documented deviation from the per-sweep covariance
 * refresh of EX/main:212) and no host round trip in between: per sweep a propose kernel (Philox normals, factor
 * mat-vec, box test; survivors packed into a list), the reactor marches of the survivors in full warps, and an
 * accept kernel.  At most 64 sweeps per call.  KINETIC_RK.  counts_dev int64[3] as for smcb_mh_accept. */
int smcb_mh_fused(smcb_handle* h, int model, double* theta_dev, int64_t ld, double* lk_dev, int64_t n, int d,
                  const double* F_host, double ratio, const double* low_host, const double* high_host,
                  double gamma, int n_sweeps, uint64_t seed, uint64_t id_offset, uint32_t stage,
                  uint32_t sweep0, uint8_t* moved_dev, int64_t* counts_dev, void* stream);

/* ---- communicator (SURVEY.md 8(b) smcb_comm_init, 8(e)): NCCL, one process per GPU ----------------------------
 * The reference has no exchange step (its only parallelism is the local ray fan-out, EX/lik:83-87).  Rank 0 makes
 * a unique id (smcb_comm_unique_id, 128 bytes), the host program hands it to every rank by any means it likes and
 * every rank calls smcb_comm_init on its handle.  All collectives below are enqueued on `stream` (no host
 * synchronisation); with world == 1 they degenerate to device copies or no-ops.  libnccl.so.2 is opened with
 * dlopen on first use (environment variable SMCB_NCCL_LIB overrides the name). */
#define SMCB_COMM_ID_BYTES 128
#define SMCB_OP_SUM 0
#define SMCB_OP_MAX 1
int smcb_comm_unique_id(void* out_host, int nbytes);
int smcb_comm_init(smcb_handle* h, const void* id_host, int nbytes, int rank, int world);
int smcb_comm_destroy(smcb_handle* h);
int smcb_comm_rank(const smcb_handle* h);
int smcb_comm_world(const smcb_handle* h);
/* recv_dev[r*bytes_per_rank ...] = send_dev[...] of rank r. */
int smcb_comm_all_gather(smcb_handle* h, const void* send_dev, void* recv_dev, int64_t bytes_per_rank, void* stream);
/* In-place all-reduce of count doubles, op = SMCB_OP_SUM | SMCB_OP_MAX. */
int smcb_comm_all_reduce_f64(smcb_handle* h, double* buf_dev, int64_t count, int op, void* stream);
int smcb_comm_broadcast(smcb_handle* h, void* buf_dev, int64_t bytes, int root, void* stream);
/* Particle migration of the sharded resampling: send_counts_host[q] doubles go to rank q from consecutive ranges
 * of send_dev (rank order), recv_counts_host[q] doubles arrive from rank q into consecutive ranges of recv_dev. */
int smcb_comm_all_to_all_v(smcb_handle* h, const double* send_dev, const int64_t* send_counts_host,
                           double* recv_dev, const int64_t* recv_counts_host, void* stream);
/* The same migration straight from / into row-major matrices, without packing: for every rank q the column range
 * [soff_q, soff_q + send_counts_host[q]) of each of the `rows` rows of send_dev (leading dimension ld_send; soff = running
 * sum of the send counts) goes to rank q, and what rank q sends lands in the column range [roff_q, roff_q +
 * recv_counts_host[q]) of the rows of recv_dev (ld_recv).  One grouped NCCL call; the part that stays is a 2-D copy. */
int smcb_comm_exchange_rows(smcb_handle* h, const double* send_dev, int64_t ld_send, const int64_t* send_counts_host,
                            double* recv_dev, int64_t ld_recv, const int64_t* recv_counts_host, int rows, void* stream);
/* Number of NCCL operations this handle has enqueued so far. */
int64_t smcb_collective_count(const smcb_handle* h);

/* Several MH sweeps in one call for ANY model, covariance refreshed every sweep (EX/main:212) because the factor stays on
 * the device: per sweep smcb_mh_propose_dev (factor = the one in blk_dev) -> smcb_mh_threshold (early_reject != 0) ->
 * smcb_loglik_bounded on the in-box proposals -> smcb_mh_accept -> smcb_moments_merged (counters of all shards +
 * moments + next factor into blk_dev).  blk_dev must hold the result of a smcb_moments_merged call on the current
 * particles at entry.  The reference's early-exit and step-halving rules (EX/main:243-248) are the caller's, between
 * calls (n_sweeps = 1 reproduces them exactly).  Uniform (box) priors, Philox random inputs. */
int smcb_mh_sweeps(smcb_handle* h, int model, double* theta_dev, int64_t ld, double* lk_dev, int64_t n, int d,
                   int64_t n_total, const double* w_cov_host, double ratio, const double* low_host,
                   const double* high_host, double gamma, int n_sweeps, int early_reject, uint64_t seed,
                   uint64_t id_offset, uint32_t stage, uint32_t sweep0, double* prop_dev, int64_t ld_prop,
                   double* lk2_dev, double* lkmin_dev, uint8_t* inbox_dev, uint8_t* moved_dev, int64_t* counts_dev,
                   double* blk_dev, void* stream);

/* ---- utilities ---------------------------------------------------------------------------- */
/* Philox draws exactly as the kernels make them, for tests: z_dev [n][d], u_dev [n] (either NULL). */
int smcb_philox_draws(smcb_handle* h, int64_t n, int d, uint64_t seed, uint64_t id_offset, uint32_t stage,
                      uint32_t sweep, double* z_dev, double* u_dev, void* stream);
/* Uniform prior sample: theta[k][i] = low_k + (high_k-low_k)*U, Philox keyed (seed, id, 0xFFFFFFFF, k). */
int smcb_sample_uniform_box(smcb_handle* h, double* theta_dev, int64_t ld, int64_t n, int d,
                            const double* low_host, const double* high_host, uint64_t seed,
                            uint64_t id_offset, void* stream);
/* Stream-ordered memset to zero (the per-stage reset of the moved mask and the MH counters, EX/main:187-190). */
int smcb_zero(smcb_handle* h, void* ptr_dev, int64_t bytes, void* stream);
/* Stream-ordered 2-D device copy, dst[k*ld_dst + i] = src[k*ld_src + i] for k < rows, i < width (unpacks a received
 * [rows][width] particle chunk into the SoA state). */
int smcb_copy_rows(smcb_handle* h, const double* src_dev, int64_t ld_src, double* dst_dev, int64_t ld_dst,
                   int64_t width, int rows, void* stream);
/* Sustained FP64 FMA and FP32 FMA rates of this device (micro-benchmark, used as the roofline
 * denominator for the compute-bound likelihood kernels): out_host[0]=FP64 FLOP/s, [1]=FP32 FLOP/s.
 * Synchronous. */
int smcb_measure_fma_peak(smcb_handle* h, double* out_host);

#ifdef __cplusplus
}
#endif
#endif /* SMCB200_H */
