// K1: Michaelis-Menten log-likelihoods.
//
// MM_PROGRESS replaces `log_likelihood_mm_multi` / `simulate_mm_on_grid` / `mm_ode`
// (reference SMC_example/Micmem_likelihood.py:14-77).  The reference integrates
// dS/dt = -Vmax*S/(Km+S) with scipy's adaptive RK45 at rtol=1e-3, atol=1e-6 and reads the
// solution through the quartic dense output, so the likelihood is defined by that controller.
// The device code below takes the same steps (scipy/_ivp/rk.py:61-180,538-567,
// common.py:110-134, ivp.py:712-728) in FP64.
//
// Mapping: one *lane* integrates one (particle, experiment) solve at a time.  Solves need
// 40..3000 RHS evaluations depending on (Vmax, Km), so a static lane->solve map would leave
// most of a warp idle.  Instead every block owns a contiguous range of solves and its lanes
// pull the next one from a shared-memory queue head (warp-aggregated atomic) whenever they
// finish; the step body itself is executed convergently by all lanes that hold a solve.
// Observation data (t, P_obs, S0) is staged once per block in shared memory.
#include "common.cuh"

namespace {

constexpr double RTOL = 1e-3, ATOL = 1e-6, SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0;

// Dormand-Prince tableau exactly as scipy spells it (rk.py:538-567); the quotients are
// evaluated by the compiler in FP64 just as CPython evaluates them.
constexpr double C2 = 1.0 / 5, C3 = 3.0 / 10, C4 = 4.0 / 5, C5 = 8.0 / 9;
constexpr double A21 = 1.0 / 5;
constexpr double A31 = 3.0 / 40, A32 = 9.0 / 40;
constexpr double A41 = 44.0 / 45, A42 = -56.0 / 15, A43 = 32.0 / 9;
constexpr double A51 = 19372.0 / 6561, A52 = -25360.0 / 2187, A53 = 64448.0 / 6561, A54 = -212.0 / 729;
constexpr double A61 = 9017.0 / 3168, A62 = -355.0 / 33, A63 = 46732.0 / 5247, A64 = 49.0 / 176,
                 A65 = -5103.0 / 18656;
constexpr double B1 = 35.0 / 384, B3 = 500.0 / 1113, B4 = 125.0 / 192, B5 = -2187.0 / 6784, B6 = 11.0 / 84;
constexpr double E1 = -71.0 / 57600, E3 = 71.0 / 16695, E4 = -71.0 / 1920, E5 = 17253.0 / 339200,
                 E6 = -22.0 / 525, E7 = 1.0 / 40;
// dense-output matrix P (7 x 4); row 2 is zero.
constexpr double P11 = 1.0, P12 = -8048581381.0 / 2820520608, P13 = 8663915743.0 / 2820520608,
                 P14 = -12715105075.0 / 11282082432;
constexpr double P32 = 131558114200.0 / 32700410799, P33 = -68118460800.0 / 10900136933,
                 P34 = 87487479700.0 / 32700410799;
constexpr double P42 = -1754552775.0 / 470086768, P43 = 14199869525.0 / 1410260304,
                 P44 = -10690763975.0 / 1880347072;
constexpr double P52 = 127303824393.0 / 49829197408, P53 = -318862633887.0 / 49829197408,
                 P54 = 701980252875.0 / 199316789632;
constexpr double P62 = -282668133.0 / 205662961, P63 = 2019193451.0 / 616988883,
                 P64 = -1453857185.0 / 822651844;
constexpr double P72 = 40617522.0 / 29380423, P73 = -110615467.0 / 29380423, P74 = 69997945.0 / 29380423;

__device__ __forceinline__ double mm_rhs(double nVmax, double Km, double S) {
    // python: -Vmax * S / (Km + S)  ==  ((-Vmax)*S)/(Km+S)
    return (nVmax * S) / (Km + S);
}

__device__ __forceinline__ double ulp10(double t) {
    // 10 * |nextafter(t, +inf) - t|
    double nx = __longlong_as_double(__double_as_longlong(t) + (t >= 0.0 ? 1 : -1));
    if (t == 0.0) nx = __longlong_as_double(1LL);
    return 10.0 * fabs(nx - t);
}

struct SolveState {
    double nVmax, Km, S0;
    double t, y, f, h_abs;
    double acc;      // residual sum of squares
    int i_eval;      // next t_eval index
    int task;        // -1 = none
    int e;
};

constexpr int BLOCK = 128;

// MODE 0: residual sum of squares into ssr[e*n + p];  MODE 1: predictions P_model into pred.
template <int MODE>
__global__ void __launch_bounds__(BLOCK)
mm_progress_kernel(const double* __restrict__ theta, int64_t ld, int64_t n,
                   const uint8_t* __restrict__ active, const double* __restrict__ g_t,
                   const double* __restrict__ g_P, const double* __restrict__ g_S0, int n_ex, int n_t,
                   int tasks_per_block, double* __restrict__ out, unsigned long long* __restrict__ stats) {
    extern __shared__ double smem[];
    double* s_t = smem;
    double* s_P = smem + (size_t)n_ex * n_t;
    double* s_S0 = s_P + (size_t)n_ex * n_t;
    __shared__ int s_next;

    for (int i = threadIdx.x; i < n_ex * n_t; i += BLOCK) {
        s_t[i] = g_t[i];
        s_P[i] = g_P[i];
    }
    for (int i = threadIdx.x; i < n_ex; i += BLOCK) s_S0[i] = g_S0[i];
    const int64_t total = n * (int64_t)n_ex;
    const int64_t task_lo = (int64_t)blockIdx.x * tasks_per_block;
    int64_t rem = total - task_lo;
    const int n_tasks = (int)(rem < tasks_per_block ? rem : tasks_per_block);
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;

    SolveState st;
    st.task = -1;
    bool exhausted = false;
    unsigned long long n_fev = 0, n_acc = 0, n_rej = 0, n_fail = 0;
    int64_t p = 0;
    const double* tt = s_t;
    const double* pp = s_P;

    while (true) {
        // ---- refill: lanes without a solve pull the next task of this block --------------
        while (true) {
            const bool want = (st.task < 0) && !exhausted;
            const unsigned need = __ballot_sync(FULL_MASK, want);
            if (need == 0) break;
            const int leader = __ffs(need) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&s_next, __popc(need));
            base = __shfl_sync(FULL_MASK, base, leader);
            if (want) {
                const int tsk = base + __popc(need & lt_mask);
                if (tsk >= n_tasks) {
                    exhausted = true;
                } else {
                    const int64_t g = task_lo + tsk;
                    p = g / n_ex;
                    const int e = (int)(g - p * n_ex);
                    if (active == nullptr || active[p]) {
                        st.task = tsk;
                        st.e = e;
                        st.nVmax = -theta[p];
                        st.Km = theta[ld + p];
                        st.S0 = s_S0[e];
                        tt = s_t + (size_t)e * n_t;
                        pp = s_P + (size_t)e * n_t;
                        st.t = tt[0];
                        st.y = st.S0;
                        st.acc = 0.0;
                        st.i_eval = 0;
                        st.f = mm_rhs(st.nVmax, st.Km, st.y);
                        // select_initial_step (common.py:110-134), n=1, direction=+1, order=4
                        const double t_bound = tt[n_t - 1];
                        const double interval = fabs(t_bound - st.t);
                        const double scale = ATOL + fabs(st.y) * RTOL;
                        const double d0 = fabs(st.y / scale);
                        const double d1 = fabs(st.f / scale);
                        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
                        h0 = (interval < h0) ? interval : h0;
                        const double y1 = st.y + h0 * st.f;
                        const double f1 = mm_rhs(st.nVmax, st.Km, y1);
                        const double d2 = fabs((f1 - st.f) / scale) / h0;
                        double h1;
                        if (d1 <= 1e-15 && d2 <= 1e-15) {
                            h1 = h0 * 1e-3;
                            h1 = (h1 > 1e-6) ? h1 : 1e-6;
                        } else {
                            const double dm = (d2 > d1) ? d2 : d1;
                            h1 = pow(0.01 / dm, 1.0 / 5.0);
                        }
                        double hh = 100 * h0;
                        hh = (h1 < hh) ? h1 : hh;
                        hh = (interval < hh) ? interval : hh;
                        st.h_abs = hh;
                        n_fev += 2;
                        if (interval == 0.0) {   // degenerate grid: nothing to integrate
                            if (MODE == 0) out[(int64_t)e * n + p] = 0.0;
                            st.task = -1;
                        }
                    }
                }
            }
        }
        const bool have = st.task >= 0;
        if (!__any_sync(FULL_MASK, have)) break;   // all lanes exhausted and idle

        if (have) {
            // ---- one step attempt (rk.py:111-176) ----------------------------------------
            const double t_bound = tt[n_t - 1];
            const double t = st.t, y = st.y;
            const double min_step = ulp10(t);
            double h_abs = st.h_abs;
            // the clamp `h_abs < min_step -> min_step` applies at the start of a scipy step
            // (before the first attempt); a *rejected* attempt that drops below fails instead.
            // st.h_abs < 0 encodes "inside a step, after a rejection".
            bool rejected = false;
            if (h_abs < 0) {
                rejected = true;
                h_abs = -h_abs;
            } else if (h_abs < min_step) {
                h_abs = min_step;
            }
            if (h_abs < min_step) {
                // TOO_SMALL_STEP: scipy returns a short solution and the reference would raise.
                n_fail++;
                if (MODE == 0) out[(int64_t)st.e * n + p] = INFINITY;
                st.task = -1;
            } else {
                double t_new = t + h_abs;
                if (t_new - t_bound > 0) t_new = t_bound;
                const double h = t_new - t;
                h_abs = fabs(h);
                const double k1 = st.f;
                const double k2 = mm_rhs(st.nVmax, st.Km, y + (k1 * A21) * h);
                const double k3 = mm_rhs(st.nVmax, st.Km, y + (k1 * A31 + k2 * A32) * h);
                const double k4 = mm_rhs(st.nVmax, st.Km, y + (k1 * A41 + k2 * A42 + k3 * A43) * h);
                const double k5 = mm_rhs(st.nVmax, st.Km, y + (k1 * A51 + k2 * A52 + k3 * A53 + k4 * A54) * h);
                const double k6 =
                    mm_rhs(st.nVmax, st.Km, y + (k1 * A61 + k2 * A62 + k3 * A63 + k4 * A64 + k5 * A65) * h);
                const double y_new = y + h * (k1 * B1 + k3 * B3 + k4 * B4 + k5 * B5 + k6 * B6);
                const double k7 = mm_rhs(st.nVmax, st.Km, y_new);
                n_fev += 6;
                const double ay = fabs(y), ayn = fabs(y_new);
                const double scale = ATOL + ((ayn > ay || ayn != ayn) ? ayn : ay) * RTOL;
                const double err =
                    fabs(((k1 * E1 + k3 * E3 + k4 * E4 + k5 * E5 + k6 * E6 + k7 * E7) * h) / scale);
                if (err < 1.0) {
                    double factor;
                    if (err == 0.0) {
                        factor = MAX_FACTOR;
                    } else {
                        factor = SAFETY * pow(err, -0.2);
                        factor = (factor < MAX_FACTOR) ? factor : MAX_FACTOR;
                    }
                    if (rejected) factor = (factor < 1.0) ? factor : 1.0;
                    st.h_abs = h_abs * factor;
                    n_acc++;
                    // ---- dense output for every t_eval in (t_old, t_new] (ivp.py:712-728) ----
                    int i = st.i_eval;
                    if (i < n_t && tt[i] <= t_new) {
                        const double q1 = k1 * P11;
                        const double q2 = k1 * P12 + k3 * P32 + k4 * P42 + k5 * P52 + k6 * P62 + k7 * P72;
                        const double q3 = k1 * P13 + k3 * P33 + k4 * P43 + k5 * P53 + k6 * P63 + k7 * P73;
                        const double q4 = k1 * P14 + k3 * P34 + k4 * P44 + k5 * P54 + k6 * P64 + k7 * P74;
                        do {
                            const double x = (tt[i] - t) / h;
                            const double x2 = x * x, x3 = x2 * x, x4 = x3 * x;
                            const double S = h * (q1 * x + q2 * x2 + q3 * x3 + q4 * x4) + y;
                            const double Pm = st.S0 - S;
                            if (MODE == 0) {
                                const double r = pp[i] - Pm;
                                st.acc += r * r;
                            } else {
                                out[(p * n_ex + st.e) * (int64_t)n_t + i] = Pm;
                            }
                            ++i;
                        } while (i < n_t && tt[i] <= t_new);
                        st.i_eval = i;
                    }
                    st.t = t_new;
                    st.y = y_new;
                    st.f = k7;
                    if (t_new - t_bound >= 0) {   // finished
                        if (MODE == 0) out[(int64_t)st.e * n + p] = st.acc;
                        st.task = -1;
                    }
                } else {
                    double factor = SAFETY * pow(err, -0.2);
                    factor = (factor > MIN_FACTOR) ? factor : MIN_FACTOR;
                    st.h_abs = -(h_abs * factor);   // stay inside this scipy step
                    n_rej++;
                }
            }
        }
    }
    // ---- work counters ---------------------------------------------------------------------
    n_fev = warp_sum_ll((long long)n_fev);
    n_acc = warp_sum_ll((long long)n_acc);
    n_rej = warp_sum_ll((long long)n_rej);
    n_fail = warp_sum_ll((long long)n_fail);
    if (lane == 0 && stats != nullptr) {
        atomicAdd(&stats[0], n_fev);
        atomicAdd(&stats[1], n_acc);
        atomicAdd(&stats[2], n_rej);
        if (n_fail) atomicAdd(&stats[3], n_fail);
        atomicAdd(&stats[4], n_fev);   // cumulative since smcb_create (never reset by a sweep)
        atomicAdd(&stats[5], n_acc);
        atomicAdd(&stats[6], n_rej);
        if (n_fail) atomicAdd(&stats[7], n_fail);
    }
}

// lk[p] = sum_e [ -0.5*n_t*log(2*pi*sigma^2) - ssr_e/(2 sigma^2) ]   (Micmem_likelihood.py:70-73)
__global__ void mm_progress_finalize(const double* __restrict__ theta, int64_t ld, int64_t n,
                                     const uint8_t* __restrict__ active, const double* __restrict__ ssr,
                                     int n_ex, int n_t, double* __restrict__ lk) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (active != nullptr && !active[p]) return;
    const double sigma = theta[2 * ld + p];
    if (sigma <= 0) {   // Micmem_likelihood.py:53-54
        lk[p] = -INFINITY;
        return;
    }
    const double s2 = sigma * sigma;
    const double c0 = -0.5 * n_t * log(2 * M_PI * s2);
    const double den = 2 * s2;
    double total = 0.0;
    for (int e = 0; e < n_ex; ++e) total += c0 - ssr[(int64_t)e * n + p] / den;
    lk[p] = total;
}

// ------------------------------------------------------------------------------ MM_RATE
// ll = -0.5*n*log(2 pi sigma^2) - sum_i (v_i - Vmax*S_i/(Km+S_i))^2 / (2 sigma^2)
// One thread per particle; observations streamed through shared memory in tiles that every
// thread of the block reads by broadcast.
constexpr int RATE_BLOCK = 128;
constexpr int RATE_TILE = 2048;

__global__ void __launch_bounds__(RATE_BLOCK)
mm_rate_kernel_f64(const double* __restrict__ theta, int64_t ld, int64_t n,
                   const uint8_t* __restrict__ active, const double* __restrict__ gS,
                   const double* __restrict__ gv, int64_t n_obs, double* __restrict__ lk) {
    __shared__ double sS[RATE_TILE];
    __shared__ double sv[RATE_TILE];
    const int64_t p = (int64_t)blockIdx.x * RATE_BLOCK + threadIdx.x;
    const bool live = p < n && (active == nullptr || active[p]);
    double Vmax = 1.0, Km = 1.0, sigma = 1.0;
    if (live) {
        Vmax = theta[p];
        Km = theta[ld + p];
        sigma = theta[2 * ld + p];
    }
    double acc = 0.0;
    for (int64_t base = 0; base < n_obs; base += RATE_TILE) {
        const int m = (int)((n_obs - base < RATE_TILE) ? (n_obs - base) : RATE_TILE);
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += RATE_BLOCK) {
            sS[i] = gS[base + i];
            sv[i] = gv[base + i];
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < m; ++i) {
            const double S = sS[i];
            const double r = sv[i] - Vmax * S / (Km + S);
            acc = fma(r, r, acc);
        }
    }
    if (live) {
        if (sigma <= 0) {
            lk[p] = -INFINITY;
        } else {
            const double s2 = sigma * sigma;
            lk[p] = -0.5 * (double)n_obs * log(2 * M_PI * s2) - acc / (2 * s2);
        }
    }
}

// FP32 arithmetic: per observation  g = S*rcp(Km+S);  r = v - Vmax*g;  acc += r*r  in FP32 over
// 64-observation tiles, tile sums accumulated in FP64.
__global__ void __launch_bounds__(RATE_BLOCK)
mm_rate_kernel_f32(const double* __restrict__ theta, int64_t ld, int64_t n,
                   const uint8_t* __restrict__ active, const float2* __restrict__ gSv, int64_t n_obs,
                   double* __restrict__ lk) {
    __shared__ float2 sSv[RATE_TILE];
    const int64_t p = (int64_t)blockIdx.x * RATE_BLOCK + threadIdx.x;
    const bool live = p < n && (active == nullptr || active[p]);
    float Vmax = 1.f, Km = 1.f;
    double sigma = 1.0;
    if (live) {
        Vmax = (float)theta[p];
        Km = (float)theta[ld + p];
        sigma = theta[2 * ld + p];
    }
    double acc = 0.0;
    for (int64_t base = 0; base < n_obs; base += RATE_TILE) {
        const int m = (int)((n_obs - base < RATE_TILE) ? (n_obs - base) : RATE_TILE);
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += RATE_BLOCK) sSv[i] = gSv[base + i];
        __syncthreads();
        for (int i0 = 0; i0 < m; i0 += 64) {
            const int i1 = (i0 + 64 < m) ? i0 + 64 : m;
            float a0 = 0.f, a1 = 0.f;
            int i = i0;
            for (; i + 1 < i1; i += 2) {
                const float2 o0 = sSv[i], o1 = sSv[i + 1];
                const float g0 = __fdividef(o0.x, Km + o0.x);
                const float g1 = __fdividef(o1.x, Km + o1.x);
                const float r0 = fmaf(-Vmax, g0, o0.y);
                const float r1 = fmaf(-Vmax, g1, o1.y);
                a0 = fmaf(r0, r0, a0);
                a1 = fmaf(r1, r1, a1);
            }
            if (i < i1) {
                const float2 o0 = sSv[i];
                const float g0 = __fdividef(o0.x, Km + o0.x);
                const float r0 = fmaf(-Vmax, g0, o0.y);
                a0 = fmaf(r0, r0, a0);
            }
            acc += (double)(a0 + a1);
        }
    }
    if (live) {
        if (sigma <= 0) {
            lk[p] = -INFINITY;
        } else {
            const double s2 = sigma * sigma;
            lk[p] = -0.5 * (double)n_obs * log(2 * M_PI * s2) - acc / (2 * s2);
        }
    }
}

}  // namespace

int launch_loglik_mm_progress(smcb_handle* h, const double* theta, int64_t ld, int64_t n,
                              const uint8_t* active, double* lk, double* pred, cudaStream_t st) {
    const MmProgressData& D = h->mmp;
    REQUIRE(h, D.t != nullptr, SMCB_ERR_STATE, "smcb_set_data_mm_progress has not been called");
    if (n == 0) return SMCB_OK;
    const size_t smem = ((size_t)2 * D.n_ex * D.n_t + D.n_ex) * sizeof(double);
    REQUIRE(h, smem <= 200 * 1024, SMCB_ERR_UNSUPPORTED, "data set too large for shared-memory staging");
    const int64_t total = n * (int64_t)D.n_ex;
    int tasks_per_block = BLOCK * 8;
    // keep at least ~4 blocks per SM in flight for small n
    while (tasks_per_block > BLOCK && total / tasks_per_block < (int64_t)h->sm_count * 4) tasks_per_block >>= 1;
    const int64_t grid = (total + tasks_per_block - 1) / tasks_per_block;
    REQUIRE(h, grid < (1LL << 31), SMCB_ERR_UNSUPPORTED, "too many particles for one launch");
    if (pred != nullptr) {
        CUDA_TRY(h, cudaFuncSetAttribute(mm_progress_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        mm_progress_kernel<1><<<(unsigned)grid, BLOCK, smem, st>>>(theta, ld, n, active, D.t, D.P, D.S0, D.n_ex,
                                                                 D.n_t, tasks_per_block, pred, nullptr);
        LAUNCH_CHECK(h);
        return SMCB_OK;
    }
    REQUIRE(h, h->ssr != nullptr && n <= h->n_max && D.n_ex <= h->ssr_rows, SMCB_ERR_STATE,
            "smcb_reserve too small for this sweep");
    CUDA_TRY(h, cudaMemsetAsync(h->stats, 0, 4 * sizeof(unsigned long long), st));
    CUDA_TRY(h, cudaFuncSetAttribute(mm_progress_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    mm_progress_kernel<0><<<(unsigned)grid, BLOCK, smem, st>>>(theta, ld, n, active, D.t, D.P, D.S0, D.n_ex, D.n_t,
                                                             tasks_per_block, h->ssr, h->stats);
    LAUNCH_CHECK(h);
    const int fb = 256;
    mm_progress_finalize<<<(unsigned)((n + fb - 1) / fb), fb, 0, st>>>(theta, ld, n, active, h->ssr, D.n_ex, D.n_t,
                                                                     lk);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

int launch_loglik_mm_rate(smcb_handle* h, const double* theta, int64_t ld, int64_t n, const uint8_t* active,
                          double* lk, cudaStream_t st) {
    const MmRateData& D = h->mmr;
    REQUIRE(h, D.S != nullptr, SMCB_ERR_STATE, "smcb_set_data_mm_rate has not been called");
    if (n == 0) return SMCB_OK;
    const int64_t grid = (n + RATE_BLOCK - 1) / RATE_BLOCK;
    if (D.precision == 32)
        mm_rate_kernel_f32<<<(unsigned)grid, RATE_BLOCK, 0, st>>>(theta, ld, n, active, D.Sv32, D.n_obs, lk);
    else
        mm_rate_kernel_f64<<<(unsigned)grid, RATE_BLOCK, 0, st>>>(theta, ld, n, active, D.S, D.v, D.n_obs, lk);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}
