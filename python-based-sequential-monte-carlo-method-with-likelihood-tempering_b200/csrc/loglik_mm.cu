// K1: Michaelis-Menten log-likelihoods.
//
// MM_PROGRESS replaces `log_likelihood_mm_multi` / `simulate_mm_on_grid` / `mm_ode` and the ray fan-out
// of `sim_particle` (reference SMC_example/Micmem_likelihood.py:14-92).  The arithmetic of one solve -
// scipy's adaptive RK45 taken step for step - lives in mm_solver.cuh; this file is the mapping onto the
// machine.
//
// Work is n_ex solves per particle, each 10 .. 1e5 attempted steps depending on (Vmax, Km): in a prior cloud
// the median solve takes 12 attempts, one in a thousand takes more than 800 and the worst of 2^20 particles
// takes 8e4 strictly sequential ones; near the posterior every solve takes ~19.  One sweep is these launches:
//
//   mm_bin / mm_binscan / mm_scatter_kernel
//                       (bounded sweeps: residual limit per particle from its early-rejection threshold, then)
//                       counting sort of the particles that need solves by Vmax/Km (heaviest first); inactive,
//                       sigma <= 0 and hopeless particles get their -inf here and leave the work list.
//   mm_bulk_kernel      persistent warps, one solve per lane.  A warp draws runs of 32 cost-ordered particles of
//                       one experiment from a global queue (all experiments of the heavy bins come first), so
//                       its lanes take nearly the same steps; a lane whose solve ends waits until a few lanes
//                       of its warp have been free for a few steps (or none is busy) and they are then set up
//                       together.  A solve that needs more than `budget` attempts is parked (its state saved), which
//                       bounds the drain time of the kernel.
//   mm_finalize_kernel  one thread per evaluated particle: sums the experiments in the reference's order
//                       (Micmem_likelihood.py:70-73), applies the particle-level bound, or lists the particle's
//                       deferred solves for the tail kernel.
//   mm_tail_kernel      one deferred solve per lane, resumed where the bulk kernel left it, stepped by mmsolve::solve_lat (the
//                       latency spelling of the step).  These few thousand solves are latency-bound (~0.23 us per
//                       step); the kernel lasts as long as its longest solve.
//   mm_collect_kernel   ordered sum for the particles the tail kernel finished.
//
// Early rejection (MH sweeps).  Residuals only accumulate, so with c0 = -n_t/2 log(2 pi sigma^2)
//     n_ex*c0                                                        (before any solve: mm_bin_kernel)
//     sum_finished (c0 - ssr_e/(2 sigma^2)) + n_deferred*c0          (mm_finalize_kernel; minus one deferred
//                                                                     solve's running term in mm_tail_kernel)
// are upper bounds of the particle's log-likelihood.  Given lkmin[p] (smcb_mh_threshold: the value below
// which the Metropolis test of Micmem_SMC_main.py:231-236 is certain to reject, with a safety margin) a
// particle whose bound falls below it reports -inf: the accept/reject decision, hence the whole run, is
// exactly what it would have been.  Stiff proposals are almost always hopeless ones, so this removes most of
// the serial tail from MH sweeps; the first sweep has no threshold and keeps it.
#include <algorithm>

#include "common.cuh"
#include "exp_table.cuh"
#include "mm_solver.cuh"

namespace {

using mmsolve::Solve;

// Build switches kept for A/B runs: the tail kernel takes its steps in the latency spelling of mm_solver.cuh (solve_lat;
// 0 = the bulk kernel's attempt() alone) with the loop's coefficients in registers (0 = wherever the compiler puts them)
#ifndef SMCB_TAIL_LATENCY_FORM
#define SMCB_TAIL_LATENCY_FORM 1
#endif
#ifndef SMCB_TAIL_HOIST
#define SMCB_TAIL_HOIST 1
#endif

constexpr int BULK_BLOCK = 128;
constexpr int TAIL_BLOCK = 32;
// markers in the per-solve result array (a residual sum is >= 0): DEFERRED = restart the solve in the tail kernel,
// PARKED - slot = resume it from park[slot] (mmsolve::park_store / park_load, mm_solver.cuh)
constexpr double DEFERRED = -1.0;
constexpr double PARKED = -2.0;
constexpr int PARK_WORDS = mmsolve::PARK_WORDS;
using mmsolve::park_load;
using mmsolve::park_store;

// Shared-memory image of the data set: obs[n_ex][n_t] (P_obs[i], t[i+1]) pairs, then per experiment
// (S0, t[0], t[n_t-1]).
struct SharedData {
    mmsolve::ObsPair* obs;
    double* S0;
    double* t0;
    double* tb;
};
__host__ __device__ inline size_t shared_data_bytes(int n_ex, int n_t) {
    return (size_t)n_ex * n_t * sizeof(mmsolve::ObsPair) + (size_t)3 * n_ex * sizeof(double);
}
__device__ __forceinline__ SharedData stage_data(void* smem, const double* g_t, const double* g_P,
                                                 const double* g_S0, int n_ex, int n_t) {
    SharedData D;
    D.obs = reinterpret_cast<mmsolve::ObsPair*>(smem);
    D.S0 = reinterpret_cast<double*>(D.obs + (size_t)n_ex * n_t);
    D.t0 = D.S0 + n_ex;
    D.tb = D.t0 + n_ex;
    for (int k = threadIdx.x; k < n_ex * n_t; k += blockDim.x) {
        const int e = k / n_t, i = k - e * n_t;
        mmsolve::fill_pairs(D.obs + (size_t)e * n_t, g_t + (size_t)e * n_t, g_P + (size_t)e * n_t, n_t, i);
    }
    for (int e = threadIdx.x; e < n_ex; e += blockDim.x) {
        D.S0[e] = g_S0[e];
        D.t0[e] = g_t[(size_t)e * n_t];
        D.tb[e] = g_t[(size_t)e * n_t + n_t - 1];
    }
    __syncthreads();
    return D;
}

// counters: see smcb_loglik_stats
__device__ __forceinline__ void flush_stats(unsigned long long* stats, unsigned n_set, unsigned n_acc,
                                            unsigned n_rej, unsigned n_fail, unsigned n_def, unsigned mx,
                                            bool is_tail = false) {
    if (stats == nullptr) return;
    const unsigned long long w_fev =
        (unsigned long long)warp_sum_ll(2LL * n_set + 6LL * ((long long)n_acc + n_rej));
    const unsigned long long w_acc = (unsigned long long)warp_sum_ll((long long)n_acc);
    const unsigned long long w_rej = (unsigned long long)warp_sum_ll((long long)n_rej);
    const unsigned long long w_fail = (unsigned long long)warp_sum_ll((long long)n_fail);
    const unsigned long long w_def = (unsigned long long)warp_sum_ll((long long)n_def);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, o));
    if ((threadIdx.x & 31) == 0) {
        if (w_fev) { atomicAdd(&stats[0], w_fev); atomicAdd(&stats[4], w_fev); }
        if (w_acc) { atomicAdd(&stats[1], w_acc); atomicAdd(&stats[5], w_acc); }
        if (w_rej) { atomicAdd(&stats[2], w_rej); atomicAdd(&stats[6], w_rej); }
        if (w_fail) { atomicAdd(&stats[3], w_fail); atomicAdd(&stats[7], w_fail); }
        if (w_def) { atomicAdd(&stats[11], w_def); atomicAdd(&stats[12], w_def); }
        if (mx) atomicMax(&stats[10], (unsigned long long)mx);
        if (is_tail && (w_acc + w_rej)) atomicAdd(&stats[15], w_acc + w_rej);   // attempted steps of the tail kernel
    }
}

// ------------------------------------------------------------------------------ ordering by cost
// The number of steps scipy takes is, per experiment, almost a function of Vmax/Km alone (in a prior cloud
// warps of consecutive particles run at 15% lane efficiency, warps of particles sorted by Vmax/Km at 93%:
// DESIGN.md K1).  Every sweep therefore bins the particles it has to evaluate by Vmax/Km (8 bins per octave,
// counting sort, heaviest first) and the bulk kernel walks that permutation.  Particles that need no solve
// (inactive, sigma <= 0, hopeless before any residual) are left out, which also compacts the work list.
constexpr int NBIN = 512;
constexpr unsigned NOBIN = 0xFFFFu;

// Bounded sweeps first turn the particle's early-rejection threshold into a residual limit:
//   cutlim[p] = residual sum of squares above which ONE solve alone proves lk[p] < lkmin[p]:
//   n_ex*c0 - ssr/(2 sigma^2) < lkmin   <=>   ssr > (n_ex*c0 - lkmin) * 2 sigma^2
// (negative: hopeless before any residual).  Neighbouring particles of a posterior cloud fall into the same bin, so
// the lanes of a warp that share a bin add to the histogram once (match_any): one shared-memory atomic per warp
// instead of 32 on one address (36 -> 9 us per 2^20 particles).
template <bool BOUNDED>
__global__ void __launch_bounds__(256)
mm_bin_kernel(const double* __restrict__ theta, int64_t ld, unsigned n, const uint8_t* __restrict__ active,
              const double* __restrict__ lkmin, int n_ex, int n_t, double* __restrict__ cutlim,
              unsigned short* __restrict__ bins, unsigned* __restrict__ hist,
              double* __restrict__ lk, unsigned long long* __restrict__ stats) {
    __shared__ unsigned s_hist[NBIN];
    for (int i = threadIdx.x; i < NBIN; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    bool cut = false;
    unsigned b = NOBIN;
    if (p < n) {
        if (active == nullptr || active[p]) {
            const double sigma = theta[2 * ld + p];
            bool hopeless = false;
            if (BOUNDED) {
                double cl = INFINITY;
                const double thr = lkmin[p];
                if (sigma > 0 && thr > -INFINITY) {
                    const double s2 = sigma * sigma;
                    const double c0 = -0.5 * n_t * log(2 * M_PI * s2);
                    cl = (n_ex * c0 - thr) * (2 * s2);   // NaN compares false below
                }
                cutlim[p] = cl;
                hopeless = cl < 0;
            }
            if (!(sigma > 0)) {
                lk[p] = -INFINITY;                        // sigma <= 0 (Micmem_likelihood.py:53-54)
            } else if (hopeless) {
                lk[p] = -INFINITY;                        // n_ex*c0 < lkmin: hopeless before any residual
                cut = true;
            } else {
                const double r = theta[p] / theta[ld + p];   // Vmax / Km
                // exponent and top 3 mantissa bits: 8 bins per octave over 2^-32 .. 2^32; heaviest first
                int k = (__double2hiint(r) >> 17) - ((1023 - 32) << 3);
                if (!(r > 0)) k = 0;                      // zero, negative, NaN: cheapest bin
                k = k < 0 ? 0 : (k > NBIN - 1 ? NBIN - 1 : k);
                b = (unsigned)(NBIN - 1 - k);
            }
        }
        bins[p] = (unsigned short)b;
    }
    {
        const unsigned peers = __match_any_sync(FULL_MASK, b);
        if (b != NOBIN && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&s_hist[b], (unsigned)__popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NBIN; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
    if (BOUNDED) {
        const unsigned mk = __ballot_sync(FULL_MASK, cut);
        if (mk != 0 && (threadIdx.x & 31) == 0) {
            atomicAdd(&stats[8], (unsigned long long)__popc(mk));
            atomicAdd(&stats[9], (unsigned long long)__popc(mk));
        }
    }
}

// exclusive scan of the histogram (one block of NBIN threads): cursor[b] = first slot of bin b, ctl[3] = total
__global__ void __launch_bounds__(NBIN)
mm_binscan_kernel(const unsigned* __restrict__ hist, unsigned* __restrict__ cursor, unsigned* __restrict__ ctl) {
    __shared__ unsigned s[NBIN];
    const int t = threadIdx.x;
    s[t] = hist[t];
    __syncthreads();
    for (int o = 1; o < NBIN; o <<= 1) {
        const unsigned v = (t >= o) ? s[t - o] : 0u;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    cursor[t] = s[t] - hist[t];
    if (t == NBIN - 1) ctl[3] = s[t];
}

__global__ void __launch_bounds__(256)
mm_scatter_kernel(unsigned n, const unsigned short* __restrict__ bins, unsigned* __restrict__ cursor,
                  unsigned* __restrict__ perm) {
    __shared__ unsigned s_cnt[NBIN];
    __shared__ unsigned s_base[NBIN];
    for (int i = threadIdx.x; i < NBIN; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    unsigned b = NOBIN, rank = 0;
    if (p < n) b = bins[p];
    {
        // the lanes of a warp that share a bin take their slots with one atomic (lane order inside the group)
        const unsigned peers = __match_any_sync(FULL_MASK, b);
        const int leader = __ffs(peers) - 1;
        unsigned base = 0;
        if (b != NOBIN && (int)lane == leader) base = atomicAdd(&s_cnt[b], (unsigned)__popc(peers));
        base = __shfl_sync(FULL_MASK, base, leader);
        rank = base + (unsigned)__popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NBIN; i += blockDim.x)
        if (s_cnt[i]) s_base[i] = atomicAdd(&cursor[i], s_cnt[i]);
    __syncthreads();
    if (b != NOBIN) perm[s_base[b] + rank] = p;
}

// ------------------------------------------------------------------------------ bulk
// No early rejection inside this kernel: a solve that stopped at a random step would leave its lane idle
// until the next refill, and with cost-ordered warps that costs more than the steps it saves (measured: sweeps
// with a third of the solves cut ran at 40% of the attempt rate of uncut ones).  Hopeless particles never get
// here (mm_bin_kernel), the bound is applied per particle in mm_finalize_kernel and per solve in mm_tail_kernel.
__global__ void __launch_bounds__(BULK_BLOCK)
mm_bulk_kernel(const double* __restrict__ theta, int64_t ld, unsigned n, const unsigned* __restrict__ perm,
               const double* __restrict__ g_t, const double* __restrict__ g_P,
               const double* __restrict__ g_S0, int n_ex, int n_t, unsigned budget, int refill_min,
               int patience, unsigned chunk, double* __restrict__ ssr_out, unsigned* __restrict__ queue,
               double* __restrict__ park, unsigned park_cap, unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SharedData D = stage_data(smem, g_t, g_P, g_S0, n_ex, n_t);

    // Work list: the m particles of `perm` (heaviest cost bin first) in runs of `chunk`, each run once per
    // experiment; queue item q = run (q / n_ex) of experiment (q % n_ex).  A warp's lanes thus hold consecutive
    // particles of one experiment, and every experiment of the heavy bins is started early.
    const unsigned m = queue[3];
    const unsigned n_items = ((m + chunk - 1) / chunk) * (unsigned)n_ex;
    unsigned w_base = 0, w_e = 0;    // first particle (index into perm) and experiment of the warp's current run
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned w_cur = 0, w_end = 0;   // the warp's current run of solves (warp-uniform)
    bool drained = false;            // queue exhausted (warp-uniform)

    Solve s;
    s.nVmax = -1.0; s.Km = 1.0; s.S0 = 0.0; s.t = 0.0; s.y = 0.0; s.f = 0.0; s.h_abs = 0.0; s.ssr = 0.0;
    s.cut_lim = INFINITY; s.i_eval = 0; s.rejected = 0; s.t_next = 0.0; s.t_bound = 0.0;
    bool have = false;
    unsigned task = 0, n_att = 0;
    const mmsolve::ObsPair* obs = D.obs;
    // work counters: a solve that cannot be parked is redone from scratch by the tail kernel, so its attempts here are
    // dropped again (acc0/rej0 = counters when the solve started) and every step is counted once
    unsigned n_set = 0, n_acc = 0, n_rej = 0, n_fail = 0, n_def = 0, mx = 0, acc0 = 0, rej0 = 0;

    int waited = 0;   // attempted steps since a lane of this warp went free (warp-uniform)
    for (;;) {
        const unsigned busy = __ballot_sync(FULL_MASK, have);
        unsigned want = drained ? 0u : (~busy);
        // Refill when the whole warp is free, or when at least refill_min lanes have been waiting for
        // `patience` steps: near the posterior the lanes of a warp finish within a step or two of each
        // other and are set up together; in a prior cloud the stragglers are not waited for.
        if (want != 0 && (busy == 0 || (__popc(want) >= refill_min && waited >= patience))) {
            waited = 0;
            // ---- 1. hand out the next solves of the queue to the free lanes ------------------------------
            bool got = false;
            unsigned e = 0, p = 0;
            while (want != 0) {
                if (w_cur == w_end) {
                    unsigned q = 0;
                    if (lane == 0) q = atomicAdd(queue, 1u);
                    q = __shfl_sync(FULL_MASK, q, 0);
                    if (q >= n_items) {
                        drained = true;
                        break;
                    }
                    const unsigned run = q / (unsigned)n_ex;
                    w_e = q - run * (unsigned)n_ex;
                    w_base = run * chunk;
                    w_cur = 0;
                    w_end = (m - w_base < chunk) ? m - w_base : chunk;
                }
                const unsigned avail = w_end - w_cur;
                const unsigned rank = __popc(want & lt_mask);
                const bool mine = ((want >> lane) & 1u) && rank < avail;
                if (mine) {
                    e = w_e;
                    p = perm[w_base + w_cur + rank];
                    task = e * n + p;
                    got = true;
                }
                const unsigned n_want = __popc(want);
                w_cur += (n_want < avail) ? n_want : avail;
                want &= ~__ballot_sync(FULL_MASK, mine);
            }
            // ---- 2. set the new solves up together ------------------------------------------------
            if (got) {
                s.nVmax = -theta[p];
                s.Km = theta[ld + p];
                s.S0 = D.S0[e];
                obs = D.obs + (size_t)e * n_t;
                n_att = 0;
                n_set++;
                acc0 = n_acc;
                rej0 = n_rej;
                if (mmsolve::setup(s, D.t0[e], D.tb[e])) {
                    have = true;
                } else {
                    ssr_out[task] = INFINITY;
                    n_fail++;
                }
            }
        }
        if (__ballot_sync(FULL_MASK, have) == 0) {
            if (drained) break;
            continue;
        }
        waited += (want != 0);
        if (have) {
            const int st = mmsolve::attempt<false>(s, obs, nullptr, n_acc, n_rej);
            ++n_att;
            if (st != mmsolve::RUNNING) {
                ssr_out[task] = (st == mmsolve::DONE) ? s.ssr : INFINITY;
                if (st == mmsolve::FAILED) n_fail++;
                mx = max(mx, n_att);
                have = false;
            } else if (n_att >= budget) {
                // over budget: the solve is parked (its state goes to `park`, the result array names the slot) and the
                // tail kernel takes it from here; without a free slot it is marked DEFERRED and restarted there
                const unsigned slot = atomicAdd(queue + 4, 1u);
                if (slot < park_cap) {
                    park_store(park + (size_t)slot * PARK_WORDS, s, n_att);
                    ssr_out[task] = PARKED - (double)slot;
                } else {
                    ssr_out[task] = DEFERRED;
                    n_set--;            // redone from scratch by the tail kernel: its attempts here are dropped again
                    n_acc = acc0;
                    n_rej = rej0;
                }
                n_def++;
                have = false;
            }
        }
    }
    flush_stats(stats, n_set, n_acc, n_rej, n_fail, n_def, mx);
}

// ------------------------------------------------------------------------------ finalize
// One thread per evaluated particle, in cost order (thread i handles perm[i]).  No deferred solve: ordered
// sum -> lk.  Otherwise, unless the bound already decides it, every deferred solve goes to solve_list (for
// mm_tail_kernel), the particle to part_list (for mm_collect_kernel) and cutlim[p] becomes the residual limit
// of ONE deferred solve given what the finished ones contributed:
//     total_finished + n_deferred*c0 - ssr_e/(2 sigma^2) < lkmin.
// Lists are appended warp by warp, so they stay (block-wise) in cost order: the lanes of a tail warp hold
// solves of similar length.
template <bool BOUNDED>
__global__ void __launch_bounds__(256)
mm_finalize_kernel(const double* __restrict__ theta, int64_t ld, unsigned n, const unsigned* __restrict__ perm,
                   const double* __restrict__ lkmin, int n_ex, int n_t, const double* __restrict__ ssr,
                   double* __restrict__ lk, double* __restrict__ cutlim, unsigned* __restrict__ solve_list,
                   unsigned* __restrict__ part_list, unsigned* __restrict__ ctl,
                   unsigned long long* __restrict__ stats) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned m = ctl[3];
    const unsigned lane = threadIdx.x & 31;
    bool cut = false;
    int n_def = 0;
    unsigned p = 0;
    if (i < m) {
        p = perm[i];
        const double sigma = theta[2 * ld + p];
        const double s2 = sigma * sigma;
        const double c0 = -0.5 * n_t * log(2 * M_PI * s2);
        const double inv_den = 1.0 / (2 * s2);
        double total = 0.0;
        for (int e = 0; e < n_ex; ++e) {
            const double v = ssr[(size_t)e * n + p];
            if (v < 0) ++n_def;
            else total += c0 - v * inv_den;   // logL_i, summed in experiment order (Micmem_likelihood.py:70-73)
        }
        const double thr = BOUNDED ? lkmin[p] : -INFINITY;
        if (n_def == 0 || total == -INFINITY) {
            lk[p] = total;
            cut = BOUNDED && total == -INFINITY;
            n_def = 0;
        } else if (BOUNDED && total + n_def * c0 < thr) {
            lk[p] = -INFINITY;
            cut = true;
            n_def = 0;
        } else {
            cutlim[p] = (BOUNDED && thr > -INFINITY) ? (total + n_def * c0 - thr) * (2 * s2) : INFINITY;
        }
    }
    // warp-aggregated append: lane order = cost order
    const unsigned has = __ballot_sync(FULL_MASK, n_def > 0);
    if (has != 0) {
        int incl = n_def;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(FULL_MASK, incl, o);
            if ((int)lane >= o) incl += up;
        }
        const int warp_total = __shfl_sync(FULL_MASK, incl, 31);
        unsigned base_s = 0, base_p = 0;
        if (lane == 0) {
            base_s = atomicAdd(&ctl[1], (unsigned)warp_total);
            base_p = atomicAdd(&ctl[2], (unsigned)__popc(has));
        }
        base_s = __shfl_sync(FULL_MASK, base_s, 0);
        base_p = __shfl_sync(FULL_MASK, base_p, 0);
        if (n_def > 0) {
            part_list[base_p + __popc(has & ((1u << lane) - 1u))] = p;
            unsigned at = base_s + (unsigned)(incl - n_def);
            for (int e = 0; e < n_ex; ++e)
                if (ssr[(size_t)e * n + p] < 0) solve_list[at++] = (unsigned)e * n + p;
        }
    }
    if (BOUNDED) {
        const unsigned mk = __ballot_sync(FULL_MASK, cut);
        if (mk != 0 && lane == 0) {
            atomicAdd(&stats[8], (unsigned long long)__popc(mk));
            atomicAdd(&stats[9], (unsigned long long)__popc(mk));
        }
    }
}

// ------------------------------------------------------------------------------ tail
// One lane per deferred solve, resumed from its parked state and run to its end.  These are the 1e3 .. 1e5-step solves: each is a
// strictly serial chain (~460 cycles per attempted step for a lane that has its scheduler to itself), so the kernel
// lasts as long as its longest solve.  Everything here serves that chain:
//   * the steps are taken by mmsolve::solve_lat (mm_solver.cuh): plain steps in a loop that is ONE basic block
//     (accept / reject are selects), steps that touch an observation time through attempt();
//   * lanes of a warp are free, warps are not: inside that loop the lanes of a warp run in lockstep whatever their
//     solves do, while a second warp on the same scheduler competes for the FP64 pipe and the issue slot (the lone
//     chain runs at 456 cycles per step, next to the three other warps of a 16-block-per-SM launch at 554).  So the
//     list is dealt to as FEW warps as possible: entry k of the (roughly heaviest-first) list goes to lane
//     (k / nb) % 32 of block k % nb, where nb = one block per scheduler (4 per SM) as long as that gives a lane at
//     most TAIL_DEPTH entries to work through, and more blocks only beyond that.  The heaviest solves thus each lead
//     a different warp; the entries a lane takes after its first are the light end of the list.  Blocks >= nb exit.
//   * a lane that leaves the loop (every solve does, once per observation time) waits at the loop's exit for the
//     other lanes of its warp, so a warp advances in rounds of [plain steps until every lane needs attempt()] +
//     [one attempt()].  The stiffest lane of a warp sets the length of nearly every round (the number of steps
//     between two observation times grows with Vmax/Km), so the warp's time is that lane's own.
//   * NO warp collective, __syncthreads or value-returning atomic after the data is staged (work counters go to a
//     per-thread record that mm_collect_kernel adds up).
constexpr int TAIL_REC = 4;     // per-thread record: set-ups | failed << 32, accepted, rejected, max attempts << 32 | its cycles/attempt
constexpr unsigned TAIL_DEPTH = 4;

__global__ void __launch_bounds__(TAIL_BLOCK)
mm_tail_kernel(const double* __restrict__ theta, int64_t ld, unsigned n, const double* __restrict__ cutlim,
               const double* __restrict__ g_t, const double* __restrict__ g_P, const double* __restrict__ g_S0,
               int n_ex, int n_t, double* __restrict__ ssr, const unsigned* __restrict__ solve_list,
               const unsigned* __restrict__ ctl, const double* __restrict__ park, unsigned long long* __restrict__ rec,
               unsigned min_blocks) {
    const unsigned count = ctl[1];
    const unsigned tid = blockIdx.x * TAIL_BLOCK + threadIdx.x;
    unsigned long long* my = rec + (size_t)tid * TAIL_REC;
    // blocks in use (see above)
    unsigned nb = (count + TAIL_BLOCK * TAIL_DEPTH - 1) / (TAIL_BLOCK * TAIL_DEPTH);
    nb = max(nb, min(min_blocks, count));
    nb = min(nb, gridDim.x);
    if (blockIdx.x >= nb) {
        my[0] = 0;
        return;
    }
    extern __shared__ __align__(16) unsigned char smem[];
#if SMCB_TAIL_HOIST
    __shared__ double s_coef[mmsolve::LAT_NCOEF];
    if (threadIdx.x == 0) mmsolve::lat_coef_fill(s_coef);   // stage_data ends with a block-wide barrier
#else
    const double* s_coef = nullptr;
#endif
    const SharedData D = stage_data(smem, g_t, g_P, g_S0, n_ex, n_t);

    unsigned n_set = 0, n_acc = 0, n_rej = 0, n_fail = 0, mx = 0, mx_cyc = 0;
    const unsigned stride = nb * TAIL_BLOCK;
    for (unsigned idx = threadIdx.x * nb + blockIdx.x; idx < count; idx += stride) {
        const unsigned g = solve_list[idx];
        const unsigned e = g / n, p = g - e * n;
        Solve s;
        s.nVmax = -theta[p];
        s.Km = theta[ld + p];
        s.S0 = D.S0[e];
        s.cut_lim = cutlim[p];
        const mmsolve::ObsPair* obs = D.obs + (size_t)e * n_t;
        const unsigned att0 = n_acc + n_rej;
        const double mark = ssr[g];
        const long long c0 = clock64();
        int st = mmsolve::RUNNING;
        unsigned att_parked = 0;
        if (mark <= PARKED) {     // handed over with its state: resume
            att_parked = park_load(park + (size_t)(unsigned)(PARKED - mark) * PARK_WORDS, s, obs, D.tb[e]);
        } else {                  // restart
            n_set++;
            if (!mmsolve::setup(s, D.t0[e], D.tb[e])) st = mmsolve::FAILED;
        }
#if SMCB_TAIL_LATENCY_FORM
        if (st == mmsolve::RUNNING) st = mmsolve::solve_lat<SMCB_TAIL_HOIST != 0>(s, obs, n_acc, n_rej, s_coef);
#else
        while (st == mmsolve::RUNNING) st = mmsolve::attempt<false>(s, obs, nullptr, n_acc, n_rej);
#endif
        const long long c1 = clock64();
        ssr[g] = (st == mmsolve::DONE) ? s.ssr : INFINITY;
        if (st == mmsolve::FAILED) n_fail++;
        const unsigned att = n_acc + n_rej - att0;
        if (att + att_parked > mx) {
            mx = att + att_parked;                                   // attempts of the whole solve,
            mx_cyc = (unsigned)((c1 - c0) / (att ? att : 1u));      // cycles per attempt of the part taken here
        }
    }
    my[0] = (unsigned long long)n_set | ((unsigned long long)n_fail << 32) | (1ull << 63);   // bit 63: record in use
    my[1] = n_acc;
    my[2] = n_rej;
    my[3] = ((unsigned long long)mx << 32) | mx_cyc;
}

// ------------------------------------------------------------------------------ collect
// Ordered sum for the particles whose deferred solves the tail kernel has now finished.
template <bool BOUNDED>
__global__ void __launch_bounds__(256)
mm_collect_kernel(const double* __restrict__ theta, int64_t ld, unsigned n, int n_ex, int n_t,
                  const double* __restrict__ ssr, double* __restrict__ lk, const unsigned* __restrict__ part_list,
                  const unsigned* __restrict__ ctl, unsigned long long* __restrict__ stats,
                  const unsigned long long* __restrict__ rec, unsigned n_rec) {
    // work counters of the tail kernel's threads: summed per warp, then one atomic per counter and warp
    {
        long long fev = 0, acc = 0, rej = 0, fail = 0;
        unsigned long long mx = 0, mxrec = 0;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rec; i += gridDim.x * blockDim.x) {
            const unsigned long long* r = rec + (size_t)i * TAIL_REC;
            if (r[0] == 0) continue;
            const unsigned long long n_set = r[0] & 0xffffffffull;
            fev += (long long)(2ull * n_set + 6ull * (r[1] + r[2]));
            acc += (long long)r[1];
            rej += (long long)r[2];
            fail += (long long)((r[0] >> 32) & 0x7fffffffull);
            mx = max(mx, r[3] >> 32);
            if ((r[3] >> 32) > 1024) mxrec = max(mxrec, r[3]);
        }
        fev = warp_sum_ll(fev);
        acc = warp_sum_ll(acc);
        rej = warp_sum_ll(rej);
        fail = warp_sum_ll(fail);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, o));
            mxrec = max(mxrec, __shfl_xor_sync(FULL_MASK, mxrec, o));
        }
        if ((threadIdx.x & 31) == 0 && fev != 0) {
            atomicAdd(&stats[0], (unsigned long long)fev); atomicAdd(&stats[4], (unsigned long long)fev);
            atomicAdd(&stats[1], (unsigned long long)acc); atomicAdd(&stats[5], (unsigned long long)acc);
            atomicAdd(&stats[2], (unsigned long long)rej); atomicAdd(&stats[6], (unsigned long long)rej);
            if (fail) { atomicAdd(&stats[3], (unsigned long long)fail); atomicAdd(&stats[7], (unsigned long long)fail); }
            atomicMax(&stats[10], mx);
            atomicAdd(&stats[15], (unsigned long long)(acc + rej));
            if (mxrec) atomicMax(&stats[16], mxrec);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        stats[13] = ctl[2];
        atomicAdd(&stats[14], (unsigned long long)ctl[2]);
    }
    const unsigned count = ctl[2];
    long long n_cut = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const unsigned p = part_list[i];
        const double sigma = theta[2 * ld + p];
        const double s2 = sigma * sigma;
        const double c0 = -0.5 * n_t * log(2 * M_PI * s2);
        const double inv_den = 1.0 / (2 * s2);
        double total = 0.0;
        for (int e = 0; e < n_ex; ++e) total += c0 - ssr[(size_t)e * n + p] * inv_den;
        lk[p] = total;
        n_cut += (BOUNDED && total == -INFINITY);
    }
    if (BOUNDED) {
        n_cut = warp_sum_ll(n_cut);
        if (n_cut != 0 && (threadIdx.x & 31) == 0) {
            atomicAdd(&stats[8], (unsigned long long)n_cut);
            atomicAdd(&stats[9], (unsigned long long)n_cut);
        }
    }
}

// ------------------------------------------------------------------------------ predictions
// P_model of every (particle, experiment): pred[(p*n_ex + e)*n_t + i]  (the `C_l_` of sim_particle)
__global__ void __launch_bounds__(128)
mm_predict_kernel(const double* __restrict__ theta, int64_t ld, unsigned n, const double* __restrict__ g_t,
                  const double* __restrict__ g_P, const double* __restrict__ g_S0, int n_ex, int n_t,
                  double* __restrict__ pred) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SharedData D = stage_data(smem, g_t, g_P, g_S0, n_ex, n_t);
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * (unsigned)n_ex) return;
    const unsigned p = g / n_ex, e = g - p * n_ex;
    Solve s;
    s.nVmax = -theta[p];
    s.Km = theta[ld + p];
    s.S0 = D.S0[e];
    s.cut_lim = INFINITY;
    double* out = pred + (size_t)g * n_t;
    unsigned a = 0, r = 0;
    int st = mmsolve::setup(s, D.t0[e], D.tb[e]) ? mmsolve::RUNNING : mmsolve::FAILED;
    while (st == mmsolve::RUNNING) st = mmsolve::attempt<true>(s, D.obs + (size_t)e * n_t, out, a, r);
    if (st == mmsolve::FAILED)   // scipy would return a short solution here
        for (int i = s.i_eval; i < n_t; ++i) out[i] = NAN;
}

// ------------------------------------------------------------------------------ exact integrator (converged mode)
// the progress curve has the closed form  S(t) = Km * omega(z),
//     z(t) = ln(S0/Km) + (S0 - Vmax t)/Km,      omega + ln(omega) = z   (Wright omega = Lambert W of e^z),
// which needs no step-size control, has no stiff tail and is the CONVERGED solution of the reference's ODE
// (Micmem_likelihood.py:14-15) - not the reference's likelihood, which is defined by scipy's RK45 at rtol 1e-3 and
// differs from the converged one by up to 2.9e-3 relative.  It is therefore a separate, labelled
// integrator (SMCB_MM_EXACT) with its own oracle twin (oracle/mm.py: loglik_progress_exact) and is never used for the
// parity line.
//
// omega is found through L = ln(omega), the root of g(L) = e^L + L - z (S0/Km reaches 1e6: e^z overflows, e^L does
// not): Halley's iteration L -= 2 g g' / (2 g'^2 - g g''), g' = 1 + e^L, g'' = e^L (order 3), with the table
// exponential of exp_table.cuh (10 FP64 operations) and a MUFU-seeded reciprocal - no log, no IEEE division on the
// path.  Along an experiment the start is the previous observation time's L moved by d L/dz = 1/(1+omega)
// (second-order Taylor step; two iterations then reach 1e-15 while |dz| <= 1/2), otherwise L0 = z - omega0 (the
// identity L = z - omega) with omega0 = e^z/(1+e^z) for z < 1 and z - ln z + ln z / z (an FP32 logarithm suffices)
// above, and three iterations (checked against scipy.special.wrightomega over z in [-700, 2e6]: 4e-15 relative).
// Below z = -37 (substrate exhausted, the usual state of a fast-kinetics particle at most observation times)
// omega = e^z to the last bit and nothing is iterated.
__device__ __forceinline__ double halley_step(double& L, double z, const double* __restrict__ etab, double& E) {
    E = expt::exp_fast(L, etab);
    const double g = E + (L - z);
    const double gp = 1.0 + E;
    const double dL = -(2.0 * g * gp) * mmsolve::rcp64(fma(2.0 * gp, gp, -g * E));
    L += dL;
    return dL;
}
// omega(z) from scratch; leaves L = ln(omega)
__device__ __forceinline__ double wright_omega(double z, double& L, const double* __restrict__ etab) {
    if (z < -37.0) {                                 // substrate exhausted: omega = e^z (1 - e^z + ...), e^z < 1e-16
        const double w = expt::exp_fast(z, etab);
        L = z - w;
        return w;
    }
    double w0;
    if (z < 1.0) {
        const double e = expt::exp_fast(z, etab);
        w0 = e * mmsolve::rcp64(1.0 + e);
    } else {
        const double l = (double)__logf((float)z);
        w0 = z - l + l * mmsolve::rcp64(z);
    }
    L = z - w0;
    double E, dL;
    halley_step(L, z, etab, E);
    halley_step(L, z, etab, E);
    dL = halley_step(L, z, etab, E);
    return E * fma(dL, fma(0.5, dL, 1.0), 1.0);      // e^(L_before + dL)
}

__global__ void __launch_bounds__(128)
mm_exact_kernel(const double* __restrict__ theta, int64_t ld, int64_t n, const uint8_t* __restrict__ active,
                const double* __restrict__ g_t, const double* __restrict__ g_P, const double* __restrict__ g_S0,
                int n_ex, int n_t, const double* __restrict__ lkmin, double* __restrict__ lk, double* __restrict__ pred) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double etab[expt::TAB_N];
    expt::load_table(etab);
    const SharedData D = stage_data(smem, g_t, g_P, g_S0, n_ex, n_t);     // ends with __syncthreads()
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || (active != nullptr && !active[p])) return;
    const double Vmax = theta[p], Km = theta[ld + p], sigma = theta[2 * ld + p];
    if (pred == nullptr && !(sigma > 0)) {
        lk[p] = -INFINITY;                            // Micmem_likelihood.py:53-54
        return;
    }
    const double iKm = 1.0 / Km, k = Vmax * iKm;
    double total = 0.0;
    const double s2 = sigma * sigma;
    const double c0 = -0.5 * n_t * log(2 * M_PI * s2), inv_den = 1.0 / (2 * s2);
    // exact early rejection (smcb_loglik_bounded): residuals only accumulate, so total so far + c0 for every
    // experiment not finished - the current one's running residual term is an upper bound of the result
    const double thr = (lkmin != nullptr && pred == nullptr) ? lkmin[p] : -INFINITY;
    for (int e = 0; e < n_ex; ++e) {
        const mmsolve::ObsPair* obs = D.obs + (size_t)e * n_t;
        const double S0 = D.S0[e];
        const double y0 = S0 * iKm;
        const double L0 = log(y0);
        const double ub0 = total + (n_ex - e) * c0;
        const double z0 = L0 + y0;                    // z at t = 0: omega(z0) = y0, ln omega = L0
        double t = D.t0[e];
        double z_prev = z0 - k * t;
        double L = L0, w = y0;                        // ln(omega) and omega at z_prev
        if (t != 0.0 && Km > 0.0) w = wright_omega(z_prev, L, etab);
        double ssr = 0.0;
        for (int i = 0; i < n_t; ++i) {
            double S;
            if (!(Km > 0.0)) {
                S = fmax(S0 - Vmax * t, 0.0);         // Km -> 0: zero-order kinetics until the substrate is gone
            } else {
                if (i > 0) {
                    const double z = z0 - k * t;
                    const double dz = z - z_prev;
                    if (fabs(dz) <= 0.5) {
                        // dL/dz = 1/(1+w), d2L/dz2 = -w/(1+w)^3
                        const double a = mmsolve::rcp64(1.0 + w);
                        L = fma(dz, fma(-0.5 * dz * w, a * a * a, a), L);
                        double E;
                        halley_step(L, z, etab, E);
                        const double dL = halley_step(L, z, etab, E);
                        w = E * fma(dL, fma(0.5, dL, 1.0), 1.0);
                    } else {
                        w = wright_omega(z, L, etab);
                    }
                    z_prev = z;
                }
                S = Km * w;
            }
            const mmsolve::ObsPair o = obs[i];
            const double Pm = S0 - S;                 // Micmem_likelihood.py:32
            if (pred != nullptr) {
                pred[((size_t)p * n_ex + e) * n_t + i] = Pm;
            } else {
                const double r = o.P - Pm;
                ssr = fma(r, r, ssr);
                if (fma(-ssr, inv_den, ub0) < thr) {  // certainly below the threshold: the proposal is rejected
                    lk[p] = -INFINITY;
                    return;
                }
            }
            t = o.t_next;
        }
        total += c0 - ssr * inv_den;                  // summed in experiment order (:70-73)
    }
    if (pred == nullptr) lk[p] = total;
}

// ------------------------------------------------------------------------------ MM_RATE
// ll = -0.5*n*log(2 pi sigma^2) - sum_i (v_i - Vmax*S_i/(Km+S_i))^2 / (2 sigma^2)
// Compute-bound: 10^4 observations per particle against 32 B of particle data.  Observations are streamed
// through shared memory in tiles that every thread of the block reads by broadcast.
constexpr int RATE_BLOCK = 128;
constexpr int RATE_TILE = 2048;

// FP64: one thread per particle.  Per observation DADD, MUFU.RCP64H + three DFMA (the reciprocal correction
// folded into -Vmax*S/(Km+S), mmsolve::mm_rate), DADD, DFMA: 7 FP64-pipe operations (a true division costs ~20).
__global__ void __launch_bounds__(RATE_BLOCK)
mm_rate_kernel_f64(const double* __restrict__ theta, int64_t ld, int64_t n,
                   const uint8_t* __restrict__ active, const double* __restrict__ gS,
                   const double* __restrict__ gv, int64_t n_obs, double* __restrict__ lk) {
    __shared__ double2 sSv[RATE_TILE];
    const int64_t p = (int64_t)blockIdx.x * RATE_BLOCK + threadIdx.x;
    const bool live = p < n && (active == nullptr || active[p]);
    double nVmax = -1.0, Km = 1.0, sigma = 1.0;
    if (live) {
        nVmax = -theta[p];
        Km = theta[ld + p];
        sigma = theta[2 * ld + p];
    }
    double acc0 = 0.0, acc1 = 0.0;
    for (int64_t base = 0; base < n_obs; base += RATE_TILE) {
        const int m = (int)((n_obs - base < RATE_TILE) ? (n_obs - base) : RATE_TILE);
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += RATE_BLOCK) sSv[i] = make_double2(gS[base + i], gv[base + i]);
        __syncthreads();
        int i = 0;
#pragma unroll 2
        for (; i + 1 < m; i += 2) {
            const double2 o0 = sSv[i], o1 = sSv[i + 1];
            const double r0 = o0.y + mmsolve::mm_rate(nVmax, Km, o0.x);
            const double r1 = o1.y + mmsolve::mm_rate(nVmax, Km, o1.x);
            acc0 = fma(r0, r0, acc0);
            acc1 = fma(r1, r1, acc1);
        }
        if (i < m) {
            const double2 o0 = sSv[i];
            const double r0 = o0.y + mmsolve::mm_rate(nVmax, Km, o0.x);
            acc0 = fma(r0, r0, acc0);
        }
    }
    if (live) {
        if (sigma <= 0) {
            lk[p] = -INFINITY;
        } else {
            const double s2 = sigma * sigma;
            lk[p] = -0.5 * (double)n_obs * log(2 * M_PI * s2) - (acc0 + acc1) / (2 * s2);
        }
    }
}

// FP32 arithmetic, FP32 accumulation over 64-observation tiles, FP64 across tiles.  One thread evaluates two PAIRS
// of particles; a pair so that (a) one MUFU.RCP serves two reciprocals, 1/a = b*rcp(a*b), 1/b = a*rcp(a*b) - the kernel was
// MUFU-bound (73% of 16 rcp/clk/SM) with one reciprocal per term - and (b) the two particles form the halves of
// packed fma.rn.f32x2 / mul.f32x2 / add.f32x2 operands, which halves the issue slots (the FMA pipe retires the same
// 128 lanes/clk/SM either way: 71 vs 64 TFLOP/s measured, profiles/ubench/fp32x2.cu).  Per term and particle:
// 5.5 FMA-pipe lanes + 0.5 MUFU; shared memory holds (S, S, v, v) so one LDS.128 feeds both packed operands.
__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

constexpr int RATE_PAIRS = 2;   // packed pairs per thread: 4 particles per thread
__global__ void __launch_bounds__(RATE_BLOCK)
mm_rate_kernel_f32(const double* __restrict__ theta, int64_t ld, int64_t n,
                   const uint8_t* __restrict__ active, const float2* __restrict__ gSv, int64_t n_obs,
                   double* __restrict__ lk) {
    __shared__ float4 sObs[RATE_TILE];   // (S, S, v, v)
    // thread t of block b owns particles base + t*2 + {0,1} (pair 0) and base + 2*RATE_BLOCK + t*2 + {0,1} (pair 1):
    // one LDS.128 (512 B returned to the warp's registers, 4 clocks of the 128 B/clk path) now feeds four
    // particles; with two it was that return path, not the FMA pipe, that bound the kernel.
    const int64_t blk = (int64_t)blockIdx.x * RATE_BLOCK * 2 * RATE_PAIRS;
    int64_t pid[RATE_PAIRS][2];
    bool live[RATE_PAIRS][2];
    unsigned long long KK[RATE_PAIRS], nVV[RATE_PAIRS];
    double sg[RATE_PAIRS][2], acc[RATE_PAIRS][2];
#pragma unroll
    for (int q = 0; q < RATE_PAIRS; ++q) {
        float V[2], K[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int64_t p = blk + (int64_t)q * 2 * RATE_BLOCK + threadIdx.x * 2 + j;
            pid[q][j] = p;
            live[q][j] = p < n && (active == nullptr || active[p]);
            V[j] = 1.f; K[j] = 1.f; sg[q][j] = 1.0; acc[q][j] = 0.0;
            if (live[q][j]) {
                V[j] = (float)theta[p];
                K[j] = (float)theta[ld + p];
                sg[q][j] = theta[2 * ld + p];
            }
        }
        KK[q] = pk(K[0], K[1]);
        nVV[q] = pk(-V[0], -V[1]);
    }
    for (int64_t base = 0; base < n_obs; base += RATE_TILE) {
        const int m = (int)((n_obs - base < RATE_TILE) ? (n_obs - base) : RATE_TILE);
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += RATE_BLOCK) {
            const float2 o = gSv[base + i];
            sObs[i] = make_float4(o.x, o.x, o.y, o.y);
        }
        __syncthreads();
        for (int i0 = 0; i0 < m; i0 += 64) {
            const int i1 = (i0 + 64 < m) ? i0 + 64 : m;
            unsigned long long a[RATE_PAIRS];
#pragma unroll
            for (int q = 0; q < RATE_PAIRS; ++q) a[q] = 0ull;   // (0.f, 0.f)
#pragma unroll 4
            for (int i = i0; i < i1; ++i) {
                const float4 o = sObs[i];
                const unsigned long long SS = pk(o.x, o.y), vv = pk(o.z, o.w);
#pragma unroll
                for (int q = 0; q < RATE_PAIRS; ++q) {
                    const unsigned long long den = add2(KK[q], SS);       // (Km0+S, Km1+S)
                    float d0, d1;
                    upk(den, d0, d1);
                    float r;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d0 * d1));
                    const unsigned long long g = mul2(SS, pk(r * d1, r * d0));   // S/(Km0+S), S/(Km1+S)
                    const unsigned long long res = fma2(nVV[q], g, vv);          // v - Vmax*g
                    a[q] = fma2(res, res, a[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < RATE_PAIRS; ++q) {
                float a0, a1;
                upk(a[q], a0, a1);
                acc[q][0] += (double)a0;
                acc[q][1] += (double)a1;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < RATE_PAIRS; ++q)
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (live[q][j]) {
                const double s2 = sg[q][j] * sg[q][j];
                lk[pid[q][j]] = (sg[q][j] <= 0) ? -INFINITY
                                               : -0.5 * (double)n_obs * log(2 * M_PI * s2) - acc[q][j] / (2 * s2);
            }
}

// Sufficient-statistic form (tables built by smcb_set_data_mm_rate_sufficient): per particle two Clenshaw sums of 14 terms.
// HBM-bound: 24 B of parameters in, 8 B out.  A Km outside the tabulated range takes the direct FP64 sum.
constexpr int SUFF_BLOCK = 256;
__global__ void __launch_bounds__(SUFF_BLOCK)
mm_rate_kernel_suff(const double* __restrict__ theta, int64_t ld, int64_t n, const uint8_t* __restrict__ active,
                    const double* __restrict__ tab, double s0, double u_lo, double u_hi, double inv_log2rho,
                    double sum_v2, const double* __restrict__ gS, const double* __restrict__ gv, int64_t n_obs,
                    double* __restrict__ lk) {
    constexpr int NI = SMCB_SUFF_INT, M = SMCB_SUFF_M;
    __shared__ double sT[NI * 2 * M + 2 * NI];
    for (int k = threadIdx.x; k < NI * 2 * M + 2 * NI; k += SUFF_BLOCK) sT[k] = tab[k];
    __syncthreads();
    // grid-stride: the 14.7 KB table is staged once per block, not once per 256 particles
    for (int64_t p = (int64_t)blockIdx.x * SUFF_BLOCK + threadIdx.x; p < n; p += (int64_t)gridDim.x * SUFF_BLOCK) {
        if (active != nullptr && !active[p]) continue;
        const double Vmax = theta[p], Km = theta[ld + p], sigma = theta[2 * ld + p];
        if (!(sigma > 0)) {
            lk[p] = -INFINITY;
            continue;
        }
        const double u = Km + s0;
        double ssr;
        if (u >= u_lo && u <= u_hi) {
            int j = (int)(log2(u / u_lo) * inv_log2rho);
            j = j < 0 ? 0 : (j > NI - 1 ? NI - 1 : j);
            const double t = (u - sT[NI * 2 * M + j]) * sT[NI * 2 * M + NI + j], t2 = 2.0 * t;
            const double* ca = sT + (size_t)(j * 2) * M;
            const double* cb = ca + M;
            double a1 = 0.0, a2 = 0.0, b1 = 0.0, b2 = 0.0;   // Clenshaw
#pragma unroll
            for (int k = M - 1; k >= 1; --k) {
                const double na = fma(t2, a1, ca[k] - a2), nb = fma(t2, b1, cb[k] - b2);
                a2 = a1; a1 = na;
                b2 = b1; b1 = nb;
            }
            const double A = fma(t, a1, ca[0] - a2), B = fma(t, b1, cb[0] - b2);
            ssr = fma(Vmax, fma(Vmax, B, -2.0 * A), sum_v2);   // sum v^2 - 2 Vmax A + Vmax^2 B
        } else {
            double acc = 0.0;
            for (int64_t i = 0; i < n_obs; ++i) {
                const double r = gv[i] + mmsolve::mm_rate(-Vmax, Km, gS[i]);
                acc = fma(r, r, acc);
            }
            ssr = acc;
        }
        const double s2 = sigma * sigma;
        lk[p] = -0.5 * (double)n_obs * log(2 * M_PI * s2) - ssr / (2 * s2);
    }
}

}  // namespace

// Optional per-kernel timing (SMCB_PARAM_PROFILE): CUDA events on the launching stream around the bulk and the
// tail kernel of every sweep, read back by smcb_profile_read.
static inline void prof_mark(smcb_handle* h, int which, cudaStream_t st) {
    if (!h->prof_on) return;
    cudaEventRecord(h->prof_ev[(size_t)h->prof_sweeps * 4 + which], st);   // room made by prof_begin_sweep
}

// sweep-local counters, the solve queue and the deferred-particle list head
__global__ void mm_reset_kernel(unsigned long long* stats, unsigned* ctl, unsigned* hist) {
    if (threadIdx.x < 4) stats[threadIdx.x] = 0;
    if (threadIdx.x == 8 || threadIdx.x == 10 || threadIdx.x == 11 || threadIdx.x == 13) stats[threadIdx.x] = 0;
    if (threadIdx.x < 8) ctl[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < NBIN; i += blockDim.x) hist[i] = 0;
}

int launch_loglik_mm_progress(smcb_handle* h, const double* theta, int64_t ld, int64_t n,
                              const uint8_t* active, const double* lkmin, double* lk, double* pred,
                              cudaStream_t st) {
    const MmProgressData& D = h->mmp;
    REQUIRE(h, D.t != nullptr, SMCB_ERR_STATE, "smcb_set_data_mm_progress has not been called");
    if (n == 0) return SMCB_OK;
    const size_t smem = shared_data_bytes(D.n_ex, D.n_t);
    REQUIRE(h, smem <= 200 * 1024, SMCB_ERR_UNSUPPORTED, "data set too large for shared-memory staging");
    REQUIRE(h, n * (int64_t)D.n_ex < (1LL << 31), SMCB_ERR_UNSUPPORTED, "too many solves for one launch");
    const unsigned un = (unsigned)n;
    if (h->mm_integrator == SMCB_MM_EXACT) {          // closed-form progress curves: no ordering, no tail
        if (smem > 48 * 1024)
            CUDA_TRY(h, cudaFuncSetAttribute(mm_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mm_exact_kernel<<<(un + 127) / 128, 128, smem, st>>>(theta, ld, n, active, D.t, D.P, D.S0, D.n_ex, D.n_t, lkmin, lk,
                                                            pred);
        LAUNCH_CHECK(h);
        return SMCB_OK;
    }
    if (pred != nullptr) {
        const unsigned tasks = un * (unsigned)D.n_ex;
        if (smem > 48 * 1024)
            CUDA_TRY(h, cudaFuncSetAttribute(mm_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mm_predict_kernel<<<(tasks + 127) / 128, 128, smem, st>>>(theta, ld, un, D.t, D.P, D.S0, D.n_ex, D.n_t, pred);
        LAUNCH_CHECK(h);
        return SMCB_OK;
    }
    REQUIRE(h, h->n_max >= n && h->ssr_rows >= D.n_ex && h->mm_defer != nullptr, SMCB_ERR_STATE,
            "smcb_reserve(n_max >= n) must be called before smcb_loglik");
    const bool bounded = lkmin != nullptr;
    if (smem > 48 * 1024 && !h->mm_smem_set) {
        CUDA_TRY(h, cudaFuncSetAttribute(mm_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(h, cudaFuncSetAttribute(mm_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->mm_smem_set = true;
    }
    if (h->mm_bulk_blocks_per_sm == 0) {
        int a = 0;
        CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, mm_bulk_kernel, BULK_BLOCK, smem));
        h->mm_bulk_blocks_per_sm = a > 0 ? a : 1;
        CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, mm_tail_kernel, TAIL_BLOCK, smem));
        h->mm_tail_blocks_per_sm = a > 0 ? a : 1;
    }
    {
        const int rc = prof_begin_sweep(h);
        if (rc) return rc;
    }
    unsigned* queue = h->mm_ctl;   // [0] solve queue head, [1] deferred solves, [2] deferred particles, [3] particles to evaluate
    unsigned* hist = h->mm_hist;   // [NBIN] histogram, then [NBIN] scatter cursors
    mm_reset_kernel<<<1, 128, 0, st>>>(h->stats, h->mm_ctl, hist);
    LAUNCH_CHECK(h);
    if (bounded)
        mm_bin_kernel<true><<<(un + 255) / 256, 256, 0, st>>>(theta, ld, un, active, lkmin, D.n_ex, D.n_t, h->mm_cutlim,
                                                             h->mm_bins, hist, lk, h->stats);
    else
        mm_bin_kernel<false><<<(un + 255) / 256, 256, 0, st>>>(theta, ld, un, active, nullptr, D.n_ex, D.n_t, nullptr,
                                                              h->mm_bins, hist, lk, h->stats);
    LAUNCH_CHECK(h);
    mm_binscan_kernel<<<1, NBIN, 0, st>>>(hist, hist + NBIN, h->mm_ctl);
    LAUNCH_CHECK(h);
    mm_scatter_kernel<<<(un + 255) / 256, 256, 0, st>>>(un, h->mm_bins, hist + NBIN, h->mm_perm);
    LAUNCH_CHECK(h);
    const unsigned tasks = un * (unsigned)D.n_ex;
    unsigned grid = (unsigned)(h->sm_count * h->mm_bulk_blocks_per_sm);
    const unsigned need = (tasks + BULK_BLOCK - 1) / BULK_BLOCK;
    if (need < grid) grid = need;
    const unsigned budget = (unsigned)h->mm_budget;
    prof_mark(h, 0, st);
    mm_bulk_kernel<<<grid, BULK_BLOCK, smem, st>>>(theta, ld, un, h->mm_perm, D.t, D.P, D.S0, D.n_ex, D.n_t, budget,
                                                  h->mm_refill_min, h->mm_patience, (unsigned)h->mm_chunk, h->ssr, queue,
                                                  h->mm_park, h->mm_park_cap, h->stats);
    LAUNCH_CHECK(h);
    prof_mark(h, 1, st);
    unsigned* solve_list = h->mm_defer;                              // [n_ex * n_max]
    unsigned* part_list = h->mm_defer + (size_t)h->ssr_rows * h->n_max;   // [n_max]
    if (bounded)
        mm_finalize_kernel<true><<<(un + 255) / 256, 256, 0, st>>>(theta, ld, un, h->mm_perm, lkmin, D.n_ex, D.n_t, h->ssr,
                                                                  lk, h->mm_cutlim, solve_list, part_list, h->mm_ctl,
                                                                  h->stats);
    else
        mm_finalize_kernel<false><<<(un + 255) / 256, 256, 0, st>>>(theta, ld, un, h->mm_perm, nullptr, D.n_ex, D.n_t,
                                                                   h->ssr, lk, h->mm_cutlim, solve_list, part_list,
                                                                   h->mm_ctl, h->stats);
    LAUNCH_CHECK(h);
    // One-warp blocks.  The grid is what is resident together (at most mm_tail_warps blocks per SM); the kernel itself
    // decides from the length of the list how many of those blocks it uses (one per scheduler unless the list is very
    // long) and walks the list with a stride, so one launch covers any list.
    const unsigned tail_grid = (unsigned)h->sm_count * (unsigned)std::min(h->mm_tail_warps, h->mm_tail_blocks_per_sm);
    prof_mark(h, 2, st);
    mm_tail_kernel<<<tail_grid, TAIL_BLOCK, smem, st>>>(theta, ld, un, h->mm_cutlim, D.t, D.P, D.S0, D.n_ex, D.n_t, h->ssr,
                                                       solve_list, h->mm_ctl, h->mm_park, h->mm_tailrec,
                                                       4u * (unsigned)h->sm_count);
    LAUNCH_CHECK(h);
    prof_mark(h, 3, st);
    const unsigned cgrid = (unsigned)h->sm_count * 8;
    if (bounded)
        mm_collect_kernel<true><<<cgrid, 256, 0, st>>>(theta, ld, un, D.n_ex, D.n_t, h->ssr, lk, part_list, h->mm_ctl,
                                                      h->stats, h->mm_tailrec, tail_grid * TAIL_BLOCK);
    else
        mm_collect_kernel<false><<<cgrid, 256, 0, st>>>(theta, ld, un, D.n_ex, D.n_t, h->ssr, lk, part_list, h->mm_ctl,
                                                       h->stats, h->mm_tailrec, tail_grid * TAIL_BLOCK);
    LAUNCH_CHECK(h);
    if (h->prof_on) h->prof_sweeps++;
    return SMCB_OK;
}

int launch_loglik_mm_rate(smcb_handle* h, const double* theta, int64_t ld, int64_t n, const uint8_t* active,
                          double* lk, cudaStream_t st) {
    const MmRateData& D = h->mmr;
    REQUIRE(h, D.S != nullptr, SMCB_ERR_STATE, "smcb_set_data_mm_rate has not been called");
    if (n == 0) return SMCB_OK;
    const int64_t grid = (n + RATE_BLOCK - 1) / RATE_BLOCK;
    const int64_t per_block = (int64_t)RATE_BLOCK * 2 * RATE_PAIRS;
    if (D.precision == 0)
        mm_rate_kernel_suff<<<(unsigned)std::min<int64_t>((n + SUFF_BLOCK - 1) / SUFF_BLOCK, (int64_t)h->sm_count * 8),
                              SUFF_BLOCK, 0, st>>>(
            theta, ld, n, active, D.suff, D.suff_s0, D.suff_ulo, D.suff_uhi, D.suff_inv_log2rho, D.sum_v2, D.S, D.v,
            D.n_obs, lk);
    else if (D.precision == 32)   // four particles per thread
        mm_rate_kernel_f32<<<(unsigned)((n + per_block - 1) / per_block), RATE_BLOCK, 0, st>>>(
            theta, ld, n, active, D.Sv32, D.n_obs, lk);
    else
        mm_rate_kernel_f64<<<(unsigned)grid, RATE_BLOCK, 0, st>>>(theta, ld, n, active, D.S, D.v, D.n_obs, lk);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}
