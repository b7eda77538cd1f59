// K3: residual-systematic resampling.  Replaces the serial Python loop of the reference
// (SMC_example/Micmem_SMC_main.py:147-184):
//
//     p_is = trunc(w*N);  w -= p_is*inv_Np;  wrand = u0*inv_Np
//     for j: sum += w[j];  if sum >= wrand: p_is[j]++, wrand += inv_Np;  emit p_is[j] copies of j
//
// Two prefix-sum arithmetics:
//   SEQUENTIAL  one warp walks the residuals with the reference's sequentially rounded FP64
//               running sum and sequentially incremented threshold: bit-exact ancestors.
//   FIXED       residuals are quantised to 2^-62 fixed point and summed exactly in integers with
//               a blocked parallel scan; the number of thresholds (u0+k)/N at or below a prefix s
//               is cross(s) = floor((s*N - u0q)/2^62)+1, evaluated in 128-bit integers.  The
//               result is independent of block structure and of how particles are sharded.
// Copy counts are expanded to the (non-decreasing) ancestor vector by marking the first slot of
// each surviving particle and running a max-scan; particle state then moves with one gather.
#include "common.cuh"

namespace {

constexpr int SB = 256;               // scan block
constexpr int IPT = 8;                // items per thread
constexpr int TILE = SB * IPT;        // 2048
constexpr double TWO62 = 4611686018427387904.0;

// ---------------------------------------------------------------------------------------------
// block-wide exclusive scan of one value per thread (sum for uint64/int64, max for int32)
struct OpAdd {
    template <typename T>
    __device__ __forceinline__ T operator()(T a, T b) const { return a + b; }
};
struct OpMax {
    template <typename T>
    __device__ __forceinline__ T operator()(T a, T b) const { return a > b ? a : b; }
};

template <typename T, typename Op>
__device__ __forceinline__ T block_exclusive_scan(T v, T identity, Op op, T* smem /*[32]*/, T* block_total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T up = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl = op(up, incl);
    }
    if (lane == 31) smem[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        T w = (lane < SB / 32) ? smem[lane] : identity;
        T wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T up = __shfl_up_sync(FULL_MASK, wi, o);
            if (lane >= o) wi = op(up, wi);
        }
        if (lane < SB / 32) smem[lane] = wi;   // inclusive over warps
    }
    __syncthreads();
    T excl = __shfl_up_sync(FULL_MASK, incl, 1);
    if (lane == 0) excl = identity;
    if (wid > 0) excl = op(smem[wid - 1], excl);
    *block_total = smem[SB / 32 - 1];
    __syncthreads();
    return excl;
}

// ---------------------------------------------------------------------------------------------
// prepare: floor counts and residuals from normalised weights; per-tile totals
template <int MODE>
__global__ void __launch_bounds__(SB)
prepare_kernel(const double* __restrict__ w, int64_t n, double Nd, double inv_Np,
               int32_t* __restrict__ floor_cnt, double* __restrict__ resid_f,
               uint64_t* __restrict__ resid_q, int64_t* __restrict__ tile_tot) {
    __shared__ long long sm_f[32];
    __shared__ long long sm_q[32];
    const int64_t base = (int64_t)blockIdx.x * TILE;
    long long tf = 0, tq = 0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const int64_t j = base + (int64_t)k * SB + threadIdx.x;   // coalesced
        if (j < n) {
            const double wj = w[j];
            const double fl = trunc(__dmul_rn(wj, Nd));              // np.trunc(p_weight*n_particle)
            const double r = __dsub_rn(wj, __dmul_rn(fl, inv_Np));   // p_weight - p_is*inv_Np
            const int32_t c = (int32_t)fl;
            floor_cnt[j] = c;
            tf += c;
            if (MODE == SMCB_SCAN_SEQUENTIAL) {
                resid_f[j] = r;
            } else {
                double rq = r * TWO62;
                rq = (rq > 0.0) ? rq : 0.0;
                const uint64_t q = (uint64_t)__double2ull_rn(rq);
                resid_q[j] = q;
                tq += (long long)q;
            }
        }
    }
    tf = warp_sum_ll(tf);
    tq = warp_sum_ll(tq);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        sm_f[wid] = tf;
        sm_q[wid] = tq;
    }
    __syncthreads();
    if (wid == 0) {
        tf = (lane < SB / 32) ? sm_f[lane] : 0;
        tq = (lane < SB / 32) ? sm_q[lane] : 0;
        tf = warp_sum_ll(tf);
        tq = warp_sum_ll(tq);
        if (lane == 0) {
            tile_tot[2 * (int64_t)blockIdx.x] = tf;
            tile_tot[2 * (int64_t)blockIdx.x + 1] = tq;
        }
    }
}

// exclusive scan over tile totals (pairs), single block; totals written to totals_out[0..1].
__global__ void __launch_bounds__(SB)
scan_tiles_pair_kernel(int64_t* __restrict__ tile_tot, int64_t n_tiles, int64_t carry0, int64_t carry1,
                       int64_t* __restrict__ totals_out) {
    __shared__ long long sm[32];
    long long run0 = carry0, run1 = carry1;
    for (int64_t b0 = 0; b0 < n_tiles; b0 += SB) {
        const int64_t i = b0 + threadIdx.x;
        long long v0 = 0, v1 = 0;
        if (i < n_tiles) {
            v0 = tile_tot[2 * i];
            v1 = tile_tot[2 * i + 1];
        }
        long long t0, t1;
        long long e0 = block_exclusive_scan<long long>(v0, 0LL, OpAdd(), sm, &t0);
        long long e1 = block_exclusive_scan<long long>(v1, 0LL, OpAdd(), sm, &t1);
        if (i < n_tiles) {
            tile_tot[2 * i] = run0 + e0;
            tile_tot[2 * i + 1] = run1 + e1;
        }
        run0 += t0;
        run1 += t1;
    }
    if (threadIdx.x == 0 && totals_out != nullptr) {
        totals_out[0] = run0 - carry0;
        totals_out[1] = run1 - carry1;
    }
}

__global__ void totals_from_tiles_kernel(const int64_t* __restrict__ tile_tot, int64_t n_tiles,
                                         int64_t* __restrict__ totals_out) {
    // single thread block, tiny: plain ordered sum of the *unscanned* tile totals
    __shared__ long long sm[32];
    long long a = 0, b = 0;
    for (int64_t i = threadIdx.x; i < n_tiles; i += blockDim.x) {
        a += tile_tot[2 * i];
        b += tile_tot[2 * i + 1];
    }
    a = warp_sum_ll(a);
    b = warp_sum_ll(b);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = a;
    __syncthreads();
    long long ta = 0;
    if (threadIdx.x == 0)
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) ta += sm[k];
    __syncthreads();
    if (lane == 0) sm[wid] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long tb = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tb += sm[k];
        totals_out[0] = ta;
        totals_out[1] = tb;
    }
}

// number of thresholds (u0 + k)/N, k>=0, at or below the fixed-point prefix s
__device__ __forceinline__ long long crossings(uint64_t s, uint64_t N, uint64_t u0q) {
    const uint64_t lo = s * N;
    const uint64_t hi = __umul64hi(s, N);
    if (hi == 0 && lo < u0q) return 0;
    const uint64_t dlo = lo - u0q;
    const uint64_t dhi = hi - (lo < u0q ? 1 : 0);
    return (long long)((dhi << 2) | (dlo >> 62)) + 1;
}

// FIXED mode: counts[j] = floor[j] + cross(s_j) - cross(s_{j-1}),  s = exact prefix of q
__global__ void __launch_bounds__(SB)
counts_fixed_kernel(const int32_t* __restrict__ floor_cnt, const uint64_t* __restrict__ resid_q, int64_t n,
                    const int64_t* __restrict__ tile_base, uint64_t N, uint64_t u0q, int first_shard,
                    int32_t* __restrict__ counts) {
    __shared__ unsigned long long sm[32];
    const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * IPT;   // blocked layout
    unsigned long long q[IPT];
    unsigned long long tsum = 0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        q[k] = (base + k < n) ? resid_q[base + k] : 0ULL;
        tsum += q[k];
    }
    unsigned long long btot;
    unsigned long long excl = block_exclusive_scan<unsigned long long>(tsum, 0ULL, OpAdd(), sm, &btot);
    unsigned long long s = (unsigned long long)tile_base[2 * (int64_t)blockIdx.x + 1] + excl;
    long long c_prev = crossings(s, N, u0q);
    if (first_shard && base == 0) c_prev = 0;   // nothing is crossed before the first particle of the run
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        s += q[k];
        const long long c = crossings(s, N, u0q);
        if (base + k < n) counts[base + k] = floor_cnt[base + k] + (int32_t)(c - c_prev);
        c_prev = c;
    }
}

// SEQUENTIAL mode: one warp; lane 0 carries the reference's running sum and threshold.
constexpr int SEQ_CHUNK = 1024;
__global__ void __launch_bounds__(32)
counts_sequential_kernel(const int32_t* __restrict__ floor_cnt, const double* __restrict__ resid_f, int64_t n,
                         double inv_Np, double* __restrict__ carry /*[2]: sum, wrand*/,
                         int32_t* __restrict__ counts, int64_t* __restrict__ totals_out) {
    __shared__ double s_w[SEQ_CHUNK];
    __shared__ unsigned char s_c[SEQ_CHUNK];
    const int lane = threadIdx.x;
    double run = carry[0], wrand = carry[1];
    long long ncross = 0, nfloor = 0;
    for (int64_t b0 = 0; b0 < n; b0 += SEQ_CHUNK) {
        const int m = (int)((n - b0 < SEQ_CHUNK) ? (n - b0) : SEQ_CHUNK);
        for (int i = lane; i < m; i += 32) s_w[i] = resid_f[b0 + i];
        __syncwarp();
        if (lane == 0) {
#pragma unroll 8
            for (int i = 0; i < m; ++i) {
                run = __dadd_rn(run, s_w[i]);
                const bool hit = run >= wrand;
                s_c[i] = hit ? 1 : 0;
                if (hit) {
                    wrand = __dadd_rn(wrand, inv_Np);
                    ++ncross;
                }
            }
        }
        __syncwarp();
        for (int i = lane; i < m; i += 32) {
            const int32_t f = floor_cnt[b0 + i];
            counts[b0 + i] = f + s_c[i];
            nfloor += f;
        }
        __syncwarp();
    }
    nfloor = warp_sum_ll(nfloor);
    if (lane == 0) {
        carry[0] = run;
        carry[1] = wrand;
        totals_out[0] = nfloor;
        totals_out[1] = ncross;
    }
}

// SEQUENTIAL mode at block speed.  The reference's running sum S_j = fl(S_{j-1} + r_j) and its threshold
// T_k = fl(T_{k-1} + 1/N) look inherently serial, but inside one binade [2^e, 2^(e+1)) every FP64 value is an integer
// multiple of u = 2^(e-52), so fl(S + r) = S + rn(r/u)*u as long as the result stays in the binade and r/u is not
// exactly half-way between two integers: the sequentially ROUNDED sum of a chunk is an exact INTEGER prefix sum of
// q_j = rn(r_j/u), which a block scans in parallel; likewise the thresholds of a chunk are T + i*rn((1/N)/u_T)*u_T,
// and the number of thresholds at or below S_j is one integer division.  One block of 1024 threads walks the
// particles in chunks of 1024, carrying (S, T) exactly; a chunk that breaks an assumption (a binade boundary of S
// or T inside it, a half-way case, S still tiny against the residuals, a particle crossing two thresholds) is done
// by thread 0 with the literal loop, 2.4% of the chunks at 2^20 particles.  Result: the reference's counts bit for
// bit (tests/test_gpu_kernels.py, oracle/smc.py::resample_sequential) in 2.65 ms instead of 4.4 ms at 2^20 and 8.6
// instead of 17.6 ms at 2^22 (profiles/resample_sequential_timing.py): the chain over the chunks keeps the work on
// ONE SM, whose issue rate (~150 instructions per particle) is now the limit, not the latency of 2^20 dependent adds.
constexpr int SQB = 1024;
__device__ __forceinline__ double pow2i(int k) { return __hiloint2double((k + 1023) << 20, 0); }   // -1022 <= k <= 1023
__device__ __forceinline__ int unit_exp(double x) { return ((__double2hiint(x) >> 20) & 0x7ff) - 1023 - 52; }

__global__ void __launch_bounds__(SQB)
counts_sequential_block_kernel(const int32_t* __restrict__ floor_cnt, const double* __restrict__ resid_f, int64_t n,
                               double inv_Np, double* __restrict__ carry /*[2]: sum, wrand*/,
                               int32_t* __restrict__ counts, int64_t* __restrict__ totals_out,
                               unsigned long long* __restrict__ dbg /*[2] chunks, literal chunks; may be null*/) {
    __shared__ double s_r[SQB];
    __shared__ unsigned long long s_c[SQB];
    __shared__ long long s_w[32];
    __shared__ double sh_S, sh_T;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        sh_S = carry[0];
        sh_T = carry[1];
    }
    long long nfloor = 0, ncross = 0, n_lit = 0, n_chunks = 0;   // ncross, n_lit, n_chunks: thread 0 / uniform
    __syncthreads();
    // the chunks depend on one another through (S, T), so nothing hides a load but the previous chunk's work: the
    // next chunk's residuals and floor counts are fetched while the current one is scanned
    double r_next = (tid < n) ? resid_f[tid] : 0.0;
    int32_t f_next = (tid < n) ? floor_cnt[tid] : 0;
    for (int64_t b0 = 0; b0 < n; b0 += SQB) {
        const int m = (int)((n - b0 < SQB) ? (n - b0) : SQB);
        const double r = r_next;
        const int32_t f_cur = f_next;
        {
            const int64_t jn = b0 + SQB + tid;
            r_next = (jn < n) ? resid_f[jn] : 0.0;
            f_next = (jn < n) ? floor_cnt[jn] : 0;
        }
        s_r[tid] = r;
        const double S = sh_S, T = sh_T;
        ++n_chunks;
        // ---- can the chunk be scanned in integers? (every condition is checked, none is assumed) ----
        bool bad = !(S > 0.0 && T > 0.0 && S < 1e300 && T < 1e300 && S > 1e-280 && T > 1e-280);
        const int uS = bad ? 0 : unit_exp(S), uT = bad ? 0 : unit_exp(T);
        const int emin = uS < uT ? uS : uT;
        const int dS = uS - emin, dT = uT - emin;
        bad = bad || (dS > 10) || (dT > 10);
        bad = bad || !(fabs(r) < S * 1024.0);                    // also catches NaN
        double t = 0.0;
        long long q = 0;
        if (!bad) {
            t = r * pow2i(-uS);                                  // exact scaling: |t| < 2^63
            bad = (t - floor(t) == 0.5);                         // half-way: the rounding would depend on the sum's parity
            q = __double2ll_rn(t);
        }
        // thresholds: T + i*step*u_T
        long long step = 0, Mt = 0, M = 0;
        if (!bad) {
            const double st = inv_Np * pow2i(-uT);
            bad = !(st >= 1.0 && st < 4.0e18) || (st - floor(st) == 0.5);
            step = __double2ll_rn(st);
            Mt = (long long)(T * pow2i(-uT));                    // exact: an integer in [2^52, 2^53)
            M = (long long)(S * pow2i(-uS));
        }
        // inclusive block scan of q
        long long incl = q;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long up = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) s_w[wid] = incl;
        int any_bad = __syncthreads_or(bad ? 1 : 0);
        if (wid == 0) {
            long long w = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long up = __shfl_up_sync(FULL_MASK, w, o);
                if (lane >= o) w += up;
            }
            s_w[lane] = w;
        }
        __syncthreads();
        const long long cum = incl + (wid > 0 ? s_w[wid - 1] : 0);
        const long long total = s_w[31];
        unsigned long long c = 0;
        const double inv_sc = (!any_bad) ? 1.0 / (double)((unsigned long long)step << dT) : 0.0;
        if (!any_bad) {
            const long long Mj = M + cum;
            bad = (Mj < (1LL << 52)) || (M + total >= (1LL << 53));       // S leaves its binade inside the chunk
            if (!bad) {
                const unsigned long long A = (unsigned long long)Mj << dS, B = (unsigned long long)Mt << dT,
                                         sc = (unsigned long long)step << dT;
                // thresholds T + i*step at or below S_j: floor((A-B)/sc) + 1, the quotient (<= 1024 here) from a
                // double estimate corrected by exact integer products (a 64-bit division costs ~100 instructions)
                if (A >= B) {
                    const unsigned long long D = A - B;
                    unsigned long long qd = (unsigned long long)((double)D * inv_sc);
                    if (qd * sc > D) --qd;
                    if ((qd + 1ULL) * sc <= D) ++qd;
                    bad = bad || (qd * sc > D) || ((qd + 1ULL) * sc <= D) || qd > 4096ULL;   // estimate off by more than one
                    c = qd + 1ULL;
                }
            }
        }
        s_c[tid] = c;
        any_bad = __syncthreads_or((any_bad || bad) ? 1 : 0);
        int flag = 0;
        if (!any_bad) {
            const unsigned long long c_prev = (tid > 0) ? s_c[tid - 1] : 0ULL;
            bad = (c < c_prev) || (c - c_prev > 1ULL);                     // the literal loop takes one threshold per particle
            flag = (int)(c - c_prev);
            // every threshold that was compared with (index <= c_last) must lie in T's binade
            if (tid == SQB - 1) bad = bad || (Mt + (long long)(c + 1) * step >= (1LL << 53));
        }
        any_bad = __syncthreads_or((any_bad || bad) ? 1 : 0);
        if (any_bad) {
            // ---- literal loop of the reference for this chunk (Micmem_SMC_main.py:165-174) ----
            if (tid == 0) {
                double run = S, wrand = T;
#pragma unroll 8
                for (int i = 0; i < m; ++i) {
                    run = __dadd_rn(run, s_r[i]);
                    const bool hit = run >= wrand;
                    s_c[i] = hit ? 1ULL : 0ULL;
                    if (hit) {
                        wrand = __dadd_rn(wrand, inv_Np);
                        ++ncross;
                    }
                }
                sh_S = run;
                sh_T = wrand;
                ++n_lit;
            }
            __syncthreads();
            flag = (tid < m) ? (int)s_c[tid] : 0;
        } else if (tid == SQB - 1) {
            sh_S = (double)(M + total) * pow2i(uS);                        // exact: below 2^53 in units of u_S
            sh_T = (double)(Mt + (long long)c * step) * pow2i(uT);
            s_w[0] = (long long)c;
        }
        __syncthreads();
        if (!any_bad && tid == 0) ncross += s_w[0];
        if (tid < m) {
            counts[b0 + tid] = f_cur + flag;
            nfloor += f_cur;
        }
        __syncthreads();
    }
    nfloor = warp_sum_ll(nfloor);
    if (lane == 0) s_w[wid] = nfloor;
    __syncthreads();
    if (tid == 0) {
        long long tf = 0;
        for (int k = 0; k < SQB / 32; ++k) tf += s_w[k];
        carry[0] = sh_S;
        carry[1] = sh_T;
        totals_out[0] = tf;
        totals_out[1] = ncross;
        if (dbg != nullptr) {
            dbg[0] = (unsigned long long)n_chunks;
            dbg[1] = (unsigned long long)n_lit;
        }

    }
}

// ---------------------------------------------------------------------------------------------
// ancestor expansion
__global__ void __launch_bounds__(SB)
count_tiles_kernel(const int32_t* __restrict__ counts, int64_t n, int64_t* __restrict__ tile_tot) {
    __shared__ long long sm[32];
    const int64_t base = (int64_t)blockIdx.x * TILE;
    long long t = 0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const int64_t j = base + (int64_t)k * SB + threadIdx.x;
        if (j < n) t += counts[j];
    }
    t = warp_sum_ll(t);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = t;
    __syncthreads();
    if (wid == 0) {
        t = (lane < SB / 32) ? sm[lane] : 0;
        t = warp_sum_ll(t);
        if (lane == 0) {
            tile_tot[2 * (int64_t)blockIdx.x] = t;
            tile_tot[2 * (int64_t)blockIdx.x + 1] = 0;
        }
    }
}

// mark[offset_j] = j+1 for every particle with count_j>0 and offset_j<m
__global__ void __launch_bounds__(SB)
mark_heads_kernel(const int32_t* __restrict__ counts, int64_t n, const int64_t* __restrict__ tile_base,
                  int64_t m, int32_t* __restrict__ mark) {
    __shared__ long long sm[32];
    const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * IPT;
    int32_t c[IPT];
    long long tsum = 0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        c[k] = (base + k < n) ? counts[base + k] : 0;
        tsum += c[k];
    }
    long long btot;
    long long off = tile_base[2 * (int64_t)blockIdx.x] +
                    block_exclusive_scan<long long>(tsum, 0LL, OpAdd(), sm, &btot);
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        if (c[k] > 0 && off < m) mark[off] = (int32_t)(base + k) + 1;
        off += c[k];
    }
}

__global__ void __launch_bounds__(SB)
max_tiles_kernel(const int32_t* __restrict__ mark, int64_t m, int32_t* __restrict__ tile_max) {
    __shared__ int sm[32];
    const int64_t base = (int64_t)blockIdx.x * TILE;
    int t = 0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const int64_t j = base + (int64_t)k * SB + threadIdx.x;
        if (j < m) t = max(t, mark[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = max(t, __shfl_xor_sync(FULL_MASK, t, o));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = t;
    __syncthreads();
    if (wid == 0) {
        t = (lane < SB / 32) ? sm[lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = max(t, __shfl_xor_sync(FULL_MASK, t, o));
        if (lane == 0) tile_max[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(SB) scan_tiles_max_kernel(int32_t* __restrict__ tile_max, int64_t n_tiles) {
    __shared__ int sm[32];
    int run = 0;
    for (int64_t b0 = 0; b0 < n_tiles; b0 += SB) {
        const int64_t i = b0 + threadIdx.x;
        int v = (i < n_tiles) ? tile_max[i] : 0;
        int tot;
        int e = block_exclusive_scan<int>(v, 0, OpMax(), sm, &tot);
        if (i < n_tiles) tile_max[i] = max(run, e);
        run = max(run, tot);
    }
}

__global__ void __launch_bounds__(SB)
fill_ancestors_kernel(const int32_t* __restrict__ mark, int64_t m, const int32_t* __restrict__ tile_max,
                      int32_t* __restrict__ anc) {
    __shared__ int sm[32];
    const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * IPT;
    int v[IPT];
    int tmax = 0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        v[k] = (base + k < m) ? mark[base + k] : 0;
        tmax = max(tmax, v[k]);
    }
    int btot;
    int run = max(tile_max[blockIdx.x], block_exclusive_scan<int>(tmax, 0, OpMax(), sm, &btot));
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        run = max(run, v[k]);
        if (base + k < m) anc[base + k] = (run > 0 ? run : 1) - 1;
    }
}

// dst[k][s] = src[k][anc[s]]
__global__ void __launch_bounds__(256)
gather_kernel(const double* __restrict__ src, int64_t ld_src, const int32_t* __restrict__ anc, int64_t m, int rows,
              double* __restrict__ dst, int64_t ld_dst) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m) return;
    const int64_t a = anc[s];
    int k = 0;
    for (; k + 4 <= rows; k += 4) {
        const double v0 = src[(int64_t)(k + 0) * ld_src + a];
        const double v1 = src[(int64_t)(k + 1) * ld_src + a];
        const double v2 = src[(int64_t)(k + 2) * ld_src + a];
        const double v3 = src[(int64_t)(k + 3) * ld_src + a];
        dst[(int64_t)(k + 0) * ld_dst + s] = v0;
        dst[(int64_t)(k + 1) * ld_dst + s] = v1;
        dst[(int64_t)(k + 2) * ld_dst + s] = v2;
        dst[(int64_t)(k + 3) * ld_dst + s] = v3;
    }
    for (; k < rows; ++k) dst[(int64_t)k * ld_dst + s] = src[(int64_t)k * ld_src + a];
}

// ---------------------------------------------------------------------------------------------
// Single-pass resampling (one shard, FIXED arithmetic): weights -> floor counts / fixed-point residuals -> copy
// counts -> output offsets -> ancestors -> gather of the particle state, in ONE kernel.
//
// The copies placed before particle j number  sum_{i<j} floor_i + cross(s_{j-1})  (the crossing counts telescope),
// so two global prefixes suffice: the floor counts and the exact fixed-point residual sums.  Tiles of 2048 particles
// take their prefixes from a decoupled look-back chain (tile ids handed out by an atomic counter, so a tile only
// ever waits for tiles that are already running); both prefixes fit one 64-bit word each with a 2-bit status on top
// (residual prefix < 2^62 because the residuals of a normalised weight vector sum to < 1; floor prefix < 2^31).
// A tile then expands its own slice of the output cooperatively: output slot s finds its particle by a binary search
// over the tile's 2048 inclusive end offsets in shared memory, which costs the same for one particle with 10^5
// copies as for 10^5 particles with one, and moves the particle's rows at once.  HBM traffic per particle:
// 8 B (log-likelihood) + 4 B (ancestor) + 2 (d+1) 8 B (state), against 8 + 4 + 4 + 2 (d+1) 8 algorithmic
//  and ~150 B for the chain of kernels this replaces on a single GPU.
constexpr unsigned long long LB_VAL = (1ULL << 62) - 1;
constexpr unsigned long long LB_READY = 1ULL << 62;

__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}

// MINB: blocks per SM the register allocation aims at; U: output slots a thread expands at a time.  The kernel issues
// ~340 instructions per particle (the FP64 exponential and division of the weight, nine 128-bit crossing counts per
// thread, an 11-step search per output slot) and waits on memory in between: ncu at 2^23 particles shows the issue
// slot 43 % busy and 5 long-scoreboard stalls per issue.  More resident warps hide more of that latency (6 blocks per SM
// at 42 registers beat the unconstrained 58 registers / 4 blocks above 2^21 particles), two independent searches and
// eight row loads in flight per thread a little more: <5, 2> measures 46 / 110 / 184 us at 2^20 / 2^22 / 2^23 particles
// against 56 / 118 / 192 for <6, 1>; <4, 4> 46 / 114 / 196; <3, 8> 67 / 153 / 265 (profiles/resample_variants_r02.log).
template <int MINB, int U>
__global__ void __launch_bounds__(SB, MINB)
resample_fused_kernel(const double* __restrict__ lk, const double* __restrict__ w_in, int64_t n, uint64_t N,
                      unsigned long long carry_q, int first_shard, int64_t m_out,
                      const double* __restrict__ max_dev, double gm, const double* __restrict__ sum_w_dev, double Nd,
                      double inv_Np, uint64_t u0q, unsigned long long* agg /*[2*tiles]: q, floor|ready*/,
                      unsigned long long* inc /*[2*tiles]*/, unsigned* __restrict__ tile_counter, const double* __restrict__ src, int64_t ld_src, int rows,
                      double* __restrict__ dst, int64_t ld_dst, int32_t* __restrict__ anc_out,
                      int32_t* __restrict__ counts_out, int64_t* __restrict__ filled_out) {
    __shared__ unsigned long long sm_q[32];
    __shared__ long long sm_f[32];
    __shared__ unsigned s_end[TILE];          // inclusive end offset of every particle of the tile, relative to the tile's first slot
    __shared__ unsigned long long s_base[2];
    __shared__ unsigned s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    const int64_t base = (int64_t)tile * TILE + (int64_t)threadIdx.x * IPT;   // blocked layout

    // ---- 1. weights, floor counts, fixed-point residuals (the arithmetic of weights_kernel + prepare_kernel) ----
    unsigned long long q[IPT];
    int32_t fl[IPT];
    unsigned long long tq = 0;
    long long tf = 0;
    const double mx = w_in ? 0.0 : max_dev[0], sw = w_in ? 1.0 : sum_w_dev[0];
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        q[k] = 0;
        fl[k] = 0;
        if (base + k < n) {
            double wj;
            if (w_in != nullptr) {
                wj = w_in[base + k];
            } else {
                const double d = __dsub_rn(lk[base + k], mx);
                wj = __ddiv_rn(exp(__dmul_rn(d, gm)), sw);              // Micmem_SMC_main.py:124-130
            }
            const double f = trunc(__dmul_rn(wj, Nd));                  // np.trunc(p_weight*n_particle)
            const double r = __dsub_rn(wj, __dmul_rn(f, inv_Np));       // p_weight - p_is*inv_Np
            double rq = r * TWO62;
            rq = (rq > 0.0) ? rq : 0.0;
            q[k] = (unsigned long long)__double2ull_rn(rq);
            fl[k] = (int32_t)f;
            tq += q[k];
            tf += fl[k];
        }
    }
    unsigned long long bq;
    long long bf;
    const unsigned long long ex_q = block_exclusive_scan<unsigned long long>(tq, 0ULL, OpAdd(), sm_q, &bq);
    const long long ex_f = block_exclusive_scan<long long>(tf, 0LL, OpAdd(), sm_f, &bf);

    // ---- 2. decoupled look-back: exclusive prefixes of this tile ----
    // Every tile publishes (q, floor) twice, each into write-once words: its own aggregate, later its inclusive
    // prefix.  The floor word carries the 2-bit "ready" mark and is stored after the q word (fence in between), so a
    // reader that has seen the mark reads a q of the same publication.
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        unsigned long long pre_q = 0, pre_f = 0;
        if (tile > 0) {
            if (lane == 0) {
                agg[2 * (size_t)tile] = bq;
                __threadfence();
                atomicExch(&agg[2 * (size_t)tile + 1], LB_READY | (unsigned long long)bf);
            }
            long long look = (long long)tile - 1;
            for (;;) {
                const long long i = look - lane;
                bool is_inc = true;                      // tiles before tile 0: inclusive prefix 0
                unsigned long long vq = 0, vf = 0;
                if (i >= 0) {
                    for (;;) {
                        vf = ld_volatile(&inc[2 * (size_t)i + 1]);
                        if (vf >> 62) break;
                        vf = ld_volatile(&agg[2 * (size_t)i + 1]);
                        if (vf >> 62) {
                            is_inc = false;
                            break;
                        }
                    }
                    __threadfence();
                    vq = ld_volatile(is_inc ? &inc[2 * (size_t)i] : &agg[2 * (size_t)i]);
                    vf &= LB_VAL;
                }
                // aggregates of the nearer tiles up to, and including, the nearest inclusive prefix
                const unsigned inc_mask = __ballot_sync(FULL_MASK, is_inc);
                const int stop = inc_mask ? (__ffs(inc_mask) - 1) : 32;
                unsigned long long cq = (lane <= stop) ? vq : 0, cf = (lane <= stop) ? vf : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    cq += __shfl_xor_sync(FULL_MASK, cq, o);
                    cf += __shfl_xor_sync(FULL_MASK, cf, o);
                }
                pre_q += cq;
                pre_f += cf;
                if (stop < 32) break;
                look -= 32;
            }
        }
        if (lane == 0) {
            inc[2 * (size_t)tile] = pre_q + bq;
            __threadfence();
            atomicExch(&inc[2 * (size_t)tile + 1], LB_READY | (pre_f + (unsigned long long)bf));
            s_base[0] = pre_q;
            s_base[1] = pre_f;
        }
    }
    __syncthreads();
    // residual prefix of the whole run: what the lower ranks hold (carry_q, 0 on one GPU) + this shard's prefix
    const unsigned long long base_q = carry_q + s_base[0];
    const long long base_f = (long long)s_base[1];

    // ---- 3. copy counts and output offsets ----
    // thresholds crossed before this shard's first particle (none before the first particle of the run)
    const long long c_start = first_shard ? 0 : crossings(carry_q, N, u0q);
    // first output slot of the tile: floor prefix + thresholds crossed on this shard before its first particle
    const long long tile_c0 = (tile == 0) ? c_start : crossings(base_q, N, u0q);
    const long long tile_off = base_f + (tile_c0 - c_start);
    unsigned long long sq = base_q + ex_q;
    long long c_prev = crossings(sq, N, u0q);
    if (base == 0) c_prev = c_start;
    long long off = base_f + ex_f + (c_prev - c_start);    // copies placed on this shard before this thread's first particle
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        sq += q[k];
        const long long c = crossings(sq, N, u0q);
        const int32_t cnt = fl[k] + (int32_t)(c - c_prev);
        c_prev = c;
        off += cnt;
        s_end[threadIdx.x * IPT + k] = (unsigned)(off - tile_off);
        if (counts_out != nullptr && base + k < n) counts_out[base + k] = cnt;
    }
    __syncthreads();
    const long long tile_total = (long long)s_end[TILE - 1];
    if (tile == (unsigned)(n_tiles - 1) && threadIdx.x == 0) filled_out[0] = tile_off + tile_total;

    // ---- 4. cooperative expansion + gather of this tile's output slots ----
    // U slots per thread at a time: their searches are independent chains of shared-memory loads and their row loads
    // are all in flight together (U = 1: one search, then one particle's rows, per trip).
    const int64_t tile_first = (int64_t)tile * TILE;
    long long lim = tile_total;                                  // slots of this tile that this shard fills
    if (m_out - tile_off < lim) lim = m_out - tile_off;          // (copies beyond m_out are dropped)
    for (long long sl0 = threadIdx.x; sl0 < lim; sl0 += (long long)U * SB) {
        int pos[U];
#pragma unroll
        for (int u = 0; u < U; ++u) pos[u] = 0;
        // number of end offsets <= sl  =  smallest i with s_end[i] > sl   (TILE is a power of two: fixed trip count)
#pragma unroll
        for (int step = TILE / 2; step > 0; step >>= 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned sl = (unsigned)(sl0 + (long long)u * SB);
                if (s_end[pos[u] + step - 1] <= sl) pos[u] += step;
            }
        }
        int64_t a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            a[u] = tile_first + (pos[u] < TILE - 1 ? pos[u] : TILE - 1);
            if (sl0 + (long long)u * SB < lim) anc_out[tile_off + sl0 + (long long)u * SB] = (int32_t)a[u];
        }
        for (int k0 = 0; k0 < rows; k0 += 4) {
            double v[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k0 + k < rows && sl0 + (long long)u * SB < lim) v[u][k] = src[(int64_t)(k0 + k) * ld_src + a[u]];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k0 + k < rows && sl0 + (long long)u * SB < lim)
                        dst[(int64_t)(k0 + k) * ld_dst + tile_off + sl0 + (long long)u * SB] = v[u][k];
        }
    }
}

// slots [filled, n) of a mis-filled resampling (the reference does not guard the rounding of its running sum,
// SURVEY.md H2) repeat the last ancestor
__global__ void resample_pad_kernel(const int64_t* __restrict__ filled, int64_t n, const double* __restrict__ src,
                                    int64_t ld_src, int rows, double* __restrict__ dst, int64_t ld_dst,
                                    int32_t* __restrict__ anc) {
    const int64_t f = filled[0];
    if (f >= n) return;
    const int32_t a = (f > 0) ? anc[f - 1] : 0;
    for (int64_t s = f + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        anc[s] = a;
        for (int k = 0; k < rows; ++k) dst[(int64_t)k * ld_dst + s] = src[(int64_t)k * ld_src + a];
    }
}

inline int64_t tiles_of(int64_t n) { return (n + TILE - 1) / TILE; }

}  // namespace

extern "C" int smcb_resample_totals(smcb_handle* h, const double* w_dev, int64_t n, int64_t n_total,
                                    int64_t* totals_dev, void* stream) {
    REQUIRE(h, h && w_dev && totals_dev && n > 0 && n_total >= n, SMCB_ERR_INVALID, "bad argument");
    REQUIRE(h, h->floor_cnt != nullptr && n <= h->n_max, SMCB_ERR_STATE, "smcb_reserve too small");
    REQUIRE(h, n_total < (1LL << 31), SMCB_ERR_UNSUPPORTED, "n_total must be below 2^31");
    cudaStream_t st = as_stream(stream);
    const int64_t nt = tiles_of(n);
    prepare_kernel<SMCB_SCAN_FIXED><<<(unsigned)nt, SB, 0, st>>>(w_dev, n, (double)n_total, 1.0 / (double)n_total,
                                                               h->floor_cnt, h->resid_f, h->resid_q, h->tile_tot);
    LAUNCH_CHECK(h);
    totals_from_tiles_kernel<<<1, SB, 0, st>>>(h->tile_tot, nt, totals_dev);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_resample_counts(smcb_handle* h, const double* w_dev, int64_t n, int64_t n_total, double u0,
                                    int mode, double* carry_host, uint64_t carry_q, int64_t id_offset,
                                    int32_t* counts_dev, int64_t* totals_dev, void* stream) {
    REQUIRE(h, h && w_dev && counts_dev && totals_dev && n > 0 && n_total >= n, SMCB_ERR_INVALID, "bad argument");
    REQUIRE(h, h->floor_cnt != nullptr && n <= h->n_max, SMCB_ERR_STATE, "smcb_reserve too small");
    REQUIRE(h, n_total < (1LL << 31), SMCB_ERR_UNSUPPORTED, "n_total must be below 2^31");
    REQUIRE(h, u0 >= 0.0 && u0 < 1.0, SMCB_ERR_INVALID, "u0 must lie in [0,1)");
    cudaStream_t st = as_stream(stream);
    const int64_t nt = tiles_of(n);
    const double Nd = (double)n_total;
    const double inv_Np = 1.0 / Nd;   // Micmem_settings.py:17
    if (mode == SMCB_SCAN_SEQUENTIAL) {
        prepare_kernel<SMCB_SCAN_SEQUENTIAL><<<(unsigned)nt, SB, 0, st>>>(w_dev, n, Nd, inv_Np, h->floor_cnt,
                                                                        h->resid_f, h->resid_q, h->tile_tot);
        LAUNCH_CHECK(h);
        double carry[2] = {0.0, u0 * inv_Np};   // wrand = rand()*inv_Np (Micmem_SMC_main.py:156)
        if (carry_host != nullptr) {
            carry[0] = carry_host[0];
            carry[1] = carry_host[1];
        }
        CUDA_TRY(h, cudaMemcpyAsync(h->seq_carry, carry, sizeof(carry), cudaMemcpyHostToDevice, st));
        // [2], [3] of seq_carry: chunks walked / chunks that needed the literal loop (diagnostics, as doubles' bits)
        counts_sequential_block_kernel<<<1, SQB, 0, st>>>(h->floor_cnt, h->resid_f, n, inv_Np, h->seq_carry, counts_dev,
                                                         totals_dev,
                                                         reinterpret_cast<unsigned long long*>(h->seq_carry + 2));
        LAUNCH_CHECK(h);
        if (carry_host != nullptr) {
            CUDA_TRY(h, cudaMemcpyAsync(carry_host, h->seq_carry, sizeof(carry), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(h, cudaStreamSynchronize(st));
        }
        return SMCB_OK;
    }
    REQUIRE(h, mode == SMCB_SCAN_FIXED, SMCB_ERR_INVALID, "unknown scan mode");
    prepare_kernel<SMCB_SCAN_FIXED><<<(unsigned)nt, SB, 0, st>>>(w_dev, n, Nd, inv_Np, h->floor_cnt, h->resid_f,
                                                               h->resid_q, h->tile_tot);
    LAUNCH_CHECK(h);
    scan_tiles_pair_kernel<<<1, SB, 0, st>>>(h->tile_tot, nt, 0, (int64_t)carry_q, totals_dev);
    LAUNCH_CHECK(h);
    const uint64_t u0q = (uint64_t)llrint(u0 * TWO62);
    counts_fixed_kernel<<<(unsigned)nt, SB, 0, st>>>(h->floor_cnt, h->resid_q, n, h->tile_tot, (uint64_t)n_total,
                                                    u0q, id_offset == 0 ? 1 : 0, counts_dev);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_ancestors(smcb_handle* h, const int32_t* counts_dev, int64_t n, int64_t m,
                              int32_t* ancestors_dev, int64_t* filled_dev, void* stream) {
    REQUIRE(h, h && counts_dev && ancestors_dev && n > 0 && m > 0, SMCB_ERR_INVALID, "bad argument");
    REQUIRE(h, h->mark != nullptr && n <= h->n_max && m <= h->n_max, SMCB_ERR_STATE, "smcb_reserve too small");
    cudaStream_t st = as_stream(stream);
    const int64_t nt = tiles_of(n), mt = tiles_of(m);
    count_tiles_kernel<<<(unsigned)nt, SB, 0, st>>>(counts_dev, n, h->tile_tot);
    LAUNCH_CHECK(h);
    scan_tiles_pair_kernel<<<1, SB, 0, st>>>(h->tile_tot, nt, 0, 0, h->tile_tot2);
    LAUNCH_CHECK(h);
    if (filled_dev != nullptr)
        CUDA_TRY(h, cudaMemcpyAsync(filled_dev, h->tile_tot2, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(h, cudaMemsetAsync(h->mark, 0, sizeof(int32_t) * m, st));
    mark_heads_kernel<<<(unsigned)nt, SB, 0, st>>>(counts_dev, n, h->tile_tot, m, h->mark);
    LAUNCH_CHECK(h);
    int32_t* tile_max = reinterpret_cast<int32_t*>(h->tile_tot2 + 2);
    max_tiles_kernel<<<(unsigned)mt, SB, 0, st>>>(h->mark, m, tile_max);
    LAUNCH_CHECK(h);
    scan_tiles_max_kernel<<<1, SB, 0, st>>>(tile_max, mt);
    LAUNCH_CHECK(h);
    fill_ancestors_kernel<<<(unsigned)mt, SB, 0, st>>>(h->mark, m, tile_max, ancestors_dev);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_gather(smcb_handle* h, const double* src_dev, int64_t ld_src, const int32_t* ancestors_dev,
                           int64_t m, int rows, double* dst_dev, int64_t ld_dst, void* stream) {
    REQUIRE(h, h && src_dev && ancestors_dev && dst_dev && m > 0 && rows > 0, SMCB_ERR_INVALID, "bad argument");
    gather_kernel<<<(unsigned)((m + 255) / 256), 256, 0, as_stream(stream)>>>(src_dev, ld_src, ancestors_dev, m, rows,
                                                                            dst_dev, ld_dst);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}


extern "C" int smcb_resample_fused(smcb_handle* h, const double* lk_dev, const double* w_dev, int64_t n, int64_t n_total,
                                   uint64_t carry_q, int64_t id_offset, int64_t m_out, const double* max_dev, double gm,
                                   const double* sum_w_dev, double u0, const double* src_dev, int64_t ld_src, int rows,
                                   double* dst_dev, int64_t ld_dst, int32_t* ancestors_dev, int32_t* counts_dev,
                                   int64_t* filled_dev, void* stream) {
    REQUIRE(h, h && src_dev && dst_dev && filled_dev && n > 0 && rows > 0 && ld_src >= n, SMCB_ERR_INVALID, "bad argument");
    REQUIRE(h, n_total >= n && m_out >= 0 && m_out <= n_total && ld_dst >= m_out, SMCB_ERR_INVALID, "bad sizes");
    REQUIRE(h, w_dev != nullptr || (lk_dev && max_dev && sum_w_dev), SMCB_ERR_INVALID,
            "give either normalised weights or (lk, max, sum_w)");
    REQUIRE(h, h->mark != nullptr && h->rs_desc != nullptr && n <= h->n_max && m_out <= h->n_max, SMCB_ERR_STATE,
            "smcb_reserve too small");
    REQUIRE(h, n_total < (1LL << 31), SMCB_ERR_UNSUPPORTED, "n_total must be below 2^31");
    REQUIRE(h, u0 >= 0.0 && u0 < 1.0, SMCB_ERR_INVALID, "u0 must lie in [0,1)");
    cudaStream_t st = as_stream(stream);
    const int64_t nt = tiles_of(n);
    // tile counter and look-back words in one buffer, cleared by one memset: [counter | agg 2*nt | inc 2*nt]
    unsigned long long* desc = h->rs_desc;
    unsigned* counter = reinterpret_cast<unsigned*>(desc);
    unsigned long long* agg = desc + 1;
    unsigned long long* inc = agg + 2 * (size_t)nt;
    CUDA_TRY(h, cudaMemsetAsync(desc, 0, sizeof(unsigned long long) * (1 + 4 * (size_t)nt), st));
    int32_t* anc = ancestors_dev ? ancestors_dev : h->mark;
    const double Nd = (double)n_total;
    const uint64_t u0q = (uint64_t)llrint(u0 * TWO62);
#define RS_ARGS lk_dev, w_dev, n, (uint64_t)n_total, carry_q, id_offset == 0 ? 1 : 0, m_out, max_dev, gm, sum_w_dev, Nd, \
                1.0 / Nd, u0q, agg, inc, counter, src_dev, ld_src, rows, dst_dev, ld_dst, anc, counts_dev, filled_dev
    resample_fused_kernel<5, 2><<<(unsigned)nt, SB, 0, st>>>(RS_ARGS);
#undef RS_ARGS
    LAUNCH_CHECK(h);
    if (m_out > 0) {
        resample_pad_kernel<<<8, 256, 0, st>>>(filled_dev, m_out, src_dev, ld_src, rows, dst_dev, ld_dst, anc);
        LAUNCH_CHECK(h);
    }
    return SMCB_OK;
}
