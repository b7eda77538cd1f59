// K4: Metropolis-Hastings mutation.  Replaces SMC_example/Micmem_SMC_main.py:209-249:
//   cov = np.cov(p_filt.T, bias=True) * w_cov          -> colsum + centred second moments (two-pass,
//                                                         like np.cov), factorised on the host (d<=32)
//   p_pred = p_filt + MVN(0,cov,N) * mhstep_ratio      -> smcb_mh_propose (z @ F, z from Philox or given)
//   p0 = prior(p_pred) > 0; replace out-of-box         -> same kernel (closed box test)
//   r = exp((lk2-lk1)*gamma)*p0 >= U; select; r_ac     -> smcb_mh_accept
#include "common.cuh"
#include "philox.cuh"


namespace {

constexpr int MB = 256;

struct MhParams {
    double F[SMCB_MAX_DIM * SMCB_MAX_DIM];   // row-major, x = z @ F
    double low[SMCB_MAX_DIM];
    double high[SMCB_MAX_DIM];
};

struct MeanVec {
    double m[SMCB_MAX_DIM];
};

// ---- column sums ------------------------------------------------------------------------------
// grid = (blocks, d): block (b, k) reduces a strided slice of row k
__global__ void __launch_bounds__(MB)
colsum_partial_kernel(const double* __restrict__ theta, int64_t ld, int64_t n, double* __restrict__ partial) {
    __shared__ double sm[32];
    const int k = blockIdx.y, d = gridDim.y;
    const double* row = theta + (int64_t)k * ld;
    double v[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * MB;
    for (int64_t i = (int64_t)blockIdx.x * MB + threadIdx.x; i < n; i += stride) v[0] += row[i];
    block_sum<1>(v, sm);
    if (threadIdx.x == 0) partial[(int64_t)blockIdx.x * d + k] = v[0];
}

// ---- centred second moments, small d: every thread keeps the D(D+1)/2 upper triangle ------------
template <int D>
__global__ void __launch_bounds__(MB)
moments_small_kernel(const double* __restrict__ theta, int64_t ld, int64_t n, const double* __restrict__ mean,
                     double* __restrict__ partial) {
    constexpr int NP = D * (D + 1) / 2;
    __shared__ double sm[NP * 32];
    double mu[D];
#pragma unroll
    for (int a = 0; a < D; ++a) mu[a] = mean[a];
    double acc[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) acc[q] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * MB;
    for (int64_t i = (int64_t)blockIdx.x * MB + threadIdx.x; i < n; i += stride) {
        double x[D];
#pragma unroll
        for (int a = 0; a < D; ++a) x[a] = theta[(int64_t)a * ld + i] - mu[a];
        int q = 0;
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int b = a; b < D; ++b) {
                acc[q] = fma(x[a], x[b], acc[q]);
                ++q;
            }
    }
    block_sum<NP>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NP; ++q) partial[(int64_t)blockIdx.x * NP + q] = acc[q];
    }
}

// ---- centred second moments, general d<=32: particles staged in shared memory, one (a,b) pair
// per thread (threads >= d(d+1)/2 idle).  Block = 544 threads covers d=32 (528 pairs).
constexpr int GEN_THREADS = 544;
constexpr int GEN_TILE = 64;   // particles per shared-memory tile
__global__ void __launch_bounds__(GEN_THREADS)
moments_general_kernel(const double* __restrict__ theta, int64_t ld, int64_t n, int d,
                       const double* __restrict__ mean, double* __restrict__ partial) {
    __shared__ double sx[SMCB_MAX_DIM][GEN_TILE + 1];
    const int npair = d * (d + 1) / 2;
    int a = 0, b = 0;
    if ((int)threadIdx.x < npair) {   // unrank (a<=b) from the linear upper-triangle index
        int q = threadIdx.x;
        while (q >= d - a) {
            q -= d - a;
            ++a;
        }
        b = a + q;
    }
    double acc = 0.0;
    const int64_t n_tiles = (n + GEN_TILE - 1) / GEN_TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i0 = tile * GEN_TILE;
        __syncthreads();
        for (int idx = threadIdx.x; idx < d * GEN_TILE; idx += GEN_THREADS) {
            const int k = idx / GEN_TILE, j = idx - k * GEN_TILE;
            sx[k][j] = (i0 + j < n) ? theta[(int64_t)k * ld + i0 + j] - mean[k] : 0.0;
        }
        __syncthreads();
        if ((int)threadIdx.x < npair) {
#pragma unroll 8
            for (int j = 0; j < GEN_TILE; ++j) acc = fma(sx[a][j], sx[b][j], acc);
        }
    }
    if ((int)threadIdx.x < npair) partial[(int64_t)blockIdx.x * npair + threadIdx.x] = acc;
}

// Final reduction of the centred second moments straight into the full symmetric matrix: one warp per (a <= b) pair
// sums that column of the block partials (lane-strided, then a shuffle tree: a fixed order) and writes [a][b], [b][a].
__global__ void __launch_bounds__(256)
final_sym_kernel(const double* __restrict__ partial, int nb, int d, double* __restrict__ out) {
    const int npair = d * (d + 1) / 2;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= npair) return;
    double v = 0.0;
    for (int b = lane; b < nb; b += 32) v += partial[(int64_t)b * npair + q];
    v = warp_sum(v);
    if (lane == 0) {
        int a = 0, r = q;
        while (r >= d - a) {
            r -= d - a;
            ++a;
        }
        const int bb = a + r;
        out[a * d + bb] = v;
        out[bb * d + a] = v;
    }
}

// Final reduction of the column sums of one shard plus the head of its all-gather row (smcb_moments_merged):
// row[0:4] = MH counters, row[4] = n_r, row[5:5+d] = column sums, mean_r = column sums / n_r.  One warp per column.
// (row_head == nullptr, mean_r == nullptr: plain column sums into sums, smcb_colsum.)
__global__ void __launch_bounds__(256)
colsum_final_row_kernel(const double* __restrict__ partial, int nb, int d, const unsigned long long* __restrict__ counts,
                        double n_local, double* __restrict__ row_head, double* __restrict__ sums,
                        double* __restrict__ mean_r) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row_head != nullptr && blockIdx.x == 0 && threadIdx.x < 5)
        row_head[threadIdx.x] = (threadIdx.x < 4) ? (counts ? (double)counts[threadIdx.x] : 0.0) : n_local;
    if (k >= d) return;
    double v = 0.0;
    for (int b = lane; b < nb; b += 32) v += partial[(int64_t)b * d + k];
    v = warp_sum(v);
    if (lane == 0) {
        sums[k] = v;
        if (mean_r != nullptr) mean_r[k] = v / n_local;        // IEEE division, as np.cov's X.mean()
    }
}

// ---- proposal -------------------------------------------------------------------------------------
// DT>0: compile-time dimension (registers, fully unrolled); DT==0: run-time d<=32 (local array).
// F_dev != nullptr: the factor is read from device memory (staged in shared memory) instead of the parameter block.
template <int DT>
__global__ void __launch_bounds__(MB)
propose_kernel(const double* __restrict__ theta, int64_t ld, int64_t n, int d_rt, const __grid_constant__ MhParams prm,
               const double* __restrict__ F_dev, double ratio, const double* __restrict__ z_ext, uint64_t seed,
               uint64_t id_offset, uint32_t stage, uint32_t sweep, double* __restrict__ prop, int64_t ld_prop,
               uint8_t* __restrict__ inbox) {
    const int d = DT ? DT : d_rt;
    __shared__ double sF[(DT ? DT * DT : SMCB_MAX_DIM * SMCB_MAX_DIM)];
    if (F_dev != nullptr) {
        for (int k = threadIdx.x; k < d * d; k += MB) sF[k] = F_dev[k];
    } else {
        for (int k = threadIdx.x; k < d * d; k += MB) sF[k] = prm.F[k];
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * MB + threadIdx.x;
    if (i >= n) return;
    double step[DT ? DT : SMCB_MAX_DIM];
#pragma unroll
    for (int k = 0; k < (DT ? DT : SMCB_MAX_DIM); ++k) step[k] = 0.0;
    // x = z @ F  accumulated row by row (j outer) so only two normals are live at a time
#pragma unroll
    for (int j = 0; j < d; j += 2) {
        double z0, z1 = 0.0;
        if (z_ext != nullptr) {
            z0 = z_ext[i * d + j];
            if (j + 1 < d) z1 = z_ext[i * d + j + 1];
        } else {
            philox_normal2(seed, id_offset + (uint64_t)i, stage, sweep, (uint32_t)(j >> 1), &z0, &z1);
        }
#pragma unroll
        for (int k = 0; k < d; ++k) step[k] += z0 * sF[j * d + k];
        if (j + 1 < d) {
#pragma unroll
            for (int k = 0; k < d; ++k) step[k] += z1 * sF[(j + 1) * d + k];
        }
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < d; ++k) {
        const double cur = theta[(int64_t)k * ld + i];
        const double x = __dadd_rn(cur, __dmul_rn(step[k], ratio));   // p_filt + MVN*mhstep_ratio
        step[k] = x;
        ok = ok && (x >= prm.low[k]) && (x <= prm.high[k]);          // closed box (uniform pdf > 0)
    }
#pragma unroll
    for (int k = 0; k < d; ++k)
        prop[(int64_t)k * ld_prop + i] = ok ? step[k] : theta[(int64_t)k * ld + i];
    inbox[i] = ok ? 1 : 0;
}

// ---- accept ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MB)
accept_kernel(double* __restrict__ theta, int64_t ld, double* __restrict__ lk, const double* __restrict__ prop,
              int64_t ld_prop, const double* __restrict__ lk2, const uint8_t* __restrict__ inbox, int64_t n, int d,
              double gamma, const double* __restrict__ u_ext, const double* __restrict__ dlp, uint64_t seed,
              uint64_t id_offset, uint32_t stage, uint32_t sweep, uint8_t* __restrict__ moved,
              unsigned long long* __restrict__ counts) {
    __shared__ long long sm[4][32];
    const int64_t i = (int64_t)blockIdx.x * MB + threadIdx.x;
    long long acc = 0, newly = 0, evald = 0, ninf = 0;
    if (i < n) {
        const double u = (u_ext != nullptr)
                             ? u_ext[i]
                             : philox_uniform(seed, id_offset + (uint64_t)i, stage, sweep, SMCB_SLOT_UNIFORM);
        double pp = 0.0;
        double l2 = 0.0;
        if (inbox[i]) {
            evald = 1;
            l2 = lk2[i];
            ninf = (l2 == -INFINITY);   // rejected early by smcb_loglik_bounded (or a genuinely impossible proposal)
            pp = exp(__dmul_rn(__dsub_rn(l2, lk[i]), gamma));   // exp(px*gamma_new)*p0
            if (dlp != nullptr) pp *= exp(dlp[i]);              // * p(theta')/p(theta)  (methanation main:369)
        }
        const bool r = pp >= u;   // NaN compares false, as in NumPy
        if (r) {
            acc = 1;
            if (inbox[i]) {
                for (int k = 0; k < d; ++k) theta[(int64_t)k * ld + i] = prop[(int64_t)k * ld_prop + i];
                lk[i] = l2;
            }
            if (!moved[i]) {
                moved[i] = 1;
                newly = 1;
            }
        }
    }
    acc = warp_sum_ll(acc);
    newly = warp_sum_ll(newly);
    evald = warp_sum_ll(evald);
    ninf = warp_sum_ll(ninf);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        sm[0][wid] = acc;
        sm[1][wid] = newly;
        sm[2][wid] = evald;
        sm[3][wid] = ninf;
    }
    __syncthreads();
    if (wid == 0) {
        acc = (lane < MB / 32) ? sm[0][lane] : 0;
        newly = (lane < MB / 32) ? sm[1][lane] : 0;
        evald = (lane < MB / 32) ? sm[2][lane] : 0;
        ninf = (lane < MB / 32) ? sm[3][lane] : 0;
        acc = warp_sum_ll(acc);
        newly = warp_sum_ll(newly);
        evald = warp_sum_ll(evald);
        ninf = warp_sum_ll(ninf);
        if (lane == 0) {
            if (acc) atomicAdd(&counts[0], (unsigned long long)acc);
            if (newly) atomicAdd(&counts[1], (unsigned long long)newly);
            if (evald) atomicAdd(&counts[2], (unsigned long long)evald);
            if (ninf) atomicAdd(&counts[3], (unsigned long long)ninf);
        }
    }
}

// ---- early-rejection threshold -----------------------------------------------------------------------
// accept_kernel takes a proposal iff exp((lk2-lk1)*gamma) >= u.  If lk2 < lk1 + log(u)/gamma - margin the
// left side is below u*exp(-margin*gamma), a relative gap (>= 1e-9) far wider than the rounding of the
// subtraction, the product and exp(), so the test fails for certain.
__global__ void __launch_bounds__(MB)
threshold_kernel(const double* __restrict__ lk, const uint8_t* __restrict__ inbox, int64_t n, double gamma,
                 const double* __restrict__ u_ext, const double* __restrict__ dlp, uint64_t seed, uint64_t id_offset,
                 uint32_t stage, uint32_t sweep, double* __restrict__ lkmin) {
    const int64_t i = (int64_t)blockIdx.x * MB + threadIdx.x;
    if (i >= n) return;
    double thr = -INFINITY;
    if (inbox == nullptr || inbox[i]) {
        const double u = (u_ext != nullptr)
                             ? u_ext[i]
                             : philox_uniform(seed, id_offset + (uint64_t)i, stage, sweep, SMCB_SLOT_UNIFORM);
        const double l1 = lk[i];
        if (u > 0.0 && gamma > 0.0 && isfinite(l1)) {
            // with a prior ratio the test is exp((lk2-lk1)*gamma + dlp) >= u: log(u) - dlp takes the place of log(u)
            const double lu = (log(u) - (dlp != nullptr ? dlp[i] : 0.0)) / gamma;
            thr = l1 + lu - 1e-9 * (1.0 + fabs(l1) + fabs(lu));
            if (!(thr == thr)) thr = -INFINITY;
        }
    }
    lkmin[i] = thr;
}

// ---- log prior ratio of independent normal components ----------------------------------------------
// out[i] = sum_k inv2var_k * ((theta_k - mu_k)^2 - (prop_k - mu_k)^2) = log p(prop) - log p(theta) over the
// normally distributed parameters (inv2var_k = 1/(2 sigma_k^2); 0 marks a uniform one, whose ratio is 1 in-box).
__global__ void __launch_bounds__(MB)
prior_logratio_kernel(const double* __restrict__ theta, int64_t ld, const double* __restrict__ prop, int64_t ld_prop,
                      int64_t n, int d, const __grid_constant__ MhParams prm, const uint8_t* __restrict__ inbox,
                      double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * MB + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    if (inbox == nullptr || inbox[i]) {
        for (int k = 0; k < d; ++k) {
            const double iv = prm.high[k];            // inv2var
            if (iv > 0.0) {
                const double a = theta[(int64_t)k * ld + i] - prm.low[k], b = prop[(int64_t)k * ld_prop + i] - prm.low[k];
                acc += iv * (a * a - b * b);
            }
        }
    }
    out[i] = acc;
}

__global__ void philox_draws_kernel(int64_t n, int d, uint64_t seed, uint64_t id_offset, uint32_t stage,
                                    uint32_t sweep, double* __restrict__ z, double* __restrict__ u) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (z != nullptr) {
        for (int j = 0; j < d; j += 2) {
            double z0, z1;
            philox_normal2(seed, id_offset + (uint64_t)i, stage, sweep, (uint32_t)(j >> 1), &z0, &z1);
            z[i * d + j] = z0;
            if (j + 1 < d) z[i * d + j + 1] = z1;
        }
    }
    if (u != nullptr) u[i] = philox_uniform(seed, id_offset + (uint64_t)i, stage, sweep, SMCB_SLOT_UNIFORM);
}

struct BoxParams {
    double low[SMCB_MAX_DIM];
    double high[SMCB_MAX_DIM];
};

__global__ void sample_box_kernel(double* __restrict__ theta, int64_t ld, int64_t n, int d, const BoxParams bx,
                                  uint64_t seed, uint64_t id_offset) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < d; ++k) {
        const double u = philox_uniform(seed, id_offset + (uint64_t)i, 0xFFFFFFFFu, (uint32_t)k, 0u);
        theta[(int64_t)k * ld + i] = __dadd_rn(bx.low[k], __dmul_rn(bx.high[k] - bx.low[k], u));   // no FMA: same bits as NumPy
    }
}

// ---- merged moments + proposal factor ---------------------------------------------------------------------------
// Row of one shard in the all-gather: [0:4] MH counters, [4] n_r, [5:5+d] column sums, [5+d:5+d+d*d] M2_r (second
// moments centred on the shard's own mean colsum_r/n_r, full symmetric matrix), then that mean (colsum_final_row_kernel).
// 1/x and 1/sqrt(x) from the MUFU seeds and one cubically convergent correction each (as csrc/kinetic.cuh): the
// Jacobi rotations below sit on the critical path of every sweep (one warp, strictly serial), where an IEEE FP64
// division or square root costs ~250 cycles of dependent instructions; the rotation angle needs no last-bit accuracy
// (Jacobi is self-correcting), the orthogonality of (c, s) is restored to 1 ulp by the correction.
__device__ __forceinline__ double rcp_fast(double x) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    const double e = fma(-x, r0, 1.0);
    return fma(r0, fma(e, e, e), r0);
}
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(-x * y0, y0, 1.0);                 // 1 - x*y0^2
    return fma(y0, e * fma(0.375, e, 0.5), y0);             // y0*(1 + e/2 + 3e^2/8)
}

struct WCov {
    double w[SMCB_MAX_DIM * SMCB_MAX_DIM];
    int use;
};

// One warp.  Merge in rank order (identical bits on every rank), then cyclic Jacobi on cov (*) w_cov.
__global__ void __launch_bounds__(32)
moments_merge_factor_kernel(const double* __restrict__ rows, int world, int stride, int d, double n_total,
                            const __grid_constant__ WCov wc, double* __restrict__ out) {
    __shared__ double A[SMCB_MAX_DIM][SMCB_MAX_DIM + 1];
    __shared__ double V[SMCB_MAX_DIM][SMCB_MAX_DIM + 1];
    __shared__ double mean[SMCB_MAX_DIM], lam[SMCB_MAX_DIM], sdev[SMCB_MAX_DIM];
    __shared__ int order[SMCB_MAX_DIM];
    const int lane = threadIdx.x;
    if (lane < 4) {
        double c = 0.0;
        for (int r = 0; r < world; ++r) c += rows[(size_t)r * stride + lane];
        out[lane] = c;
    }
    if (lane < d) {
        double m;
        if (world == 1) {
            m = rows[5 + d + d * d + lane];          // the shard mean the centred pass used (bit-identical to round 1)
        } else {
            double sum = 0.0;
            for (int r = 0; r < world; ++r) sum += rows[(size_t)r * stride + 5 + lane];
            m = sum / n_total;
        }
        mean[lane] = m;
        out[4 + lane] = m;
    }
    __syncwarp();
    for (int idx = lane; idx < d * d; idx += 32) {
        int a = idx / d, b = idx - a * d;
        if (a > b) { const int c = a; a = b; b = c; }          // canonical (a <= b): exactly symmetric output
        double acc = 0.0;
        for (int r = 0; r < world; ++r) {
            const double* row = rows + (size_t)r * stride;
            double m2 = row[5 + d + a * d + b];
            if (world > 1) {
                const double nr = row[4];
                const double da = row[5 + a] / nr - mean[a], db = row[5 + b] / nr - mean[b];
                m2 = fma(nr * da, db, m2);                     // Chan et al.: + n_r (mean_r - mean)(mean_r - mean)^T
            }
            acc += m2;
        }
        out[4 + d + idx] = acc;
        const int i = idx / d, j = idx - i * d;
        A[i][j] = (acc / n_total) * (wc.use ? wc.w[idx] : 1.0);   // cov (*) w_cov (Micmem_SMC_main.py:212-215)
        V[i][j] = (i == j) ? 1.0 : 0.0;
    }
    __syncwarp();
    // symmetrise against an asymmetric w_cov (the reference's is symmetric)
    for (int idx = lane; idx < d * d; idx += 32) {
        const int i = idx / d, j = idx - i * d;
        if (i < j) {
            const double v = 0.5 * (A[i][j] + A[j][i]);
            A[i][j] = v;
            A[j][i] = v;
        }
    }
    __syncwarp();
    // Scale to unit diagonal first, C = D^1/2 R D^1/2: parameters of very different magnitude (methanation: 1e8 next to
    // 1) would otherwise push the small eigenvalues of C below the rounding noise of the large ones.  The factor is
    // then F = sqrt(|Lambda|) V^T D^1/2 with R = V Lambda V^T, and F^T F = C.
    if (lane < d) sdev[lane] = (A[lane][lane] > 0.0) ? sqrt(A[lane][lane]) : 0.0;
    __syncwarp();
    for (int idx = lane; idx < d * d; idx += 32) {
        const int i = idx / d, j = idx - i * d;
        const double si = sdev[i], sj = sdev[j];
        double v = (si > 0.0 && sj > 0.0) ? A[i][j] / (si * sj) : 0.0;
        if (i == j) v = (si > 0.0) ? 1.0 : 0.0;
        A[i][j] = v;
    }
    __syncwarp();
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, dg = 0.0;
        if (lane < d)
            for (int a = 0; a < d; ++a) {
                const double v = A[a][lane];
                if (a == lane) dg = fma(v, v, dg);
                else off = fma(v, v, off);
            }
        off = warp_sum(off);
        dg = warp_sum(dg);
        if (!(off > 1e-40 * dg)) break;                         // off-diagonal below 1e-20 relative (or all zero / NaN)
        for (int p = 0; p < d - 1; ++p)
            for (int q = p + 1; q < d; ++q) {
                const double apq = A[p][q], app = A[p][p], aqq = A[q][q];
                __syncwarp();
                if (apq == 0.0) continue;                       // warp-uniform
                // |apq| so small that tau overflows the seeds' range: the rotation is the identity to working precision
                if (fabs(apq) < 1e-140 * fabs(aqq - app) || fabs(apq) < 1e-290) continue;
                const double tau = (aqq - app) * rcp_fast(2.0 * apq);
                const double t2 = fma(tau, tau, 1.0);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) * rcp_fast(fabs(tau) + t2 * rsqrt_fast(t2));
                const double c = rsqrt_fast(fma(t, t, 1.0)), sn = t * c;
                if (lane < d) {                                  // columns p, q
                    const double akp = A[lane][p], akq = A[lane][q];
                    A[lane][p] = c * akp - sn * akq;
                    A[lane][q] = sn * akp + c * akq;
                    const double vkp = V[lane][p], vkq = V[lane][q];
                    V[lane][p] = c * vkp - sn * vkq;
                    V[lane][q] = sn * vkp + c * vkq;
                }
                __syncwarp();
                if (lane < d) {                                  // rows p, q
                    const double apk = A[p][lane], aqk = A[q][lane];
                    A[p][lane] = c * apk - sn * aqk;
                    A[q][lane] = sn * apk + c * aqk;
                }
                __syncwarp();
                if (lane == 0) {
                    A[p][q] = 0.0;
                    A[q][p] = 0.0;
                }
                __syncwarp();
            }
    }
    if (lane < d) lam[lane] = A[lane][lane];
    __syncwarp();
    double lmax = 0.0;
    for (int k = 0; k < d; ++k) lmax = fmax(lmax, fabs(lam[k]));
    if (lane < d) {                                              // descending order, index breaks ties
        int rank = 0;
        for (int k = 0; k < d; ++k) rank += (lam[k] > lam[lane]) || (lam[k] == lam[lane] && k < lane);
        order[rank] = lane;
    }
    __syncwarp();
    double* F = out + 4 + d + d * d;
    if (lane < d) {
        const int j = order[lane];
        double big = -1.0, sgn = 1.0;
        for (int k = 0; k < d; ++k) {
            const double v = V[k][j];
            if (fabs(v) > big) {
                big = fabs(v);
                sgn = (v < 0.0) ? -1.0 : 1.0;
            }
        }
        const double al = fabs(lam[j]);
        const double sc = (al <= 1e-13 * lmax) ? 0.0 : sgn * sqrt(al);
        for (int k = 0; k < d; ++k) F[lane * d + k] = (sc * V[k][j]) * sdev[k];
    }
}

inline int moments_grid(const smcb_handle* h, int64_t n) {
    int64_t nb = (n + MB - 1) / MB;
    const int64_t cap = (int64_t)h->sm_count * 4;
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    return (int)nb;
}

}  // namespace

extern "C" int smcb_colsum(smcb_handle* h, const double* theta_dev, int64_t ld, int64_t n, int d, double* out_dev,
                           void* stream) {
    REQUIRE(h, h && theta_dev && out_dev && n > 0 && d >= 1 && d <= SMCB_MAX_DIM && ld >= n, SMCB_ERR_INVALID,
            "bad argument");
    REQUIRE(h, h->partial != nullptr, SMCB_ERR_STATE, "smcb_reserve has not been called");
    cudaStream_t st = as_stream(stream);
    const int nb = moments_grid(h, n);
    REQUIRE(h, (int64_t)nb * d <= h->partial_len, SMCB_ERR_STATE, "reduction scratch too small");
    colsum_partial_kernel<<<dim3(nb, d), MB, 0, st>>>(theta_dev, ld, n, h->partial);
    LAUNCH_CHECK(h);
    colsum_final_row_kernel<<<(d + 7) / 8, 256, 0, st>>>(h->partial, nb, d, nullptr, (double)n, nullptr, out_dev, nullptr);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_centered_moments(smcb_handle* h, const double* theta_dev, int64_t ld, int64_t n, int d,
                                     const double* mean_dev, double* out_dev, void* stream) {
    REQUIRE(h, h && theta_dev && mean_dev && out_dev && n > 0 && d >= 1 && d <= SMCB_MAX_DIM && ld >= n,
            SMCB_ERR_INVALID, "bad argument");
    REQUIRE(h, h->partial != nullptr, SMCB_ERR_STATE, "smcb_reserve has not been called");
    cudaStream_t st = as_stream(stream);
    const int npair = d * (d + 1) / 2;
    int nb = moments_grid(h, n);
    switch (d) {
#define CASE(D)                                                                                        \
    case D:                                                                                            \
        moments_small_kernel<D><<<nb, MB, 0, st>>>(theta_dev, ld, n, mean_dev, h->partial);            \
        break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6)
#undef CASE
        default: {
            const int64_t tiles = (n + GEN_TILE - 1) / GEN_TILE;
            nb = (int)((tiles < (int64_t)h->sm_count * 2) ? tiles : (int64_t)h->sm_count * 2);
            REQUIRE(h, ((int64_t)nb + 1) * npair <= h->partial_len, SMCB_ERR_STATE, "reduction scratch too small");
            moments_general_kernel<<<nb, GEN_THREADS, 0, st>>>(theta_dev, ld, n, d, mean_dev, h->partial);
        }
    }
    LAUNCH_CHECK(h);
    final_sym_kernel<<<(npair + 7) / 8, 256, 0, st>>>(h->partial, nb, d, out_dev);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_mh_propose(smcb_handle* h, const double* theta_dev, int64_t ld, int64_t n, int d,
                               const double* F_host, double ratio, const double* low_host, const double* high_host,
                               const double* z_dev, uint64_t seed, uint64_t id_offset, uint32_t stage,
                               uint32_t sweep, double* prop_dev, int64_t ld_prop, uint8_t* inbox_dev, void* stream) {
    REQUIRE(h, h && theta_dev && F_host && low_host && high_host && prop_dev && inbox_dev, SMCB_ERR_INVALID,
            "null pointer");
    REQUIRE(h, n > 0 && d >= 1 && d <= SMCB_MAX_DIM && ld >= n && ld_prop >= n, SMCB_ERR_INVALID, "bad size");
    REQUIRE(h, sweep < (1u << 24), SMCB_ERR_INVALID, "sweep index too large");
    MhParams prm;
    memset(&prm, 0, sizeof(prm));
    memcpy(prm.F, F_host, sizeof(double) * d * d);
    memcpy(prm.low, low_host, sizeof(double) * d);
    memcpy(prm.high, high_host, sizeof(double) * d);
    const unsigned grid = (unsigned)((n + MB - 1) / MB);
    cudaStream_t st = as_stream(stream);
#define PROPOSE(DT)                                                                                              \
    propose_kernel<DT><<<grid, MB, 0, st>>>(theta_dev, ld, n, d, prm, nullptr, ratio, z_dev, seed, id_offset, stage, \
                                            sweep, prop_dev, ld_prop, inbox_dev)
    switch (d) {
        case 1: PROPOSE(1); break;
        case 2: PROPOSE(2); break;
        case 3: PROPOSE(3); break;
        case 4: PROPOSE(4); break;
        case 5: PROPOSE(5); break;
        case 6: PROPOSE(6); break;
        case 8: PROPOSE(8); break;
        default: PROPOSE(0);
    }
#undef PROPOSE
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_mh_propose_dev(smcb_handle* h, const double* theta_dev, int64_t ld, int64_t n, int d,
                                   const double* F_dev, double ratio, const double* low_host, const double* high_host,
                                   const double* z_dev, uint64_t seed, uint64_t id_offset, uint32_t stage,
                                   uint32_t sweep, double* prop_dev, int64_t ld_prop, uint8_t* inbox_dev, void* stream) {
    REQUIRE(h, h && theta_dev && F_dev && low_host && high_host && prop_dev && inbox_dev, SMCB_ERR_INVALID,
            "null pointer");
    REQUIRE(h, n > 0 && d >= 1 && d <= SMCB_MAX_DIM && ld >= n && ld_prop >= n, SMCB_ERR_INVALID, "bad size");
    REQUIRE(h, sweep < (1u << 24), SMCB_ERR_INVALID, "sweep index too large");
    MhParams prm;
    memset(&prm, 0, sizeof(prm));
    memcpy(prm.low, low_host, sizeof(double) * d);
    memcpy(prm.high, high_host, sizeof(double) * d);
    const unsigned grid = (unsigned)((n + MB - 1) / MB);
    cudaStream_t st = as_stream(stream);
#define PROPOSE(DT)                                                                                              \
    propose_kernel<DT><<<grid, MB, 0, st>>>(theta_dev, ld, n, d, prm, F_dev, ratio, z_dev, seed, id_offset, stage, \
                                            sweep, prop_dev, ld_prop, inbox_dev)
    switch (d) {
        case 1: PROPOSE(1); break;
        case 2: PROPOSE(2); break;
        case 3: PROPOSE(3); break;
        case 4: PROPOSE(4); break;
        case 5: PROPOSE(5); break;
        case 6: PROPOSE(6); break;
        case 8: PROPOSE(8); break;
        default: PROPOSE(0);
    }
#undef PROPOSE
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_moments_merged(smcb_handle* h, const double* theta_dev, int64_t ld, int64_t n, int d, int64_t n_total,
                                   const int64_t* counts_dev, const double* w_cov_host, double* out_dev, void* stream) {
    REQUIRE(h, h && theta_dev && out_dev && n > 0 && d >= 1 && d <= SMCB_MAX_DIM && ld >= n && n_total >= n,
            SMCB_ERR_INVALID, "bad argument");
    int rc = comm_staging(h);
    if (rc) return rc;
    const int stride = 5 + d + d * d + d;           // the shard mean rides along at the end of the row (world == 1 path)
    REQUIRE(h, stride <= SMCB_COMM_ROW, SMCB_ERR_UNSUPPORTED, "row too long for the staging buffer");
    cudaStream_t st = as_stream(stream);
    double* row = h->comm_send;
    double* mean_r = row + 5 + d + d * d;
    REQUIRE(h, h->partial != nullptr, SMCB_ERR_STATE, "smcb_reserve has not been called");
    // five launches (+ the all-gather on more than one GPU): column-sum partials, their final reduction together
    // with the head of the row and the shard mean, centred partials, their final reduction into the symmetric
    // matrix, merge + factor
    const int nbc = moments_grid(h, n);
    REQUIRE(h, (int64_t)nbc * d <= h->partial_len, SMCB_ERR_STATE, "reduction scratch too small");
    colsum_partial_kernel<<<dim3(nbc, d), MB, 0, st>>>(theta_dev, ld, n, h->partial);
    LAUNCH_CHECK(h);
    colsum_final_row_kernel<<<(d + 7) / 8, 256, 0, st>>>(h->partial, nbc, d, reinterpret_cast<const unsigned long long*>(counts_dev),
                                                       (double)n, row, row + 5, mean_r);
    LAUNCH_CHECK(h);
    if ((rc = smcb_centered_moments(h, theta_dev, ld, n, d, mean_r, row + 5 + d, stream))) return rc;
    const double* rows = row;
    if (h->world > 1) {
        if ((rc = comm_all_gather_f64(h, row, h->comm_recv, stride, st))) return rc;
        rows = h->comm_recv;
    }
    WCov wc;
    wc.use = w_cov_host != nullptr;
    if (wc.use) memcpy(wc.w, w_cov_host, sizeof(double) * d * d);
    moments_merge_factor_kernel<<<1, 32, 0, st>>>(rows, h->world, stride, d, (double)n_total, wc, out_dev);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_prior_logratio(smcb_handle* h, const double* theta_dev, int64_t ld, const double* prop_dev,
                                   int64_t ld_prop, int64_t n, int d, const double* mu_host, const double* inv2var_host,
                                   const uint8_t* inbox_dev, double* out_dev, void* stream) {
    REQUIRE(h, h && theta_dev && prop_dev && mu_host && inv2var_host && out_dev, SMCB_ERR_INVALID, "null pointer");
    REQUIRE(h, n > 0 && d >= 1 && d <= SMCB_MAX_DIM && ld >= n && ld_prop >= n, SMCB_ERR_INVALID, "bad size");
    MhParams prm;
    memset(&prm, 0, sizeof(prm));
    for (int k = 0; k < d; ++k) {
        prm.low[k] = mu_host[k];
        prm.high[k] = inv2var_host[k];
    }
    prior_logratio_kernel<<<(unsigned)((n + MB - 1) / MB), MB, 0, as_stream(stream)>>>(theta_dev, ld, prop_dev, ld_prop, n,
                                                                                   d, prm, inbox_dev, out_dev);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_mh_threshold(smcb_handle* h, const double* lk_dev, const uint8_t* inbox_dev, int64_t n, double gamma,
                                 const double* u_dev, const double* dlp_dev, uint64_t seed, uint64_t id_offset,
                                 uint32_t stage, uint32_t sweep, double* lkmin_dev, void* stream) {
    REQUIRE(h, h && lk_dev && lkmin_dev && n > 0, SMCB_ERR_INVALID, "bad argument");
    threshold_kernel<<<(unsigned)((n + MB - 1) / MB), MB, 0, as_stream(stream)>>>(
        lk_dev, inbox_dev, n, gamma, u_dev, dlp_dev, seed, id_offset, stage, sweep, lkmin_dev);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_mh_accept(smcb_handle* h, double* theta_dev, int64_t ld, double* lk_dev, const double* prop_dev,
                              int64_t ld_prop, const double* lk2_dev, const uint8_t* inbox_dev, int64_t n, int d,
                              double gamma, const double* u_dev, const double* dlp_dev, uint64_t seed,
                              uint64_t id_offset, uint32_t stage, uint32_t sweep, uint8_t* moved_dev,
                              int64_t* counts_dev, void* stream) {
    REQUIRE(h, h && theta_dev && lk_dev && prop_dev && lk2_dev && inbox_dev && moved_dev && counts_dev,
            SMCB_ERR_INVALID, "null pointer");
    REQUIRE(h, n > 0 && d >= 1 && d <= SMCB_MAX_DIM && ld >= n && ld_prop >= n, SMCB_ERR_INVALID, "bad size");
    accept_kernel<<<(unsigned)((n + MB - 1) / MB), MB, 0, as_stream(stream)>>>(
        theta_dev, ld, lk_dev, prop_dev, ld_prop, lk2_dev, inbox_dev, n, d, gamma, u_dev, dlp_dev, seed, id_offset,
        stage, sweep, moved_dev, reinterpret_cast<unsigned long long*>(counts_dev));
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_philox_draws(smcb_handle* h, int64_t n, int d, uint64_t seed, uint64_t id_offset, uint32_t stage,
                                 uint32_t sweep, double* z_dev, double* u_dev, void* stream) {
    REQUIRE(h, h && n > 0 && d >= 1 && d <= SMCB_MAX_DIM, SMCB_ERR_INVALID, "bad argument");
    philox_draws_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(n, d, seed, id_offset, stage,
                                                                                  sweep, z_dev, u_dev);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_sample_uniform_box(smcb_handle* h, double* theta_dev, int64_t ld, int64_t n, int d,
                                       const double* low_host, const double* high_host, uint64_t seed,
                                       uint64_t id_offset, void* stream) {
    REQUIRE(h, h && theta_dev && low_host && high_host && n > 0 && d >= 1 && d <= SMCB_MAX_DIM && ld >= n,
            SMCB_ERR_INVALID, "bad argument");
    BoxParams bx;
    memset(&bx, 0, sizeof(bx));
    memcpy(bx.low, low_host, sizeof(double) * d);
    memcpy(bx.high, high_host, sizeof(double) * d);
    sample_box_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(theta_dev, ld, n, d, bx, seed,
                                                                                id_offset);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}
