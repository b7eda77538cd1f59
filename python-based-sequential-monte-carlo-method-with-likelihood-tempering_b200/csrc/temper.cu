// K2: tempering reductions.  Replaces the NumPy lines of the reference's back-off loop
// (SMC_example/Micmem_SMC_main.py:116-134): max(lk), and for a batch of candidate increments
// gm_k the sums  sum_i exp((lk_i-max)*gm_k)  and  sum_i exp(...)^2  from which the host forms
// ESS_k = (sum w)^2 / (N * sum w^2) and the log-evidence increment.  HBM-bound: 8 B per particle
// per pass, K candidates share one read.  Reductions are two-level (block partials, then one
// block sums the partials in a fixed order) so results do not depend on scheduling.
#include "common.cuh"
#include "exp_table.cuh"

namespace {

constexpr int RB = 256;   // reduction block

// vec: lk is 16-byte aligned and pairs are read with one 128-bit load; otherwise (odd leading dimension) the same
// pairs are read with two 64-bit loads - the element-to-thread mapping, hence every sum, is the same either way.
__global__ void __launch_bounds__(RB) max_partial_kernel(const double* __restrict__ lk, int64_t n,
                                                         double* __restrict__ partial, bool vec) {
    __shared__ double sm[32];
    double m = -INFINITY;
    const int64_t stride = (int64_t)gridDim.x * RB * 2;
    for (int64_t i = ((int64_t)blockIdx.x * RB + threadIdx.x) * 2; i < n; i += stride) {
        if (i + 1 < n) {
            const double2 v = vec ? *reinterpret_cast<const double2*>(lk + i) : make_double2(lk[i], lk[i + 1]);
            m = fmax(m, fmax(v.x, v.y));
        } else {
            m = fmax(m, lk[i]);
        }
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = (threadIdx.x < RB / 32) ? sm[threadIdx.x] : -INFINITY;
        m = warp_max(m);
        if (threadIdx.x == 0) partial[blockIdx.x] = m;
    }
}

__global__ void __launch_bounds__(RB) max_final_kernel(const double* __restrict__ partial, int nb,
                                                       double* __restrict__ out) {
    __shared__ double sm[32];
    double m = -INFINITY;
    for (int i = threadIdx.x; i < nb; i += RB) m = fmax(m, partial[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = (threadIdx.x < RB / 32) ? sm[threadIdx.x] : -INFINITY;
        m = warp_max(m);
        if (threadIdx.x == 0) out[0] = m;
    }
}

struct GmList {
    double gm[SMCB_MAX_CAND];
};

template <int K>
__global__ void __launch_bounds__(RB)
temper_partial_kernel(const double* __restrict__ lk, int64_t n, const double* __restrict__ max_dev,
                      const GmList gms, double* __restrict__ partial, bool vec) {
    __shared__ double sm[2 * K * 32];
    // exp() is what this pass costs (K per particle): the table exponential of exp_table.cuh (10 FP64 operations,
    // < 1.1 ulp, 0 below -708 where the library returns subnormals) instead of the library's ~22
    __shared__ double etab[expt::TAB_N];
    expt::load_table(etab);
    __syncthreads();
    const double mx = max_dev[0];
    double acc[2 * K];
#pragma unroll
    for (int k = 0; k < 2 * K; ++k) acc[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * RB * 2;
    for (int64_t i = ((int64_t)blockIdx.x * RB + threadIdx.x) * 2; i < n; i += stride) {
        double d0, d1;
        bool two = i + 1 < n;
        if (two) {
            const double2 v = vec ? *reinterpret_cast<const double2*>(lk + i) : make_double2(lk[i], lk[i + 1]);
            d0 = v.x - mx;
            d1 = v.y - mx;
        } else {
            d0 = lk[i] - mx;
            d1 = 0.0;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double w0 = expt::exp_fast(d0 * gms.gm[k], etab);
            acc[2 * k] += w0;
            acc[2 * k + 1] = fma(w0, w0, acc[2 * k + 1]);
            if (two) {
                const double w1 = expt::exp_fast(d1 * gms.gm[k], etab);
                acc[2 * k] += w1;
                acc[2 * k + 1] = fma(w1, w1, acc[2 * k + 1]);
            }
        }
    }
    block_sum<2 * K>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 2 * K; ++k) partial[(int64_t)blockIdx.x * (2 * K) + k] = acc[k];
    }
}

// out[c] = sum_b partial[b*ncol + c], summed in a fixed order (thread-strided, then tree).
// out[c] = sum_b partial[b*ncol + c]: one warp per column (lane-strided partial sums in block order, then the warp
// butterfly), fixed order.  Round 1 walked the columns one after the other in a single block: 11 us per call against 3.
__global__ void __launch_bounds__(RB) colsum_final_kernel(const double* __restrict__ partial, int nb,
                                                          int ncol, double* __restrict__ out) {
    const int c = blockIdx.x * (RB / 32) + (threadIdx.x >> 5);
    if (c >= ncol) return;
    double v = 0.0;
    for (int b = threadIdx.x & 31; b < nb; b += 32) v += partial[(int64_t)b * ncol + c];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) out[c] = v;
}

__global__ void weights_kernel(const double* __restrict__ lk, int64_t n, const double* __restrict__ max_dev,
                               double gm, const double* __restrict__ sum_w_dev, double* __restrict__ w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // p_weight = exp(d_lk*gm); p_weight = p_weight/sum_weight   (Micmem_SMC_main.py:124-130)
    const double d = __dsub_rn(lk[i], max_dev[0]);
    w[i] = __ddiv_rn(exp(__dmul_rn(d, gm)), sum_w_dev[0]);
}

// Merge of the per-shard rows (max_r, S1_r[k], S2_r[k]) of one tempering round, in rank order (every rank computes
// the same bits): logsumexp rescale to the global maximum.  One block; thread k owns candidate k.
struct GmWide {
    double gm[3 * SMCB_MAX_CAND];
};
__global__ void temper_merge_kernel(const double* __restrict__ rows, int world, int stride, int n_cand,
                                    const GmWide g, double* __restrict__ out) {
    double mx = -INFINITY;
    for (int r = 0; r < world; ++r) mx = fmax(mx, rows[(size_t)r * stride]);
    if (threadIdx.x == 0) out[0] = mx;
    const int k = threadIdx.x;
    if (k >= n_cand) return;
    double s1 = 0.0, s2 = 0.0;
    for (int r = 0; r < world; ++r) {
        const double* row = rows + (size_t)r * stride;
        const double mr = row[0];
        if (world > 1 && mr == -INFINITY) continue;       // a shard without a finite likelihood carries no weight
        const double f = exp((mr - mx) * g.gm[k]);        // exactly 1 for the shard that holds the maximum
        s1 += row[1 + 2 * k] * f;
        s2 += row[2 + 2 * k] * (f * f);
    }
    out[2 + 2 * k] = s1;
    out[3 + 2 * k] = s2;
}

inline int reduce_grid(const smcb_handle* h, int64_t n) {
    int64_t nb = (n + RB * 2 - 1) / (RB * 2);
    const int64_t cap = (int64_t)h->sm_count * 8;
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    return (int)nb;
}

}  // namespace

extern "C" int smcb_lk_max(smcb_handle* h, const double* lk_dev, int64_t n, double* out_dev, void* stream) {
    REQUIRE(h, h && lk_dev && out_dev && n > 0, SMCB_ERR_INVALID, "null pointer or n<=0");
    REQUIRE(h, h->partial != nullptr, SMCB_ERR_STATE, "smcb_reserve has not been called");
    const bool vec = (reinterpret_cast<uintptr_t>(lk_dev) & 15) == 0;
    const int nb = reduce_grid(h, n);
    cudaStream_t st = as_stream(stream);
    max_partial_kernel<<<nb, RB, 0, st>>>(lk_dev, n, h->partial, vec);
    LAUNCH_CHECK(h);
    max_final_kernel<<<1, RB, 0, st>>>(h->partial, nb, out_dev);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}

extern "C" int smcb_temper_sums(smcb_handle* h, const double* lk_dev, int64_t n, const double* max_dev,
                                const double* gm_host, int n_cand, double* out_dev, void* stream) {
    REQUIRE(h, h && lk_dev && max_dev && gm_host && out_dev && n > 0, SMCB_ERR_INVALID, "null pointer or n<=0");
    REQUIRE(h, n_cand >= 1 && n_cand <= SMCB_MAX_CAND, SMCB_ERR_INVALID, "n_cand out of range");
    REQUIRE(h, h->partial != nullptr, SMCB_ERR_STATE, "smcb_reserve has not been called");
    const bool vec = (reinterpret_cast<uintptr_t>(lk_dev) & 15) == 0;
    GmList g;
    for (int k = 0; k < SMCB_MAX_CAND; ++k) g.gm[k] = (k < n_cand) ? gm_host[k] : 0.0;
    const int nb = reduce_grid(h, n);
    cudaStream_t st = as_stream(stream);
    int K;
    if (n_cand <= 1) {
        K = 1;
        temper_partial_kernel<1><<<nb, RB, 0, st>>>(lk_dev, n, max_dev, g, h->partial, vec);
    } else if (n_cand <= 2) {
        K = 2;
        temper_partial_kernel<2><<<nb, RB, 0, st>>>(lk_dev, n, max_dev, g, h->partial, vec);
    } else if (n_cand <= 4) {
        K = 4;
        temper_partial_kernel<4><<<nb, RB, 0, st>>>(lk_dev, n, max_dev, g, h->partial, vec);
    } else if (n_cand <= 8) {
        K = 8;
        temper_partial_kernel<8><<<nb, RB, 0, st>>>(lk_dev, n, max_dev, g, h->partial, vec);
    } else {
        K = 16;
        temper_partial_kernel<16><<<nb, RB, 0, st>>>(lk_dev, n, max_dev, g, h->partial, vec);
    }
    LAUNCH_CHECK(h);
    // partial is [nb][2K]; only the first 2*n_cand columns are wanted, but they are the leading
    // columns of each row, so reduce with row stride 2K.
    colsum_final_kernel<<<(2 * K + RB / 32 - 1) / (RB / 32), RB, 0, st>>>(h->partial, nb, 2 * K, h->partial + (int64_t)nb * 2 * K);
    LAUNCH_CHECK(h);
    CUDA_TRY(h, cudaMemcpyAsync(out_dev, h->partial + (int64_t)nb * 2 * K, sizeof(double) * 2 * n_cand,
                                cudaMemcpyDeviceToDevice, st));
    return SMCB_OK;
}

extern "C" int smcb_weights(smcb_handle* h, const double* lk_dev, int64_t n, const double* max_dev, double gm,
                            const double* sum_w_dev, double* w_dev, void* stream) {
    REQUIRE(h, h && lk_dev && max_dev && sum_w_dev && w_dev && n > 0, SMCB_ERR_INVALID, "null pointer or n<=0");
    weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(lk_dev, n, max_dev, gm, sum_w_dev,
                                                                             w_dev);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}


extern "C" int smcb_temper_eval(smcb_handle* h, const double* lk_dev, int64_t n, const double* gm_host, int n_cand,
                                double* out_dev, void* stream) {
    REQUIRE(h, h && lk_dev && gm_host && out_dev && n > 0, SMCB_ERR_INVALID, "null pointer or n<=0");
    REQUIRE(h, n_cand >= 1 && n_cand <= 3 * SMCB_MAX_CAND, SMCB_ERR_INVALID, "n_cand out of range");
    int rc = comm_staging(h);
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    double* row = h->comm_send;                      // [0] = shard max, [1+2k], [2+2k] = sums relative to it
    if ((rc = smcb_lk_max(h, lk_dev, n, row, stream))) return rc;
    for (int o = 0; o < n_cand; o += SMCB_MAX_CAND) {
        const int len = (n_cand - o < SMCB_MAX_CAND) ? n_cand - o : SMCB_MAX_CAND;
        if ((rc = smcb_temper_sums(h, lk_dev, n, row, gm_host + o, len, row + 1 + 2 * o, stream))) return rc;
    }
    const int stride = 1 + 2 * n_cand;
    if ((rc = comm_all_gather_f64(h, row, h->comm_recv, stride, st))) return rc;
    GmWide g;
    for (int k = 0; k < 3 * SMCB_MAX_CAND; ++k) g.gm[k] = (k < n_cand) ? gm_host[k] : 0.0;
    temper_merge_kernel<<<1, 64, 0, st>>>(h->comm_recv, h->world, stride, n_cand, g, out_dev);
    LAUNCH_CHECK(h);
    return SMCB_OK;
}
