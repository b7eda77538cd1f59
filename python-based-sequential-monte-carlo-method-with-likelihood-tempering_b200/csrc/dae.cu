// K1'' (SURVEY.md 8(f) N3): the reference's transient fixed-bed reactor model, one thread block per
// (particle, operating condition).
//
// Replaces the body of my_model (SMC_methanation/methanation_set_likelihood.py:144-277): 7 x 51 = 357 unknowns
// (five concentrations, temperature, velocity on 51 axial nodes), residual `reaction` (:69-139) restated node by
// node below and pinned against the reference's own function (tests/golden/methanation_dae_residual.npz through
// oracle/methanation_dae.py, which this file twins).  The reference integrates with SUNDIALS IDA (variable-order
// BDF, absent here); the integrator is builder-defined: implicit Euler on a geometric time grid from 0 to 75 s
// (host-supplied; a grid step whose Newton iteration fails is retried in smaller pieces), finite-difference
// Jacobian.  Node j's equations involve nodes j-1, j, j+1 only, so with the unknowns ordered node by node the
// Jacobian is block tridiagonal with 7x7 blocks:
//   * residuals and the 3 x 7 perturbed residuals per node are independent tasks spread over the block's threads
//     (the reaction rate, the expensive part, is reused when a neighbour is perturbed);
//   * the linear solve is a twisted block Thomas sweep (from both ends to the middle node): each node's 7 x 21
//     system [D' | C | I] is reduced by Gauss-Jordan with row pivoting in shared memory, 147 threads on one element
//     of each of the two chains, and the factors are kept;
//   * Newton is the modified kind: the factors serve the following iterations and the following pieces of the
//     same size (one residual pass, one parallel product and two register recurrences run by two warps) until the
//     update stops shrinking by 0.3x, then they are refreshed.
// Everything lives in shared memory (73 KB per block, three blocks per SM).  Cost: 35 steps x (1-2 Jacobians +
// ~6 substitutions) per march; this is the like-for-like physics mode for reference-sized particle counts, the
// plug-flow RK4 march of kinetic.cu is the throughput mode.
#include <vector>

#include "common.cuh"
#include "kinetic.cuh"

namespace {

constexpr int NX = 51;        // methanation_set_conditon.py:44
constexpr int NV = 7;         // C_H2, C_CO2, C_CH4, C_H2O, C_Ar, T, u
constexpr int NB = NV * NV;
constexpr int DAE_THREADS = 160;
constexpr int MCOLS = 3 * NV; // elimination scratch [D' | C | I]: 7 x 21, one thread per element
constexpr int MROW = 22;      // its row stride
constexpr int NEWTON_MAX = 40;
constexpr int MAX_RETRY = 8;  // failed pieces allowed within one grid step
constexpr int MID = NX / 2;   // node where the two elimination chains meet

constexpr double DZ_DISP = 0.95e-5;   // Dz   set_conditon.py:76
constexpr double RHOS = 5075.0;       //      :77
constexpr double CPS = 698.0;         //      :83
constexpr double KEFF = 0.72;         //      :84
constexpr double T_BED0 = 400.0;      // SMC_methanation.py:421
constexpr double NEWTON_TOL = 1e-10, FD_REL = 1e-7;

struct Case {
    double Cin[5], T_in, T_j, u_in, voidf, dz, P0;
    double inv_dz, inv_dz2;   // 1/dz, 1/dz^2
    double kA[4], nEoR[4];    // prefactors and -E/R of (kf, ks, kCO2, kH2O)
};

__device__ __forceinline__ double floor_of(int v) { return v < 5 ? 1e-3 : (v == 5 ? 1.0 : 1e-4); }

// func_rCH4 (set_likelihood.py:44-58); quotients as MUFU-seeded reciprocals (kin::rcp, < 1 ulp)
__device__ __forceinline__ double rate_ch4(const Case& c, double T, double Ca, double Cb, double Cc, double Cd) {
    const double RT6 = kin::R_GAS * T * 1e-6;
    const double PH2 = Ca * RT6, PCO2 = Cb * RT6, PCH4 = Cc * RT6, PH2O = Cd * RT6;
    const double iT = kin::rcp(T);
    const double kf = c.kA[0] * exp(c.nEoR[0] * iT);
    const double ks = c.kA[1] * exp(c.nEoR[1] * iT);
    const double kC = c.kA[2] * exp(c.nEoR[2] * iT);
    const double kW = c.kA[3] * exp(c.nEoR[3] * iT);
    const double dC = 1.0 + kC * PCO2, dW = 1.0 + kW * PH2O;
    const double rf = 5075e3 * kf * kC * PCO2 * kin::sqrt_fast(fmax(0.001, PH2)) * kin::rcp(dC * dC);
    const double rr = 5075e3 * ks * kW * PH2O * (PCH4 * PCH4) * kin::rcp(dW * dW);
    return rf - rr;
}

// func_rohg (:61-66)
__device__ __forceinline__ double density(const Case& c, const double* y) {
    return c.P0 / kin::R_GAS * kin::rcp(y[5]) * (y[0] * 2 + y[1] * 44 + y[2] * 16 + y[3] * 18 + y[4] * 40) *
           kin::rcp(y[0] + y[1] + y[2] + y[3] + y[4]) * 0.001;
}

// Equations of node j (`reaction` :69-139) given the unknowns of nodes j-1 (yl), j (yc), j+1 (yr), the previous
// time level of node j and 1/dt.  out[0..4] species balances, out[5] the equation the reference keeps in the T slot
// (continuity; u condition at the outlet), out[6] the one in the u slot (energy; u = u_in at the inlet, T condition
// at the outlet).  r, rho: rate and gas density of node j.
__device__ __forceinline__ void node_equations(const Case& c, int j, const double* yl, const double* yc,
                                               const double* yr, const double* yold, double inv_dt, double r,
                                               double rho, double* out) {
    const double sc[5] = {-4.0, -1.0, 1.0, 2.0, 0.0};
    if (j == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) out[k] = (yc[k] - yold[k]) * inv_dt;
        out[6] = yc[6] - c.u_in;
        return;
    }
    if (j == NX - 1) {
#pragma unroll
        for (int k = 0; k < 5; ++k) out[k] = yc[k] - yl[k];
        out[5] = yc[6] - yl[6];
        out[6] = yc[5] - yl[5];
        return;
    }
    const double idz = c.inv_dz, idz2 = c.inv_dz2, vd = c.voidf;
    const double T = yc[5], u = yc[6], Tl = yl[5], ul = yl[6], Tr = yr[5];
    const double iT = kin::rcp(T), iTl = kin::rcp(Tl), iTr = kin::rcp(Tr);
    const double dT = (T - yold[5]) * inv_dt;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double lap = (j == 1) ? (yr[k] - yc[k]) : (yr[k] - 2 * yc[k] + yl[k]);
        out[k] = -vd * ((yc[k] - yold[k]) * inv_dt) - (u * yc[k] - ul * yl[k]) * idz + vd * DZ_DISP * lap * idz2 +
                 (1 - vd) * sc[k] * r;
    }
    double cont = -u * c.P0 * (iT - iTl) * idz - c.P0 * iT * (u - ul) * idz +
                  vd * DZ_DISP * c.P0 * (iTr - 2 * iT + iTl) * idz2 + (1 - vd) * kin::R_GAS * (-2) * r;
    if (j == 1) cont += c.P0 * vd * (iT * iT) * dT;
    out[5] = cont;
    const double store = (j == 1) ? 1.0 : 0.1;
    out[6] = -store * (vd * rho * kin::CPG + (1 - vd) * RHOS * CPS) * dT - rho * kin::CPG * (T * u - Tl * ul) * idz +
             KEFF * (Tr - 2 * T + Tl) * idz2 + (1 - vd) * (-kin::HR) * r - 2 * kin::U_WALL / kin::DINT * (T - c.T_j);
}

struct Smem {
    double Y[NX * NV];      // unknowns, node by node
    double Yold[NX * NV];
    double F[NX * NV];      // residual, then Thomas g / Newton update
    double A[NX * NB];      // dF_j/dY_{j-1}, then L_j = inv D'_j A_j
    double D[NX * NB];      // dF_j/dY_j, then inv D'_j of the block factorisation
    double C[NX * NB];      // dF_j/dY_{j+1}, then W_j = inv D'_j C_j
    double M[4 * NV * MROW];              // elimination scratch: two systems, double-buffered; between
                                          // factorisations its first NX*NV words hold h = -inv D F
    double rr[NX], rho[NX];
    double red[DAE_THREADS / 32];
    Case cs;
    int fail;
};

__global__ void __launch_bounds__(DAE_THREADS, 3)
dae_march_kernel(const double* __restrict__ theta, int64_t ld, int64_t n, const unsigned* __restrict__ list,
                 const unsigned* __restrict__ count, const double* __restrict__ cond,
                 const double* __restrict__ obs, int n_cond, const double* __restrict__ base,
                 const int* __restrict__ inv_pos, const double* __restrict__ dts, int n_dt,
                 double* __restrict__ ssr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x;
    const int64_t m = (count != nullptr) ? (int64_t)*count : n;
    for (int64_t item = blockIdx.x; item < m * n_cond; item += gridDim.x) {
        const int ci = (int)(item / m);
        const int64_t q = item - (int64_t)ci * m;
        const int64_t p = (list != nullptr) ? (int64_t)list[q] : q;
        __syncthreads();   // previous item's readers of shared memory are done
        if (tid == 0) {
            const double* row = cond + (int64_t)ci * SMCB_KIN_NCOND_FIELDS;
            double csum = 0.0;
            for (int k = 0; k < 5; ++k) {
                s.cs.Cin[k] = row[k];
                csum += row[k] * kin::R_GAS * row[5];
            }
            s.cs.T_in = row[5];
            s.cs.T_j = row[6];
            s.cs.u_in = row[7];
            s.cs.voidf = row[8];
            s.cs.dz = row[9] / (NX - 1);
            s.cs.inv_dz = 1.0 / s.cs.dz;
            s.cs.inv_dz2 = s.cs.inv_dz * s.cs.inv_dz;
            s.cs.P0 = csum;
            for (int k = 0; k < 8; ++k) {
                const int ip = inv_pos[k];
                const double val = (ip >= 0) ? theta[(int64_t)ip * ld + p] : base[k];
                if (k & 1) s.cs.nEoR[k >> 1] = -val / kin::R_GAS;
                else s.cs.kA[k >> 1] = val;
            }
            s.fail = 0;
        }
        __syncthreads();
        const Case& c = s.cs;
        for (int e = tid; e < NX * NV; e += DAE_THREADS) {   // start state (SMC_methanation.py:412-423)
            const int j = e / NV, v = e - j * NV;
            s.Y[e] = (v < 5) ? c.Cin[v] : (v == 5 ? (j == 0 ? c.T_in : T_BED0) : c.u_in);
        }
        __syncthreads();
        // Time stepping.  One pass of the loop below is one attempted implicit-Euler piece.  A grid step H whose Newton
        // iteration fails is retried from the same state in pieces: the piece is quartered after a failure and doubled
        // after a success; more than MAX_RETRY failures within one grid step fail the march.
        bool failed = false, steady = false;
        double dt_fac = 0.0, H_prev = 0.0;   // step size the stored factors belong to; previous grid step
        int step = 0, n_retry = 0;
        double H = dts[0], t_left = H, dt = H;
        while (!failed && !steady) {
            dt = fmin(dt, t_left);
            const double inv_dt = 1.0 / dt;
            const bool whole_repeat = dt == H && H == H_prev;
            for (int e = tid; e < NX * NV; e += DAE_THREADS) s.Yold[e] = s.Y[e];
            __syncthreads();
            // the factors of the previous piece serve this one when the step size is the same (the 5 s steps of the
            // second half of the march, where the bed is close to its steady state)
            bool converged = false, bad = false, need_jac = !(dt == dt_fac);
            double prev_worst = INFINITY;
            for (int it = 0; it < NEWTON_MAX && !converged && !bad; ++it) {
                // ---- residual of every node, rate and density kept for the Jacobian
                if (tid < NX) {
                    const int j = tid;
                    const double* yc = s.Y + j * NV;
                    const double* yl = s.Y + (j > 0 ? j - 1 : j) * NV;
                    const double* yr = s.Y + (j < NX - 1 ? j + 1 : j) * NV;
                    double r = 0.0, rho = 0.0, out[NV];
                    if (j > 0 && j < NX - 1) {
                        r = rate_ch4(c, yc[5], yc[0], yc[1], yc[2], yc[3]);
                        rho = density(c, yc);
                    }
                    s.rr[j] = r;
                    s.rho[j] = rho;
                    node_equations(c, j, yl, yc, yr, s.Yold + j * NV, inv_dt, r, rho, out);
                    bool fin = true;
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        s.F[j * NV + k] = out[k];
                        fin = fin && isfinite(out[k]);
                    }
                    if (!fin) s.fail = 1;
                }
                __syncthreads();
                if (s.fail) {
                    bad = true;
                    break;
                }
                if (need_jac) {
                    need_jac = false;
                    dt_fac = dt;
                    // ---- finite-difference blocks: task = (node j, neighbour nb, variable v)
                    for (int t = tid; t < NX * 3 * NV; t += DAE_THREADS) {
                        const int j = t / (3 * NV), rem = t - j * 3 * NV, nb = rem / NV, v = rem - nb * NV;
                        const int jn = j - 1 + nb;
                        double* blk = (nb == 0 ? s.A : (nb == 1 ? s.D : s.C)) + j * NB;
                        if (jn < 0 || jn >= NX) {
#pragma unroll
                            for (int k = 0; k < NV; ++k) blk[k * NV + v] = 0.0;
                            continue;
                        }
                        double yl[NV], yc[NV], yr[NV];
#pragma unroll
                        for (int k = 0; k < NV; ++k) {
                            yc[k] = s.Y[j * NV + k];
                            yl[k] = s.Y[(j > 0 ? j - 1 : j) * NV + k];
                            yr[k] = s.Y[(j < NX - 1 ? j + 1 : j) * NV + k];
                        }
                        const double y0 = s.Y[jn * NV + v];
                        const double yp = y0 + FD_REL * fmax(fabs(y0), floor_of(v));
                        const double delta = yp - y0;
                        double r = s.rr[j], rho = s.rho[j];
#pragma unroll
                        for (int k = 0; k < NV; ++k) {   // compile-time indices keep the copies in registers
                            if (k == v) {
                                if (nb == 0) yl[k] = yp;
                                else if (nb == 1) yc[k] = yp;
                                else yr[k] = yp;
                            }
                        }
                        if (nb == 1 && j > 0 && j < NX - 1) {
                            r = rate_ch4(c, yc[5], yc[0], yc[1], yc[2], yc[3]);
                            rho = density(c, yc);
                        }
                        double out[NV];
                        node_equations(c, j, yl, yc, yr, s.Yold + j * NV, inv_dt, r, rho, out);
                        const double inv_delta = 1.0 / delta;
#pragma unroll
                        for (int k = 0; k < NV; ++k) blk[k * NV + v] = (out[k] - s.F[j * NV + k]) * inv_delta;
                    }
                    __syncthreads();
                    // ---- twisted block factorisation: the elimination runs from both ends towards the middle node MID,
                    // two independent chains of half the length.  Step t reduces, by Gauss-Jordan with row pivoting
                    // and one thread per element (of both systems),
                    //   top    node t:       [D - A W_{t-1} | C | I] -> [I | W | inv D']    (W replaces C, inv D' replaces D)
                    //   bottom node NX-1-t:  [D - C V_{b+1} | A | I] -> [I | V | inv D'']   (V replaces A, inv D'' replaces D)
                    // The scratch is double-buffered (one barrier per pivot step); the last barrier of a step covers
                    // the write-back of the factors and the assembly of the next two systems, which take W and V
                    // straight from the scratch.  The middle node closes with inv(D - A W_{MID-1} - C V_{MID+1}).
                    const int row = tid / MCOLS, col = tid - row * MCOLS;
                    const bool elem = tid < NV * MCOLS;
                    const int lane = tid & 31;
                    double* Mc[2] = {s.M, s.M + 2 * NV * MROW};
                    double* Mn[2] = {s.M + NV * MROW, s.M + 3 * NV * MROW};
                    if (elem) {
                        const int jb = NX - 1;
                        Mc[0][row * MROW + col] = col < NV ? s.D[row * NV + col]
                                                           : (col < 2 * NV ? s.C[row * NV + (col - NV)]
                                                                           : (col - 2 * NV == row ? 1.0 : 0.0));
                        Mc[1][row * MROW + col] = col < NV ? s.D[jb * NB + row * NV + col]
                                                           : (col < 2 * NV ? s.A[jb * NB + row * NV + (col - NV)]
                                                                           : (col - 2 * NV == row ? 1.0 : 0.0));
                    }
                    __syncthreads();
                    for (int t = 0; t <= MID; ++t) {
                        const bool last = t == MID;          // middle node: one system only
                        const int nsys = last ? 1 : 2;
                        const int jt = t, jb = NX - 1 - t;
                        unsigned used[2] = {0, 0}, rowof[2] = {0, 0};   // every thread tracks the same pivot choices
                        for (int pv = 0; pv < NV; ++pv) {
#pragma unroll
                            for (int y = 0; y < 2; ++y) {
                                if (y < nsys) {
                                    // pivot row: largest magnitude among the unused rows, found by every warp for
                                    // itself from the high words of the seven candidates
                                    unsigned key = 0;
                                    if (lane < NV && !((used[y] >> lane) & 1u))
                                        key = ((unsigned)__double2hiint(Mc[y][lane * MROW + pv]) & 0x7ffffff8u) | (unsigned)lane;
                                    key = __reduce_max_sync(0xffffffffu, key);
                                    int piv = (int)(key & 7u);
                                    if (key < 8u || key >= 0x7ff00000u) {   // singular or non-finite block
                                        piv = 0;
                                        while ((used[y] >> piv) & 1u) ++piv;
                                        if (tid == 0) s.fail = 1;
                                    }
                                    used[y] |= 1u << piv;
                                    rowof[y] |= (unsigned)piv << (3 * pv);
                                    if (elem) {
                                        const double* M0 = Mc[y];
                                        const double a = M0[row * MROW + pv], b2 = M0[piv * MROW + col],
                                                     pvv = M0[piv * MROW + pv], mine = M0[row * MROW + col];
                                        const double bn = b2 * kin::rcp(pvv);
                                        Mn[y][row * MROW + col] = (row == piv) ? bn : mine - a * bn;
                                    }
                                }
                            }
                            __syncthreads();
#pragma unroll
                            for (int y = 0; y < 2; ++y) {
                                double* tmp = Mc[y];
                                Mc[y] = Mn[y];
                                Mn[y] = tmp;
                            }
                        }
                        // Mc[y] = [I | W or V | inverse] with unknown pu in row rowof[y][pu]
                        if (tid < NV * 14) {
                            const int pu = tid / 14, cc = tid - pu * 14;
                            const double vt = Mc[0][((rowof[0] >> (3 * pu)) & 7u) * MROW + NV + cc];
                            if (last) {
                                if (cc >= NV) s.D[MID * NB + pu * NV + (cc - NV)] = vt;
                            } else {
                                const double vb = Mc[1][((rowof[1] >> (3 * pu)) & 7u) * MROW + NV + cc];
                                if (cc < NV) {
                                    s.C[jt * NB + pu * NV + cc] = vt;
                                    s.A[jb * NB + pu * NV + cc] = vb;
                                } else {
                                    s.D[jt * NB + pu * NV + (cc - NV)] = vt;
                                    s.D[jb * NB + pu * NV + (cc - NV)] = vb;
                                }
                            }
                        }
                        if (elem && !last) {
                            const int nt = jt + 1, nbm = jb - 1;   // next top / bottom nodes; they coincide at MID
                            const bool mid_next = nt == MID;
                            double vt, vb = 0.0;
                            if (col < NV) {
                                vt = s.D[nt * NB + row * NV + col];
#pragma unroll
                                for (int k = 0; k < NV; ++k)
                                    vt -= s.A[nt * NB + row * NV + k] * Mc[0][((rowof[0] >> (3 * k)) & 7u) * MROW + NV + col];
                                if (mid_next) {
#pragma unroll
                                    for (int k = 0; k < NV; ++k)
                                        vt -= s.C[nt * NB + row * NV + k] * Mc[1][((rowof[1] >> (3 * k)) & 7u) * MROW + NV + col];
                                } else {
                                    vb = s.D[nbm * NB + row * NV + col];
#pragma unroll
                                    for (int k = 0; k < NV; ++k)
                                        vb -= s.C[nbm * NB + row * NV + k] * Mc[1][((rowof[1] >> (3 * k)) & 7u) * MROW + NV + col];
                                }
                            } else if (col < 2 * NV) {
                                vt = mid_next ? 0.0 : s.C[nt * NB + row * NV + (col - NV)];
                                vb = mid_next ? 0.0 : s.A[nbm * NB + row * NV + (col - NV)];
                            } else {
                                vt = vb = (col - 2 * NV == row) ? 1.0 : 0.0;
                            }
                            Mn[0][row * MROW + col] = vt;
                            Mn[1][row * MROW + col] = vb;
                        }
                        __syncthreads();
#pragma unroll
                        for (int y = 0; y < 2; ++y) {
                            double* tmp = Mc[y];
                            Mc[y] = Mn[y];
                            Mn[y] = tmp;
                        }
                    }
                    if (s.fail) {
                        bad = true;
                        break;
                    }
                    // L_j = inv D'_j A_j replaces A_j above the middle, U_j = inv D''_j C_j replaces C_j below it: each
                    // forward recurrence is then one 7x7 product per node
                    for (int j0 = 0; j0 < NX; j0 += 13) {   // 13 whole nodes (637 products) per pass, four per thread
                        const int e_end = (j0 + 13 < NX ? j0 + 13 : NX) * NB;
                        double lv[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int e = j0 * NB + tid + i * DAE_THREADS;
                            double acc = 0.0;
                            if (e < e_end) {
                                const int j = e / NB, rc = e - j * NB, pu = rc / NV, cc = rc - pu * NV;
                                const double* src = (j < MID ? s.A : s.C) + j * NB;
                                for (int k = 0; k < NV; ++k) acc += s.D[j * NB + pu * NV + k] * src[k * NV + cc];
                            }
                            lv[i] = acc;
                        }
                        __syncthreads();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int e = j0 * NB + tid + i * DAE_THREADS;
                            if (e < e_end) {
                                const int j = e / NB;
                                if (j < MID) s.A[e] = lv[i];
                                else if (j > MID) s.C[e] = lv[i];
                            }
                        }
                        __syncthreads();
                    }
                }
                // ---- substitution with the stored factors.  h_j = -inv D_j F_j for all nodes at once (the raw -F at
                // the middle); then warp 0 runs g_j = h_j - L_j g_{j-1} down from node 0 while warp 1 runs
                // q_j = h_j - U_j q_{j+1} up from node NX-1, seven lanes each, the running vector in registers and
                // exchanged by shuffles; the middle node closes the system; then x_j = g_j - W_j x_{j+1} and
                // x_j = q_j - V_j x_{j-1} run back outwards.  F ends up holding the Newton update.
                double* const H = s.M;
                static_assert(4 * NV * MROW >= NX * NV, "h does not fit the elimination scratch");
                for (int e = tid; e < NX * NV; e += DAE_THREADS) {
                    const int j = e / NV, pu = e - j * NV;
                    double acc = 0.0;
                    if (j == MID) acc = -s.F[e];
                    else
                        for (int k = 0; k < NV; ++k) acc -= s.D[j * NB + pu * NV + k] * s.F[j * NV + k];
                    H[e] = acc;
                }
                __syncthreads();
                {
                    const int lane = tid & 31, ln = lane < NV ? lane : 0, w = tid >> 5;
                    if (w == 0) {
                        double g = H[ln];
                        if (lane < NV) s.F[ln] = g;
                        for (int j = 1; j < MID; ++j) {
                            const double* Lr = s.A + j * NB + ln * NV;
                            double acc = H[j * NV + ln];
#pragma unroll
                            for (int k = 0; k < NV; ++k) acc -= Lr[k] * __shfl_sync(0xffffffffu, g, k);
                            g = acc;
                            if (lane < NV) s.F[j * NV + ln] = g;
                        }
                    } else if (w == 1) {
                        double q = H[(NX - 1) * NV + ln];
                        if (lane < NV) s.F[(NX - 1) * NV + ln] = q;
                        for (int j = NX - 2; j > MID; --j) {
                            const double* Ur = s.C + j * NB + ln * NV;
                            double acc = H[j * NV + ln];
#pragma unroll
                            for (int k = 0; k < NV; ++k) acc -= Ur[k] * __shfl_sync(0xffffffffu, q, k);
                            q = acc;
                            if (lane < NV) s.F[j * NV + ln] = q;
                        }
                    }
                    __syncthreads();
                    if (w == 0) {   // middle node: x = inv(Dm) (b - A g_{MID-1} - C q_{MID+1})
                        double tmpv = 0.0;
                        if (lane < NV) {
                            tmpv = H[MID * NV + ln];
                            for (int k = 0; k < NV; ++k)
                                tmpv -= s.A[MID * NB + ln * NV + k] * s.F[(MID - 1) * NV + k] +
                                        s.C[MID * NB + ln * NV + k] * s.F[(MID + 1) * NV + k];
                        }
                        double xm = 0.0;
#pragma unroll
                        for (int k = 0; k < NV; ++k) xm += s.D[MID * NB + ln * NV + k] * __shfl_sync(0xffffffffu, tmpv, k);
                        if (lane < NV) s.F[MID * NV + ln] = xm;
                    }
                    __syncthreads();
                    if (w == 0) {
                        double x = s.F[MID * NV + ln];
                        for (int j = MID - 1; j >= 0; --j) {
                            const double* Wr = s.C + j * NB + ln * NV;
                            double acc = s.F[j * NV + ln];
#pragma unroll
                            for (int k = 0; k < NV; ++k) acc -= Wr[k] * __shfl_sync(0xffffffffu, x, k);
                            x = acc;
                            if (lane < NV) s.F[j * NV + ln] = x;
                        }
                    } else if (w == 1) {
                        double x = s.F[MID * NV + ln];
                        for (int j = MID + 1; j < NX; ++j) {
                            const double* Vr = s.A + j * NB + ln * NV;
                            double acc = s.F[j * NV + ln];
#pragma unroll
                            for (int k = 0; k < NV; ++k) acc -= Vr[k] * __shfl_sync(0xffffffffu, x, k);
                            x = acc;
                            if (lane < NV) s.F[j * NV + ln] = x;
                        }
                    }
                }
                __syncthreads();
                // ---- update and convergence test
                double worst = 0.0;
                bool fin = true;
                for (int e = tid; e < NX * NV; e += DAE_THREADS) {
                    const int v = e % NV;
                    const double dy = s.F[e], y = s.Y[e] + dy;
                    s.Y[e] = y;
                    fin = fin && isfinite(y);
                    worst = fmax(worst, fabs(dy) / (fabs(y) + floor_of(v)));
                }
                if (!fin) worst = INFINITY;
                worst = warp_max(worst);
                if ((tid & 31) == 0) s.red[tid >> 5] = worst;
                __syncthreads();
                worst = 0.0;
                for (int w = 0; w < DAE_THREADS / 32; ++w) worst = fmax(worst, s.red[w]);
                __syncthreads();
                if (!(worst < INFINITY)) bad = true;
                else if (worst < NEWTON_TOL) {
                    converged = true;
                    // nothing moved over a whole, repeated grid step: the state is steady, the rest are no-ops
                    if (it == 0 && whole_repeat) steady = true;
                }
                // the factors are kept for the next iteration unless the update stopped shrinking (modified Newton)
                else if (worst > 0.3 * prev_worst) need_jac = true;
                prev_worst = worst;
            }
            if (converged) {
                t_left -= dt;
                dt *= 2.0;
                if (!(t_left > 1e-12 * H)) {   // grid step complete
                    H_prev = H;
                    n_retry = 0;
                    if (++step >= n_dt) break;
                    H = dts[step];
                    t_left = H;
                    dt = H;
                }
            } else {   // back to the state before the piece, smaller piece, fresh factors
                __syncthreads();
                for (int e = tid; e < NX * NV; e += DAE_THREADS) s.Y[e] = s.Yold[e];
                if (tid == 0) s.fail = 0;
                __syncthreads();
                dt_fac = 0.0;
                dt *= 0.25;
                if (++n_retry > MAX_RETRY) failed = true;
            }
        }
        if (tid == 0) {
            // outlet flows (set_likelihood.py:204-208), -10000 on failure (:244), squared residuals of the five species
            const double* yo = s.Y + (NX - 1) * NV;
            const double T = yo[5], u = yo[6];
            double F[5];
            bool ok = !failed;
            for (int k = 0; k < 5; ++k) {
                F[k] = yo[k] * kin::S_TUBE * u * 60 * kin::R_GAS * T / c.P0 * 1e6 * c.P0 / kin::P_STP * 298 / T;
                ok = ok && isfinite(F[k]);
            }
            double acc = 0.0;
            for (int k = 0; k < 5; ++k) {
                const double f = ok ? F[k] : kin::FAIL_FLOW;
                const double rsd = f - obs[(int64_t)k * n_cond + ci];
                acc += rsd * rsd;
            }
            ssr[(int64_t)ci * n + p] = acc;
        }
    }
}

}  // namespace

int launch_loglik_dae(smcb_handle* h, const double* theta, int64_t ld, int64_t n, int d, const uint8_t* active,
                      double* lk, cudaStream_t st) {
    const KineticData& D = h->kin;
    REQUIRE(h, D.cond != nullptr, SMCB_ERR_STATE, "smcb_set_data_kinetic has not been called");
    REQUIRE(h, D.n_pairs == 4, SMCB_ERR_UNSUPPORTED, "the transient reactor model has the reference's 8 kinetic parameters");
    REQUIRE(h, d == D.d, SMCB_ERR_INVALID, "d differs from the d given to smcb_set_data_kinetic");
    if (n == 0) return SMCB_OK;
    REQUIRE(h, h->ssr != nullptr && n <= h->n_max && D.n_cond <= h->ssr_rows, SMCB_ERR_STATE,
            "smcb_reserve too small for this sweep");
    if (!h->dae_dts) {   // fixed time grid: 1 ms growing by 1.5x up to 5 s, last step clipped at 75 s
        std::vector<double> dts;
        double t = 0.0, dt = 1e-3;
        while (t < 75.0) {
            const double k = dt < 75.0 - t ? dt : 75.0 - t;
            dts.push_back(k);
            t += k;
            dt = dt * 1.5 < 5.0 ? dt * 1.5 : 5.0;
        }
        CUDA_TRY(h, cudaMalloc((void**)&h->dae_dts, sizeof(double) * dts.size()));
        CUDA_TRY(h, cudaMemcpy(h->dae_dts, dts.data(), sizeof(double) * dts.size(), cudaMemcpyHostToDevice));
        h->dae_n_dt = (int)dts.size();
        CUDA_TRY(h, cudaFuncSetAttribute(dae_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(Smem)));
    }
    unsigned* list = nullptr;
    unsigned* count = nullptr;
    int rc = kinetic_pack_active(h, active, n, st, &list, &count);
    if (rc != SMCB_OK) return rc;
    int64_t blocks = n * (int64_t)D.n_cond;
    if (blocks > (int64_t)h->sm_count * 3) blocks = (int64_t)h->sm_count * 3;
    dae_march_kernel<<<(unsigned)blocks, DAE_THREADS, sizeof(Smem), st>>>(theta, ld, n, list, count, D.cond, D.obs,
                                                                         D.n_cond, D.base, D.est_pos, h->dae_dts,
                                                                         h->dae_n_dt, h->ssr);
    LAUNCH_CHECK(h);
    return kinetic_finalize(h, theta, ld, n, active, lk, st);
}
