// Methanation-style kinetic / reactor model, device side.
//
// Physics restated from the reference (SMC_methanation/methanation_set_likelihood.py):
//   rate law      func_rCH4  :44-58   Langmuir-Hinshelwood forward/reverse Sabatier rate
//   gas density   func_rohg  :61-66
//   outlet flows  my_model   :204-208 [sccm]
//   log-lik       my_loglike :289-298 (no 2*pi term), failure penalty -10000 (:244)
// and constants from methanation_set_conditon.py:74-89.
//
// Reactor (builder-defined): the reference integrates a 357-unknown transient
// DAE with SUNDIALS IDA; here the steady plug-flow limit of the same balances (axial dispersion
// and conduction dropped) is marched along z with classical fixed-step RK4:
//     d(u C_k)/dz     = (1-void) sc_k r                       k = H2, CO2, CH4, H2O, Ar
//     d(u P0/(R T))/dz= -2 (1-void) r                         (continuity at constant pressure)
//     rho_g Cpg d(T u)/dz = (1-void)(-Hr) r - (2U/dint)(T - T_jacket)
// With one reaction the species fluxes are N_k = N_k0 + sc_k*xi, so the state is (xi, G=T*u):
//     T = sqrt(G P0 / (R sum_k N_k)),  u = G/T,  C_k = N_k/u.
// The rate is a sum of M Langmuir-Hinshelwood channels, each with the reference's four Arrhenius
// pairs (kf, ks, kCO2, kH2O); M=1 is exactly func_rCH4, M=4 gives the 32-parameter family.
#pragma once
#include "common.cuh"
#include "exp_table.cuh"

namespace kin {

constexpr double R_GAS = 8.3144589;       // J/mol/K    (set_conditon.py:79)
constexpr double RHOS_CAT = 5075.0;       // kg/m3      (:77) -- enters as 5075e3 in the rate law
constexpr double HR = -164940.0;          // J/mol      (:78)
constexpr double CPG = 2800.0;            // J/kg/K     (:82)
constexpr double U_WALL = 68.2480;        // W/m2/K     (:86)
constexpr double DINT = 0.005;            // m          (:85)
constexpr double P_STP = 1.013e5;         // Pa         (:89)
constexpr double RR_TUBE = 0.01 / 2;      // m          (:80)
constexpr double S_TUBE = M_PI * RR_TUBE * RR_TUBE;   // m2 (:81)
constexpr double FAIL_FLOW = -10000.0;    // set_likelihood.py:244

constexpr int MAX_PAIRS = 16;

// 1/x and sqrt(x) from the MUFU seeds (20 bits) and one cubically convergent correction each: 4 and 7 FP64-pipe
// operations against ~20 for the IEEE division / square root sequences.  The right-hand side below has nine
// quotients and two roots per evaluation; with true divisions they were 60% of the kernel's instructions.
// Error < 1 ulp (reciprocal) / < 1.5 ulp (root), i.e. the size of the rounding differences between NumPy's and
// CUDA's exp(); parity with the oracle is asserted at 1e-9 on well-conditioned particles.
__device__ __forceinline__ double rcp(double x) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    const double e = fma(-x, r0, 1.0);
    return fma(r0, fma(e, e, e), r0);
}
__device__ __forceinline__ double sqrt_fast(double x) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double xy = x * y0;
    const double e = fma(-xy, y0, 1.0);                     // 1 - x*y0^2
    return fma(xy, e * fma(0.375, e, 0.5), xy);             // x*y0*(1 + e/2 + 3e^2/8)
}


struct Cond {
    double N0[5];     // inlet molar fluxes u_in*C_k_in
    double P0;        // total pressure (sum C_in) R T_in
    double G0;        // T_in*u_in
    double Tj;        // jacket temperature
    double omv;       // 1 - void
    double dz;        // length / n_steps
};

__device__ __forceinline__ Cond load_cond(const double* __restrict__ c, int n_steps) {
    Cond o;
    const double Tin = c[5], uin = c[7];
    double csum = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        o.N0[k] = uin * c[k];
        csum += c[k];
    }
    o.P0 = csum * R_GAS * Tin;
    o.G0 = Tin * uin;
    o.Tj = c[6];
    o.omv = 1.0 - c[8];
    o.dz = c[9] / n_steps;
    return o;
}

// per-particle kinetic constants: A_j and -E_j/R for 4*M Arrhenius pairs; inv_t_lim = 708 / max_j |E_j/R|: for
// |1/T| below it every Arrhenius exponent is inside the range the table exp needs no checks for
template <int M>
struct Kin {
    double A[4 * M];
    double nEoR[4 * M];
    double inv_t_lim;
    __device__ __forceinline__ void set_limit() {
        double mx = 0.0;
#pragma unroll
        for (int j = 0; j < 4 * M; ++j) mx = fmax(mx, fabs(nEoR[j]));
        inv_t_lim = 708.0 / mx;   // +inf when every E is zero; NaN parameters give NaN, which sends rate() to the checked path
    }
};

template <int M>
__device__ __forceinline__ double rate(const Kin<M>& K, double T, double Ca, double Cb, double Cc, double Cd,
                                       const double* __restrict__ etab) {
    const double RT6 = R_GAS * T * 1e-6;
    const double PH2 = Ca * RT6, PCO2 = Cb * RT6, PCH4 = Cc * RT6, PH2O = Cd * RT6;
    const double sH2 = sqrt_fast(fmax(0.001, PH2));
    const double invT = rcp(T);
    double r = 0.0;
    // M = 4: one comparison instead of a range test per factor (7% faster); with M = 1 the second code path costs
    // registers the march needs (65 -> 104 ms), so the single channel keeps the per-factor selects
    if (M > 1 && fabs(invT) < K.inv_t_lim) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const double kf = K.A[4 * m + 0] * expt::exp_core(K.nEoR[4 * m + 0] * invT, etab);
            const double ks = K.A[4 * m + 1] * expt::exp_core(K.nEoR[4 * m + 1] * invT, etab);
            const double kC = K.A[4 * m + 2] * expt::exp_core(K.nEoR[4 * m + 2] * invT, etab);
            const double kW = K.A[4 * m + 3] * expt::exp_core(K.nEoR[4 * m + 3] * invT, etab);
            const double dC = 1.0 + kC * PCO2, dW = 1.0 + kW * PH2O;
            const double rf = 5075e3 * kf * kC * PCO2 * sH2 * rcp(dC * dC);
            const double rr = 5075e3 * ks * kW * PH2O * (PCH4 * PCH4) * rcp(dW * dW);
            r += rf - rr;
        }
    } else {   // (unrolled as well: a rolled loop would index K dynamically and push it to local memory)
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const double kf = K.A[4 * m + 0] * expt::exp_fast(K.nEoR[4 * m + 0] * invT, etab);
            const double ks = K.A[4 * m + 1] * expt::exp_fast(K.nEoR[4 * m + 1] * invT, etab);
            const double kC = K.A[4 * m + 2] * expt::exp_fast(K.nEoR[4 * m + 2] * invT, etab);
            const double kW = K.A[4 * m + 3] * expt::exp_fast(K.nEoR[4 * m + 3] * invT, etab);
            const double dC = 1.0 + kC * PCO2, dW = 1.0 + kW * PH2O;
            const double rf = 5075e3 * kf * kC * PCO2 * sH2 * rcp(dC * dC);
            const double rr = 5075e3 * ks * kW * PH2O * (PCH4 * PCH4) * rcp(dW * dW);
            r += rf - rr;
        }
    }
    return r;
}

struct Local {
    double T, u, C[5];
};

__device__ __forceinline__ Local local_state(const Cond& c, double xi, double G) {
    Local s;
    const double Na = c.N0[0] - 4.0 * xi, Nb = c.N0[1] - xi, Nc = c.N0[2] + xi, Nd = c.N0[3] + 2.0 * xi,
                 Ne = c.N0[4];
    const double Ns = Na + Nb + Nc + Nd + Ne;
    s.T = sqrt_fast(G * c.P0 * rcp(R_GAS * Ns));
    const double iT = rcp(s.T);
    s.u = G * iT;
    const double iu = rcp(s.u);
    s.C[0] = Na * iu; s.C[1] = Nb * iu; s.C[2] = Nc * iu; s.C[3] = Nd * iu; s.C[4] = Ne * iu;
    return s;
}

template <int M>
__device__ __forceinline__ void rhs(const Kin<M>& K, const Cond& c, double xi, double G, double* dxi, double* dG,
                                    const double* __restrict__ etab) {
    const Local s = local_state(c, xi, G);
    const double r = rate<M>(K, s.T, s.C[0], s.C[1], s.C[2], s.C[3], etab);
    const double csum = s.C[0] + s.C[1] + s.C[2] + s.C[3] + s.C[4];
    const double rho = c.P0 / R_GAS * rcp(s.T) *
                       (s.C[0] * 2 + s.C[1] * 44 + s.C[2] * 16 + s.C[3] * 18 + s.C[4] * 40) * rcp(csum) * 0.001;
    *dxi = c.omv * r;
    *dG = (c.omv * (-HR) * r - 2 * U_WALL / DINT * (s.T - c.Tj)) * rcp(rho * CPG);
}

// integrate one operating condition; returns sum_k (F_k - obs_k)^2 over the five species
template <int M>
__device__ __forceinline__ double condition_ssr(const Kin<M>& K, const double* __restrict__ cond_row, int n_steps,
                                                const double* __restrict__ obs, int n_cond, int ci,
                                                const double* __restrict__ etab) {
    const Cond c = load_cond(cond_row, n_steps);
    double xi = 0.0, G = c.G0;
    const double h = c.dz;
    for (int s = 0; s < n_steps; ++s) {
        double a1, b1, a2, b2, a3, b3, a4, b4;
        rhs<M>(K, c, xi, G, &a1, &b1, etab);
        rhs<M>(K, c, xi + 0.5 * h * a1, G + 0.5 * h * b1, &a2, &b2, etab);
        rhs<M>(K, c, xi + 0.5 * h * a2, G + 0.5 * h * b2, &a3, &b3, etab);
        rhs<M>(K, c, xi + h * a3, G + h * b3, &a4, &b4, etab);
        xi += h / 6.0 * (a1 + 2.0 * a2 + 2.0 * a3 + a4);
        G += h / 6.0 * (b1 + 2.0 * b2 + 2.0 * b3 + b4);
    }
    const Local s = local_state(c, xi, G);
    double F[5];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        // y*S*u*60*R*T/P*1e6*P/P_stp*298/T   (set_likelihood.py:204-208)
        F[k] = s.C[k] * S_TUBE * s.u * 60 * R_GAS * s.T / c.P0 * 1e6 * c.P0 / P_STP * 298 / s.T;
        ok = ok && isfinite(F[k]);
    }
    double ssr = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double f = ok ? F[k] : FAIL_FLOW;
        const double r = f - obs[(int64_t)k * n_cond + ci];
        ssr += r * r;
    }
    return ssr;
}

}  // namespace kin
