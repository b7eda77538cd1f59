"""Methanation-style kinetic/reactor model on the CPU (oracle; test infrastructure only).

PARITY UNPINNED against the reference's own forward model: the reference integrates a transient
357-unknown DAE with SUNDIALS IDA through `assimulo` (`methanation_set_likelihood.py:69-277`);
neither assimulo nor the operating-conditions file `methanation_data/information.csv` is available.
What *is* restated exactly:
    rate law      func_rCH4  `methanation_set_likelihood.py:44-58`
    gas density   func_rohg  `:61-66`
    outlet flows             `:204-208`
    log-lik       my_loglike `:289-298`;  failure penalty -10000 `:244`
    constants                `methanation_set_conditon.py:74-89`
The reactor integration is the builder-defined steady plug-flow RK4 march documented in DESIGN.md
and in `csrc/kinetic.cuh`; this file is its NumPy twin (vectorised over particles).
"""
import numpy as np

R = 8.3144589
HR = -164940.0
CPG = 2800.0
U_WALL = 68.2480
DINT = 0.005
P_STP = 1.013 * 10 ** 5
RR = 0.01 / 2
S_TUBE = np.pi * RR ** 2
FAIL_FLOW = -10000.0

BASEPARAMS = np.array([13.04, 52.2e3, 1.147e5, 96.7e3, 23.34, -6, 0.72, -2.51e3])   # set_conditon.py:56-57
SIGMA_TRUE = 5.0
HIGH_K = np.array([25, 1, 30, 2, 1, -2, 1, -2, 2], dtype=float)                      # :62-63
LOW_K = np.array([4, 1, 4, 1, 1, -2, 1, -2, 0.9], dtype=float)
EST_POSITION = [0, 1, 2, 3, 8]                                                       # :19,34


def reference_box():
    """Prior box of the reference for the estimated positions (`set_conditon.py:64-68`)."""
    use = np.append(BASEPARAMS, SIGMA_TRUE)
    high = use + use * HIGH_K
    low = use - use * LOW_K
    return low[EST_POSITION], high[EST_POSITION]


def synthetic_conditions(n_cond=30, seed=20250205):
    """Builder-chosen operating conditions (the reference's input file is missing).
    Columns: Ca,Cb,Cc,Cd,Ce [mol/m3], T_in, T_jacket [K], u_in [m/s], void, length [m]."""
    rs = np.random.RandomState(seed)
    # mild, Ar-diluted conditions: the steady plug-flow energy balance has no solid thermal mass,
    # so undiluted feeds run away thermally (adiabatic rise of the Sabatier reaction is ~900 K)
    T = rs.uniform(493.0, 553.0, n_cond)
    Tj = T + rs.uniform(-3.0, 3.0, n_cond)
    P = rs.uniform(0.1, 0.5, n_cond) * 1e6 + 101325.0
    ratio = rs.uniform(4.2, 6.0, n_cond)                 # H2:CO2 (H2 in excess: the rate law clamps P_H2 at 0.001 MPa)
    x_ar = rs.uniform(0.70, 0.90, n_cond)
    x_co2 = (1 - x_ar) / (1 + ratio)
    x_h2 = x_co2 * ratio
    ctot = P / R / T
    sccm = rs.uniform(200.0, 1000.0, n_cond)
    u = sccm * 1.667e-8 / S_TUBE * (101325 * T) / (P * 298)      # set_conditon.py:190
    void = np.full(n_cond, 0.4)
    L = rs.uniform(0.004, 0.03, n_cond)
    z = np.zeros(n_cond)
    return np.stack([ctot * x_h2, ctot * x_co2, z, z, ctot * x_ar, T, Tj, u, void, L], axis=1)


def _rate(A, nEoR, T, Ca, Cb, Cc, Cd):
    """Sum of M Langmuir-Hinshelwood channels; A, nEoR: [n, 4M]."""
    RT6 = R * T * 1e-6
    PH2, PCO2, PCH4, PH2O = Ca * RT6, Cb * RT6, Cc * RT6, Cd * RT6
    sH2 = np.sqrt(np.maximum(0.001, PH2))
    invT = 1.0 / T
    r = np.zeros_like(T)
    for m in range(A.shape[1] // 4):
        kf = A[:, 4 * m + 0] * np.exp(nEoR[:, 4 * m + 0] * invT)
        ks = A[:, 4 * m + 1] * np.exp(nEoR[:, 4 * m + 1] * invT)
        kC = A[:, 4 * m + 2] * np.exp(nEoR[:, 4 * m + 2] * invT)
        kW = A[:, 4 * m + 3] * np.exp(nEoR[:, 4 * m + 3] * invT)
        dC, dW = 1.0 + kC * PCO2, 1.0 + kW * PH2O
        rf = 5075e3 * kf * kC * PCO2 * sH2 / (dC * dC)
        rr = 5075e3 * ks * kW * PH2O * (PCH4 * PCH4) / (dW * dW)
        r = r + (rf - rr)
    return r


def _local(c, xi, G):
    N0 = c["N0"]
    Na, Nb, Nc, Nd, Ne = N0[0] - 4.0 * xi, N0[1] - xi, N0[2] + xi, N0[3] + 2.0 * xi, N0[4] + 0.0 * xi
    Ns = Na + Nb + Nc + Nd + Ne
    T = np.sqrt(G * c["P0"] / (R * Ns))
    u = G / T
    iu = 1.0 / u
    return T, u, (Na * iu, Nb * iu, Nc * iu, Nd * iu, Ne * iu)


def _rhs(A, nEoR, c, xi, G):
    T, u, C = _local(c, xi, G)
    r = _rate(A, nEoR, T, C[0], C[1], C[2], C[3])
    csum = C[0] + C[1] + C[2] + C[3] + C[4]
    rho = c["P0"] / R / T * (C[0] * 2 + C[1] * 44 + C[2] * 16 + C[3] * 18 + C[4] * 40) / csum * 0.001
    dxi = c["omv"] * r
    dG = (c["omv"] * (-HR) * r - 2 * U_WALL / DINT * (T - c["Tj"])) / (rho * CPG)
    return dxi, dG


def outlet_flows(full, cond, n_steps=50):
    """full: [n, 8M+1] full parameter vectors; cond: [n_cond, 10].  Returns F[n, 5, n_cond] sccm."""
    full = np.atleast_2d(np.asarray(full, dtype=np.float64))
    n = full.shape[0]
    A = full[:, 0:-1:2]
    nEoR = -full[:, 1:-1:2] / R
    out = np.empty((n, 5, cond.shape[0]))
    with np.errstate(all="ignore"):
        for ci, row in enumerate(cond):
            Tin, uin = row[5], row[7]
            c = dict(N0=uin * row[:5], P0=np.sum(row[:5]) * R * Tin, Tj=row[6], omv=1.0 - row[8])
            h = row[9] / n_steps
            xi = np.zeros(n)
            G = np.full(n, Tin * uin)
            for _ in range(n_steps):
                a1, b1 = _rhs(A, nEoR, c, xi, G)
                a2, b2 = _rhs(A, nEoR, c, xi + 0.5 * h * a1, G + 0.5 * h * b1)
                a3, b3 = _rhs(A, nEoR, c, xi + 0.5 * h * a2, G + 0.5 * h * b2)
                a4, b4 = _rhs(A, nEoR, c, xi + h * a3, G + h * b3)
                xi = xi + h / 6.0 * (a1 + 2.0 * a2 + 2.0 * a3 + a4)
                G = G + h / 6.0 * (b1 + 2.0 * b2 + 2.0 * b3 + b4)
            T, u, C = _local(c, xi, G)
            F = np.stack([Ck * S_TUBE * u * 60 * R * T / c["P0"] * 1e6 * c["P0"] / P_STP * 298 / T for Ck in C],
                         axis=1)
            bad = ~np.all(np.isfinite(F), axis=1)
            F[bad] = FAIL_FLOW
            out[:, :, ci] = F
    return out


def assemble(theta, base, est_pos):
    """`p_pred_bases[:, est_position] = particle` (methanation_functions.py:80)."""
    theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    full = np.tile(np.asarray(base, dtype=np.float64), (theta.shape[0], 1))
    full[:, list(est_pos)] = theta
    return full


def loglik(theta, cond, obs, base, est_pos, n_steps=50):
    """my_loglike over all five species (`set_likelihood.py:289-298`); theta: [n, d] -> lk[n]."""
    full = assemble(theta, base, est_pos)
    sigma = full[:, -1]
    F = outlet_flows(full, cond, n_steps)
    n_cond = cond.shape[0]
    with np.errstate(all="ignore"):
        ssr = np.sum((F - obs[None, :, :]) ** 2, axis=(1, 2))
        return -(0.5 / sigma ** 2) * ssr - 5.0 * n_cond * np.log(sigma)


def synthetic_observations(cond, base, sigma=SIGMA_TRUE, n_steps=50, seed=20250205):
    """data = model(baseparams) + N(0, sigma^2) as in SMC_methanation_main.py:89-95."""
    rs = np.random.RandomState(seed + 1)
    F = outlet_flows(np.asarray(base)[None, :], cond, n_steps)[0]
    return F + sigma * rs.standard_normal(F.shape)


def base_vector(n_pairs=4, seed=7):
    """M=1: the reference's baseparams + sigma.  M=4 (32 kinetic parameters): channel 0 is the
    reference's, channels 1..3 are perturbed copies scaled down so the total rate stays comparable."""
    if n_pairs == 4:
        return np.append(BASEPARAMS, SIGMA_TRUE)
    rs = np.random.RandomState(seed)
    chans = []
    for m in range(n_pairs // 4):
        p = BASEPARAMS.copy()
        if m > 0:
            p[0::2] *= rs.uniform(0.6, 1.4, 4)       # pre-exponentials
            p[1::2] *= rs.uniform(0.9, 1.1, 4)       # activation energies
        p[0] *= 1.0 / (n_pairs // 4)                 # share the forward rate between channels
        p[2] *= 1.0 / (n_pairs // 4)
        chans.append(p)
    return np.append(np.concatenate(chans), SIGMA_TRUE)
