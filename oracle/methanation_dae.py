"""Transient fixed-bed methanation reactor of the reference, on the CPU (oracle; test infrastructure only).

SURVEY.md 8(f) N3.  The reference integrates, per particle and operating condition, a 7 x 51 = 357-unknown
index-1 DAE with SUNDIALS IDA through `assimulo` (`SMC_methanation/methanation_set_likelihood.py:69-139`
`reaction`, `:144-277` `my_model`).  Restated here:

    residual      `reaction`   :69-139   (pinned: tests/golden/methanation_dae_residual.npz holds residuals computed by
                                          the reference's own function text, see tests/golden/make_dae_fixture.py)
    rate, density `func_rCH4`  :44-58, `func_rohg` :61-66
    start state   SMC_methanation.py:412-423 (feed composition everywhere, bed at 400 K behind the inlet node)
    end time      75 s (:198), outlet flows :204-208, failure penalty -10000 (:244)
    constants     methanation_set_conditon.py:74-89, NX = 51 (:44)

PARITY OF THE TIME INTEGRATION IS UNPINNED: assimulo / IDA (variable-order BDF) is not installed and the
reference stores no trajectories.  The integrator below is builder-defined and is what the device kernel
(`csrc/dae.cu`) twins: implicit Euler (BDF1) on a fixed geometric time grid from 0 to 75 s, modified Newton (a
finite-difference block-tridiagonal Jacobian kept while the update shrinks by 0.3x per iteration); a grid step whose
Newton iteration does not converge is retried in smaller pieces, and a march that still cannot advance fails
(-10000 flows, the reference's own penalty for IDA failures).  The bed's thermal time constant is ~8 s
((1-void) rho_s Cps 0.1 / (2U/dint)), so at 75 s the state is the steady state of the reference's discretised
balances to ~1e-4, which is what both integrators converge to.

Unknowns are kept as Y[7, NX] = rows (C_H2, C_CO2, C_CH4, C_H2O, C_Ar, T, u), which is the reference's variable-major
vector X[7*NX] reshaped (`to_reference_order` / `from_reference_order`).
"""
import numpy as np
from scipy.linalg import solve_banded

from . import kinetic

NX = 51                      # set_conditon.py:44
R = kinetic.R
SC = np.array([-4.0, -1.0, 1.0, 2.0, 0.0])   # :75
DZ_DISP = 0.95e-5            # Dz, m2/s :76
RHOS = 5075.0                # :77
HR = kinetic.HR              # :78
CPG = kinetic.CPG            # :82
CPS = 698.0                  # :83
KEFF = 0.72                  # :84
DINT = kinetic.DINT          # :85
U_WALL = kinetic.U_WALL      # :86
T_BED0 = 400.0               # SMC_methanation.py:421
T_FINAL = 75.0               # set_likelihood.py:198

# builder-defined time grid and Newton controls (twinned by csrc/dae.cu)
DT0, DT_GROW, DT_MAX = 1e-3, 1.5, 5.0
NEWTON_MAX, NEWTON_TOL = 40, 1e-10
MAX_RETRY = 8                # failed pieces allowed within one grid step
FLOOR = np.array([1e-3, 1e-3, 1e-3, 1e-3, 1e-3, 1.0, 1e-4])     # scale floors of (C x5, T, u)
FD_REL = 1e-7


def time_grid():
    """Step sizes: 1 ms growing by 1.5x up to 5 s, last step clipped at 75 s."""
    dts, t, dt = [], 0.0, DT0
    while t < T_FINAL:
        dt_k = min(dt, T_FINAL - t)
        dts.append(dt_k)
        t += dt_k
        dt = min(dt * DT_GROW, DT_MAX)
    return np.array(dts)


def to_reference_order(Y):
    return np.asarray(Y).reshape(7 * NX)


def from_reference_order(X):
    return np.asarray(X, dtype=np.float64).reshape(7, NX).copy()


def rate(T, Ca, Cb, Cc, Cd, k8):
    """func_rCH4 (:44-58); k8 = (A_f, E_f, A_s, E_s, A_CO2, E_CO2, A_H2O, E_H2O)."""
    RT6 = R * T * 1e-6
    PH2, PCO2, PCH4, PH2O = Ca * RT6, Cb * RT6, Cc * RT6, Cd * RT6
    kf = k8[0] * np.exp(-k8[1] / R / T)
    ks = k8[2] * np.exp(-k8[3] / R / T)
    kC = k8[4] * np.exp(-k8[5] / R / T)
    kW = k8[6] * np.exp(-k8[7] / R / T)
    rf = 5075e3 * kf * kC * PCO2 * np.sqrt(np.maximum(0.001, PH2)) / (1 + kC * PCO2) ** 2
    rr = 5075e3 * ks * kW * PH2O * PCH4 ** 2 / (1 + kW * PH2O) ** 2
    return rf - rr


def density(C, T, P0):
    """func_rohg (:61-66); C: [5, ...]."""
    return P0 / R / T * (C[0] * 2 + C[1] * 44 + C[2] * 16 + C[3] * 18 + C[4] * 40) / (C[0] + C[1] + C[2] + C[3] + C[4]) * 0.001


def residual(Y, dY, cond_row, k8):
    """F(Y, dY) of `reaction` (:69-139), node-major: rows 0..4 species balances, row 5 the equation the reference
    stores in the T slot (continuity; at the outlet node the u condition), row 6 the one in the u slot (energy; at
    the inlet node u = u_in, at the outlet node the T condition)."""
    Cin, T_in, T_j, u_in, void, length = cond_row[:5], cond_row[5], cond_row[6], cond_row[7], cond_row[8], cond_row[9]
    dz = length / (NX - 1)
    P0 = np.sum(Cin * R * T_in)
    C, T, u = Y[:5], Y[5], Y[6]
    dC, dT = dY[:5], dY[5]
    F = np.zeros((7, NX))
    # inlet node: held at its start values, velocity prescribed
    F[:5, 0] = dC[:, 0]
    F[5, 0] = dT[0]
    F[6, 0] = u[0] - u_in
    i = np.arange(1, NX - 1)
    r = rate(T[i], C[0, i], C[1, i], C[2, i], C[3, i], k8)
    rho = density(C[:, i], T[i], P0)
    lap_C = C[:, i + 1] - 2 * C[:, i] + C[:, i - 1]
    lap_C[:, 0] = C[:, 2] - C[:, 1]                      # node 1: one-sided dispersion (:99-103)
    F[:5, i] = (-void * dC[:, i] - (u[i] * C[:, i] - u[i - 1] * C[:, i - 1]) / dz + void * DZ_DISP * lap_C / dz ** 2
                + (1 - void) * SC[:, None] * r)
    cont = (-u[i] * P0 * (1 / T[i] - 1 / T[i - 1]) / dz - P0 / T[i] * (u[i] - u[i - 1]) / dz
            + void * DZ_DISP * P0 * (1 / T[i + 1] - 2 / T[i] + 1 / T[i - 1]) / dz ** 2 + (1 - void) * R * (-2) * r)
    cont[0] += P0 * void * T[1] ** (-2) * dT[1]          # node 1 keeps the density storage term (:104)
    F[5, i] = cont
    store = np.full(NX - 2, 0.1)                         # interior nodes: storage term scaled by 0.1 (:121)
    store[0] = 1.0                                       # node 1: not scaled (:105)
    F[6, i] = (-store * (void * rho * CPG + (1 - void) * RHOS * CPS) * dT[i]
               - rho * CPG * (T[i] * u[i] - T[i - 1] * u[i - 1]) / dz + KEFF * (T[i + 1] - 2 * T[i] + T[i - 1]) / dz ** 2
               + (1 - void) * (-HR) * r - 2 * U_WALL / DINT * (T[i] - T_j))
    # outlet node: zero gradient
    n = NX - 1
    F[:5, n] = C[:, n] - C[:, n - 1]
    F[5, n] = u[n] - u[n - 1]
    F[6, n] = T[n] - T[n - 1]
    return F


def start_state(cond_row):
    Y = np.empty((7, NX))
    Y[:5] = cond_row[:5, None]
    Y[5] = T_BED0
    Y[5, 0] = cond_row[5]
    Y[6] = cond_row[7]
    return Y


def _jacobian_banded(Y, Y_old, dt, cond_row, k8, F0):
    """Finite-difference Jacobian of G(Y) = F(Y, (Y - Y_old)/dt) in scipy's band storage, node-major unknown order
    (index 7*node + variable): block tridiagonal, so 13 sub- and super-diagonals; 21 residual evaluations (variable x
    node colour mod 3)."""
    kl = ku = 13
    n = 7 * NX
    ab = np.zeros((kl + ku + 1, n))
    for v in range(7):
        for col in range(3):
            nodes = np.arange(col, NX, 3)
            Yp = Y.copy()
            delta = FD_REL * np.maximum(np.abs(Y[v, nodes]), FLOOR[v])
            Yp[v, nodes] += delta
            delta = Yp[v, nodes] - Y[v, nodes]
            dF = residual(Yp, (Yp - Y_old) / dt, cond_row, k8) - F0
            for jn, d in zip(nodes, delta):
                cidx = 7 * jn + v
                for rn in (jn - 1, jn, jn + 1):
                    if 0 <= rn < NX:
                        rows = 7 * rn + np.arange(7)
                        ab[ku + rows - cidx, cidx] = dF[:, rn] / d
    return ab, kl, ku


def _attempt(Y, Y_old, dt, cond_row, k8, fac):
    """One implicit-Euler step of size dt from Y_old by modified Newton (start iterate Y): the Jacobian of an
    iteration is kept for the following ones, and for later steps of the same size (fac["dt"]), until the update
    stops shrinking by 0.3x.  Returns (Y, converged, iterations)."""
    need_jac, prev_worst = fac.get("dt") != dt, np.inf
    for it in range(NEWTON_MAX):
        F0 = residual(Y, (Y - Y_old) / dt, cond_row, k8)
        if not np.all(np.isfinite(F0)):
            break
        if need_jac:
            need_jac = False
            fac["ab"], fac["dt"] = _jacobian_banded(Y, Y_old, dt, cond_row, k8, F0), dt
        ab, kl, ku = fac["ab"]
        try:
            dx = solve_banded((kl, ku), ab, -F0.T.reshape(-1), check_finite=True)
        except (ValueError, np.linalg.LinAlgError):
            break
        dY = dx.reshape(NX, 7).T
        Y = Y + dY
        if not np.all(np.isfinite(Y)):
            break
        worst = np.max(np.abs(dY) / (np.abs(Y) + FLOOR[:, None]))
        if worst < NEWTON_TOL:
            return Y, True, it + 1
        if worst > 0.3 * prev_worst:
            need_jac = True
        prev_worst = worst
    fac["dt"] = None
    return Y, False, NEWTON_MAX


def integrate(cond_row, k8, return_history=False):
    """Implicit-Euler march of one operating condition to 75 s (twinned by csrc/dae.cu).  A grid step whose Newton
    iteration fails is retried from the same state in pieces: the piece is quartered after a failure and doubled after
    a success; more than MAX_RETRY failures within one grid step fail the march.  Once a whole grid step of the same
    size as the previous one converges at its first iteration the state is steady and the remaining steps are
    skipped.  Returns (Y, ok) or (Y, ok, history of the grid points)."""
    Y = start_state(cond_row)
    hist, fac, steady, H_prev = [], {}, False, 0.0
    with np.errstate(all="ignore"):
        for H in time_grid():
            if not steady:
                t_left, dt, n_fail = H, H, 0
                while t_left > 1e-12 * H and not steady:
                    dt = min(dt, t_left)
                    Y_new, ok, its = _attempt(Y.copy(), Y, dt, cond_row, k8, fac)
                    if ok:
                        steady = its == 1 and dt == H and H == H_prev
                        Y, t_left, dt = Y_new, t_left - dt, dt * 2
                    else:
                        dt, n_fail = dt * 0.25, n_fail + 1
                        if n_fail > MAX_RETRY:
                            return (Y, False, hist) if return_history else (Y, False)
                H_prev = H
            if return_history:
                hist.append(Y.copy())
    return (Y, True, hist) if return_history else (Y, True)


def outlet_flows(full, cond):
    """full: [n, 9] (8 kinetic parameters + sigma); cond: [n_cond, 10].  F[n, 5, n_cond] in sccm (:204-208),
    -10000 where the march failed (:244)."""
    full = np.atleast_2d(np.asarray(full, dtype=np.float64))
    out = np.empty((full.shape[0], 5, cond.shape[0]))
    for p, th in enumerate(full):
        for ci, row in enumerate(cond):
            Y, ok = integrate(row, th[:8])
            P0 = np.sum(row[:5]) * R * row[5]
            T, u = Y[5, -1], Y[6, -1]
            F = Y[:5, -1] * kinetic.S_TUBE * u * 60 * R * T / P0 * 1e6 * P0 / kinetic.P_STP * 298 / T
            out[p, :, ci] = F if ok and np.all(np.isfinite(F)) else kinetic.FAIL_FLOW
    return out


def loglik(theta, cond, obs, base, est_pos):
    """my_loglike (:289-298) over the transient model's outlet flows; theta: [n, d] -> lk[n]."""
    full = kinetic.assemble(theta, base, est_pos)
    sigma = full[:, -1]
    F = outlet_flows(full, cond)
    with np.errstate(all="ignore"):
        ssr = np.sum((F - obs[None, :, :]) ** 2, axis=(1, 2))
        return -(0.5 / sigma ** 2) * ssr - 5.0 * cond.shape[0] * np.log(sigma)
