"""Michaelis-Menten likelihoods on the CPU (oracle; test infrastructure only).

`loglik_progress_*` restates `log_likelihood_mm_multi`
(`/root/reference/SMC_example/Micmem_likelihood.py:35-77`) with
`simulate_mm_on_grid` (`:17-33`) and `mm_ode` (`:14-15`):

    ll(Vmax,Km,sigma) = sum_e [ -0.5*n_t*log(2*pi*sigma^2) - sum_t r_et^2/(2 sigma^2) ],
    r_et = P_obs_et - (S0_e - S_e(t)),   dS/dt = -Vmax*S/(Km+S),  S(0)=S0_e,

integrated with RK45 at rtol=1e-3/atol=1e-6 and sampled by dense output.

Two evaluators of the same arithmetic:
  * `loglik_progress_scipy`  - calls the installed scipy (this IS the reference
    arithmetic; it is what `bench.py --impl reference` times),
  * `loglik_progress_twin`   - the scalar restatement in `oracle.dopri5`
    (what the device kernel mirrors).

`loglik_rate` is the builder-defined rate-law observation model for the
synthetic 10k-observation configuration: observations
(S_i, v_i), v_i ~ N(Vmax*S_i/(Km+S_i), sigma^2).  It has no reference
counterpart; it follows the same Gaussian form as `Micmem_likelihood.py:70-71`.
"""
import math

import numpy as np

from . import dopri5


def _rhs(Vmax, Km):
    return lambda t, S: -Vmax * S / (Km + S)


def loglik_progress_scipy(theta, data_t, data_P, data_S0, return_pred=False):
    """theta = (Vmax, Km, sigma); data_* as loaded by `Micmem_settings.py:103-115`."""
    from scipy.integrate import solve_ivp
    Vmax, Km, sigma = (float(v) for v in theta)
    if sigma <= 0:
        return (-np.inf, None) if return_pred else -np.inf
    n_t = data_t.shape[1]
    total = 0.0
    preds = []
    for e in range(data_t.shape[0]):
        t = data_t[e]
        S0 = float(data_S0[e])
        sol = solve_ivp(fun=lambda tt, S: -Vmax * S / (Km + S), t_span=(t[0], t[-1]),
                        y0=[S0], t_eval=t, method="RK45")
        if sol.y.shape[1] != n_t:            # solver failed: the reference would raise
            return (-np.inf, None) if return_pred else -np.inf
        P_model = S0 - sol.y[0]
        r = data_P[e] - P_model
        total += -0.5 * n_t * np.log(2 * np.pi * sigma ** 2) - np.sum(r ** 2) / (2 * sigma ** 2)
        preds.append(P_model)
    return (total, np.array(preds)) if return_pred else total


def loglik_progress_twin(theta, data_t, data_P, data_S0, stats=None):
    Vmax, Km, sigma = (float(v) for v in theta)
    if sigma <= 0:
        return -math.inf
    n_t = data_t.shape[1]
    total = 0.0
    nfev = 0
    for e in range(data_t.shape[0]):
        st = {}
        S0 = float(data_S0[e])
        ys, ok = dopri5.solve_on_grid(_rhs(Vmax, Km), S0, data_t[e], st)
        nfev += st.get("nfev", 0)
        if not ok:
            return -math.inf
        r = data_P[e] - (S0 - np.asarray(ys))
        total += -0.5 * n_t * math.log(2 * math.pi * sigma ** 2) - np.sum(r ** 2) / (2 * sigma ** 2)
    if stats is not None:
        stats["nfev"] = nfev
    return total


def sweep_progress(particles, data_t, data_P, data_S0, which="scipy"):
    """`sim_particle` (`Micmem_likelihood.py:79-92`) without ray: lk[N]."""
    f = loglik_progress_scipy if which == "scipy" else loglik_progress_twin
    return np.array([f(p, data_t, data_P, data_S0) for p in particles], dtype=np.float64)


def _chunk(args):
    chunk, data_t, data_P, data_S0, which = args
    return sweep_progress(chunk, data_t, data_P, data_S0, which)


def sweep_progress_parallel(particles, data_t, data_P, data_S0, pool, n_chunks, which="scipy"):
    """Fan a sweep out over a process pool, one chunk of particles per task.

    The reference used one `ray` task per particle on `n_cores=30`
    (`Micmem_likelihood.py:83-87`, `Micmem_settings.py:15`); chunking only
    changes scheduling, not arithmetic.
    """
    chunks = np.array_split(np.asarray(particles), n_chunks)
    parts = pool.map(_chunk, [(c, data_t, data_P, data_S0, which) for c in chunks if len(c)])
    return np.concatenate(parts)


def loglik_rate(theta, S, v, dtype=np.float64):
    """Rate-law model, vectorised over particles.  theta: [N,3] -> lk[N].

    ll = -0.5*n*log(2 pi sigma^2) - sum_i (v_i - Vmax*S_i/(Km+S_i))^2 / (2 sigma^2)
    """
    theta = np.asarray(theta, dtype=dtype)
    S = np.asarray(S, dtype=dtype)
    v = np.asarray(v, dtype=dtype)
    Vmax, Km, sigma = theta[:, 0:1], theta[:, 1:2], theta[:, 2]
    n = S.shape[0]
    out = np.empty(theta.shape[0], dtype=np.float64)
    step = max(1, (1 << 22) // max(n, 1))
    for a in range(0, theta.shape[0], step):
        b = min(theta.shape[0], a + step)
        r = v[None, :] - Vmax[a:b] * S[None, :] / (Km[a:b] + S[None, :])
        ssr = np.sum(r.astype(np.float64) ** 2, axis=1)
        sg = sigma[a:b].astype(np.float64)
        out[a:b] = -0.5 * n * np.log(2 * np.pi * sg ** 2) - ssr / (2 * sg ** 2)
    out[theta[:, 2] <= 0] = -np.inf
    return out


def loglik_progress_exact(theta, data_t, data_P, data_S0):
    """The same likelihood with the CONVERGED solution of `mm_ode` (`Micmem_likelihood.py:14-15`) in closed form,
    S(t) = Km * wrightomega(ln(S0/Km) + (S0 - Vmax t)/Km)   - the oracle of the engine's throughput
    integrator SMCB_MM_EXACT.  It is NOT the reference's number: that is defined by scipy's RK45 at rtol 1e-3 and
    differs from this by up to ~3e-3 relative.  theta: [n, 3] -> lk[n]."""
    from scipy.special import wrightomega
    theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    Vmax, Km, sigma = theta[:, 0:1], theta[:, 1:2], theta[:, 2]
    n_t = data_t.shape[1]
    total = np.zeros(theta.shape[0])
    with np.errstate(all="ignore"):
        for e in range(data_t.shape[0]):
            S0 = float(data_S0[e])
            z = np.log(S0 / Km) + (S0 - Vmax * data_t[e][None, :]) / Km
            S = np.where(Km > 0, Km * wrightomega(z).real, np.maximum(S0 - Vmax * data_t[e][None, :], 0.0))
            r = data_P[e][None, :] - (S0 - S)
            total += -0.5 * n_t * np.log(2 * np.pi * sigma ** 2) - np.sum(r * r, axis=1) / (2 * sigma ** 2)
    return np.where(sigma > 0, total, -np.inf)
