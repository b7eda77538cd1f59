"""Scalar twin of scipy's RK45 (`solve_ivp(method="RK45")`), test infrastructure.

The reference integrates `dS/dt = -Vmax*S/(Km+S)` with
`scipy.integrate.solve_ivp(..., method="RK45", t_eval=t)` at default
tolerances (`/root/reference/SMC_example/Micmem_likelihood.py:24-30`).  The
log-likelihood is therefore *defined* by scipy's adaptive controller, so the
device kernel has to take the same steps.  This module restates that
controller for one scalar ODE, operation for operation, following scipy 1.18.1:

  * tableau C/A/B/E and dense-output matrix P ... `_ivp/rk.py:538-567`
  * first step .................................. `_ivp/common.py:110-134`
  * step loop / error control ................... `_ivp/rk.py:111-176`
  * stage evaluation (dot first, then *h) ....... `_ivp/rk.py:61-71`
  * error estimate h*(K.E) ...................... `_ivp/rk.py:105-109`
  * dense output at t_eval ...................... `_ivp/rk.py:178-180,723-737`,
                                                  `_ivp/ivp.py:712-728`

`tests/test_oracle_dopri5.py` pins it against scipy itself.
"""
import math

RTOL = 1e-3
ATOL = 1e-6
SAFETY = 0.9
MIN_FACTOR = 0.2
MAX_FACTOR = 10.0
ERR_EXP = -1.0 / 5.0       # -1/(error_estimator_order+1)

C = (0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0)
A = (
    (),
    (1 / 5,),
    (3 / 40, 9 / 40),
    (44 / 45, -56 / 15, 32 / 9),
    (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
    (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656),
)
B = (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84)
E = (-71 / 57600, 0.0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40)
P = (
    (1.0, -8048581381 / 2820520608, 8663915743 / 2820520608, -12715105075 / 11282082432),
    (0.0, 0.0, 0.0, 0.0),
    (0.0, 131558114200 / 32700410799, -68118460800 / 10900136933, 87487479700 / 32700410799),
    (0.0, -1754552775 / 470086768, 14199869525 / 1410260304, -10690763975 / 1880347072),
    (0.0, 127303824393 / 49829197408, -318862633887 / 49829197408, 701980252875 / 199316789632),
    (0.0, -282668133 / 205662961, 2019193451 / 616988883, -1453857185 / 822651844),
    (0.0, 40617522 / 29380423, -110615467 / 29380423, 69997945 / 29380423),
)


def initial_step(fun, t0, y0, t_bound, f0):
    """`select_initial_step` for n=1, direction=+1, max_step=inf, order=4."""
    interval = abs(t_bound - t0)
    if interval == 0.0:
        return 0.0
    scale = ATOL + abs(y0) * RTOL
    d0 = abs(y0 / scale)
    d1 = abs(f0 / scale)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = 1e-6
    else:
        h0 = 0.01 * d0 / d1
    h0 = min(h0, interval)
    y1 = y0 + h0 * f0
    f1 = fun(t0 + h0, y1)
    d2 = abs((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / 5.0)
    return min(100 * h0, h1, interval)


def solve_on_grid(fun, y0, t_eval, stats=None):
    """Integrate y' = fun(t, y) from t_eval[0] to t_eval[-1]; return y(t_eval).

    Returns (values, ok).  `ok` is False when the step size underflows (scipy
    status -1); the values list is then shorter than t_eval, exactly like
    `sol.y[0]` would be.
    """
    t = float(t_eval[0])
    t_bound = float(t_eval[-1])
    y = float(y0)
    f = fun(t, y)
    nfev = 1
    h_abs = initial_step(fun, t, y, t_bound, f)
    nfev += 1
    out = []
    i_eval = 0
    n_eval = len(t_eval)
    K = [0.0] * 7
    nstep = nrej = 0
    finished = (t == t_bound)
    if finished:
        # scipy: step() marks finished immediately, t_old = t; dense output is
        # never built, every t_eval <= t is emitted through sol(t) with h=0.
        # Not reachable for the reference data (t spans 0..10).
        while i_eval < n_eval and t_eval[i_eval] <= t:
            out.append(y)
            i_eval += 1
    while not finished:
        min_step = 10 * abs(math.nextafter(t, math.inf) - t)
        if h_abs < min_step:
            h_abs = min_step
        accepted = False
        rejected = False
        while not accepted:
            if h_abs < min_step:
                if stats is not None:
                    stats.update(nfev=nfev, nstep=nstep, nrej=nrej)
                return out, False
            h = h_abs
            t_new = t + h
            if t_new - t_bound > 0:
                t_new = t_bound
            h = t_new - t
            h_abs = abs(h)
            # stages
            K[0] = f
            for s in range(1, 6):
                acc = 0.0
                for j in range(s):
                    acc += K[j] * A[s][j]
                dy = acc * h
                K[s] = fun(t + C[s] * h, y + dy)
            acc = 0.0
            for j in range(6):
                acc += K[j] * B[j]
            y_new = y + h * acc
            f_new = fun(t + h, y_new)
            K[6] = f_new
            nfev += 6
            scale = ATOL + max(abs(y), abs(y_new)) * RTOL
            acc = 0.0
            for j in range(7):
                acc += K[j] * E[j]
            err = abs((acc * h) / scale)
            if err < 1:
                if err == 0:
                    factor = MAX_FACTOR
                else:
                    factor = min(MAX_FACTOR, SAFETY * err ** ERR_EXP)
                if rejected:
                    factor = min(1.0, factor)
                h_abs *= factor
                accepted = True
            else:
                h_abs *= max(MIN_FACTOR, SAFETY * err ** ERR_EXP)
                rejected = True
                nrej += 1
        nstep += 1
        t_old, y_old = t, y
        t, y, f = t_new, y_new, f_new
        finished = (t - t_bound >= 0)
        # dense output for all t_eval in (t_old, t]  (and t_eval[0]==t0 on step 1)
        if i_eval < n_eval and t_eval[i_eval] <= t:
            hh = t - t_old
            Q = [0.0, 0.0, 0.0, 0.0]
            for m in range(4):
                acc = 0.0
                for j in range(7):
                    acc += K[j] * P[j][m]
                Q[m] = acc
            while i_eval < n_eval and t_eval[i_eval] <= t:
                x = (t_eval[i_eval] - t_old) / hh
                p1 = x
                p2 = p1 * x
                p3 = p2 * x
                p4 = p3 * x
                out.append(hh * (Q[0] * p1 + Q[1] * p2 + Q[2] * p3 + Q[3] * p4) + y_old)
                i_eval += 1
    if stats is not None:
        stats.update(nfev=nfev, nstep=nstep, nrej=nrej)
    return out, True
