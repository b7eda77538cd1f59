"""Tempered-SMC sampler loop on the CPU (oracle; test infrastructure only).

Function-by-function restatement of the reference loop
`/root/reference/SMC_example/Micmem_SMC_main.py:105-262` (textually the same
algorithm as `SMC_methanation/SMC_methanation_main.py:201-418`), with every
random input made explicit so the CUDA path can be driven by the same numbers:

    temper_backoff       <- :111-144   (geometric back-off on normalised ESS)
    resample_sequential  <- :147-184   (residual-systematic, sequential FP64 sum)
    proposal_factor      <- numpy legacy `multivariate_normal` (SVD factor) used at :220
    mh_sweep             <- :212-241   (one Metropolis-Hastings sweep)
    run                  <- :105-262   (whole loop) + a log-evidence accumulator

`ReferenceStream` replays the legacy global NumPy stream in the order the
reference driver consumes it (`rand()` :156, `multivariate_normal` :220,
`uniform(0,1,N)` :235) so `run` reproduces the golden run bit for bit.
"""
import math
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Settings:
    """Names and defaults of `Micmem_settings.py:15-31,90` / `methanation_set_conditon.py:107-125`."""
    n_particle: int = 1000
    ess_limit: float = 0.5
    mhstep_factor: float = 0.5
    mhstep_factor_cov: float = 0.5
    ad_mhstep_num: int = 20
    mhstep_num: int = 5
    r_threshold: float = 0.5
    r_threshold_f: float = 0.7
    r_threshold_min: float = 0.1
    d_gamma_max: float = 1
    gm_reduction_itr: int = 80
    gm_reduction_rate: float = 0.7
    itr_max: int = 50

    def w_cov(self, d):
        w = np.full((d, d), self.mhstep_factor_cov, dtype=np.float64)
        np.fill_diagonal(w, self.mhstep_factor)
        return w


# --------------------------------------------------------------------------- tempering
def temper_backoff(lk, gamma_old, cfg: Settings):
    """Returns dict(gamma_new, p_weight, ess, sum_weight, max_lk, n_backoff)."""
    lk = np.asarray(lk, dtype=np.float64)
    N = lk.shape[0]
    gamma_new = gamma_old + cfg.d_gamma_max
    if gamma_new > 1.0:
        gamma_new = 1.0
    max_lk = np.max(lk)
    d_lk = lk - max_lk
    n_backoff = 0
    for i in range(cfg.gm_reduction_itr):
        gm = gamma_new - gamma_old
        p_weight = np.exp(d_lk * gm)
        sum_weight = np.sum(p_weight)
        p_weight = p_weight / sum_weight
        ess = np.sum(p_weight ** 2)
        ess = 1.0 / ess / N
        if ess > cfg.ess_limit:
            break
        gamma_new = (gamma_new - gamma_old) * cfg.gm_reduction_rate + gamma_old
        n_backoff += 1
    # NB: after gm_reduction_itr failures the reference proceeds with the weights
    # of the last *tested* increment while gamma_new has been reduced once more.
    return dict(gamma_new=gamma_new, p_weight=p_weight, ess=ess, sum_weight=sum_weight,
                max_lk=max_lk, n_backoff=n_backoff, gm_used=gm)


def backoff_candidates(gamma_old, cfg: Settings):
    """The fixed candidate list gamma_new_k (k=0..itr-1) the back-off visits."""
    g = gamma_old + cfg.d_gamma_max
    if g > 1.0:
        g = 1.0
    out = []
    for _ in range(cfg.gm_reduction_itr):
        out.append(g)
        g = (g - gamma_old) * cfg.gm_reduction_rate + gamma_old
    return out


def temper_bisect(lk, gamma_old, ess_target, tol=1e-12, max_iter=200):
    """Bisection on Delta-gamma for ESS/N = ess_target (north_star's alternative rule).

    Canonical definition shared with the device host code: lo=0, hi=1-gamma_old;
    if ESS(hi) >= target take hi; else iterate mid=(lo+hi)/2 `n_iter` fixed times.
    """
    lk = np.asarray(lk, dtype=np.float64)
    N = lk.shape[0]
    max_lk = np.max(lk)
    d = lk - max_lk

    def ess_of(gm):
        w = np.exp(d * gm)
        s = np.sum(w)
        return (s * s) / np.sum(w * w) / N, s

    hi = 1.0 - gamma_old
    e, s = ess_of(hi)
    if e >= ess_target:
        return dict(gamma_new=1.0, ess=e, sum_weight=s, max_lk=max_lk, gm_used=hi)
    lo = 0.0
    for _ in range(max_iter):
        mid = 0.5 * (lo + hi)
        e, s = ess_of(mid)
        if e >= ess_target:
            lo = mid
        else:
            hi = mid
        if hi - lo <= tol:
            break
    e, s = ess_of(lo)
    return dict(gamma_new=gamma_old + lo, ess=e, sum_weight=s, max_lk=max_lk, gm_used=lo)


# --------------------------------------------------------------------------- resampling
def resample_sequential(p_weight, u0):
    """Residual-systematic resampling exactly as the reference does it.

    p_weight: normalised weights f64[N];  u0: one U[0,1) draw.
    Returns (ancestors int64[n_filled], counts int64[N], info).  `ancestors` is
    non-decreasing.  `info['n_filled']` may differ from N through rounding; the
    reference does not guard this (over-run raises IndexError, under-fill keeps
    stale rows) - callers decide (the engine clamps / pads with the last ancestor).
    """
    w = np.array(p_weight, dtype=np.float64)
    N = w.shape[0]
    inv_Np = 1 / N
    p_is = np.trunc(w * N).astype(np.int64)
    w = w - p_is * inv_Np
    n_floor = int(np.sum(p_is))
    wrand = u0 * inv_Np
    run = 0.0
    wl = w.tolist()
    cnt = p_is.tolist()
    n_cross = 0
    for j in range(N):
        run += wl[j]
        if run >= wrand:
            cnt[j] += 1
            wrand += inv_Np
            n_cross += 1
    counts = np.array(cnt, dtype=np.int64)
    ancestors = np.repeat(np.arange(N, dtype=np.int64), counts)
    return ancestors, counts, dict(n_floor=n_floor, n_cross=n_cross, n_filled=int(counts.sum()))


def fit_ancestors(ancestors, N):
    """Engine rule for a mis-filled resample: clamp to N, pad with the last ancestor."""
    a = np.asarray(ancestors, dtype=np.int64)
    if a.shape[0] >= N:
        return a[:N].copy()
    pad = np.full(N - a.shape[0], a[-1] if a.shape[0] else 0, dtype=np.int64)
    return np.concatenate([a, pad])


# --------------------------------------------------------------------------- MH mutation
def proposal_factor(cov):
    """Factor F (d x d) with x = z @ F, as numpy's legacy multivariate_normal builds it
    (SVD, not Cholesky): F = sqrt(s)[:,None] * Vt."""
    cov = np.asarray(cov, dtype=np.float64)
    (u, s, v) = np.linalg.svd(cov)
    return np.sqrt(s)[:, None] * v


def proposal_factor_eig(cov):
    """The engine's device-side factor, restated (csrc/mh.cu `moments_merge_factor_kernel`): scale to unit diagonal,
    cov = D^1/2 R D^1/2; symmetric eigen-decomposition R = V diag(lam) V^T, eigenvalues in descending order, every
    eigenvector signed so that its largest component is positive, |lam_j| <= 1e-13 max|lam| treated as 0;
    F[j] = sqrt(|lam_j|) v_j D^1/2, so F^T F = cov: the covariance NumPy's SVD-based sampler draws from.  For given
    normals the proposals differ from `proposal_factor`'s (LAPACK's ordering and signs, no scaling); in distribution
    they are the same."""
    cov = np.atleast_2d(np.asarray(cov, dtype=np.float64))
    cov = 0.5 * (cov + cov.T)
    dg = np.diag(cov)
    sdev = np.where(dg > 0, np.sqrt(np.where(dg > 0, dg, 1.0)), 0.0)
    pos = sdev > 0
    R = np.zeros_like(cov)
    R[np.ix_(pos, pos)] = cov[np.ix_(pos, pos)] / (sdev[pos][:, None] * sdev[pos][None, :])
    R[np.diag_indices_from(R)] = pos.astype(np.float64)
    lam, V = np.linalg.eigh(R)
    order = np.argsort(-lam, kind="stable")
    lam, V = lam[order], V[:, order]
    lmax = np.max(np.abs(lam)) if lam.size else 0.0
    F = np.zeros_like(cov)
    for j in range(lam.size):
        v = V[:, j]
        k = int(np.argmax(np.abs(v)))
        sgn = -1.0 if v[k] < 0 else 1.0
        al = abs(lam[j])
        sc = 0.0 if al <= 1e-13 * lmax else sgn * math.sqrt(al)
        F[j] = (sc * v) * sdev
    return F


def particle_cov(p_filt):
    """Population covariance, `np.cov(p_filt.T, bias=True)` (`Micmem_SMC_main.py:212`)."""
    return np.atleast_2d(np.cov(np.asarray(p_filt).T, bias=True))


def in_box(theta, low, high):
    """`cal_prior(...) > 0` for independent uniform priors on the closed box
    (`Micmem_SMC_main.py:79-82`: scipy.stats.uniform.pdf is >0 on [low, high])."""
    theta = np.asarray(theta)
    return np.all((theta >= low) & (theta <= high), axis=1)


def mh_sweep(p_filt, lk1, gamma_new, F, Z, U, mhstep_ratio, loglik, low, high, log_prior_ratio=None):
    """One sweep.  Z: f64[N,d] standard normals, U: f64[N] uniforms.
    log_prior_ratio(theta_new, theta_old) -> log p(new) - log p(old) for priors with normal components: the
    reference's `pp = np.exp(px * gamma_new) * (p0_2/p0_1)` (`SMC_methanation_main.py:359-375`; never reached
    in its shipped configuration), written with the log ratio so that far-out particles do not give 0/0.
    Returns (p_filt', lk1', r int32[N], n_eval)."""
    step = np.dot(Z, F)
    p_pred = p_filt + step * mhstep_ratio
    p0 = in_box(p_pred, low, high).astype(np.int32)
    p_pred = p_pred * p0[:, None] + p_filt * (1.0 - p0[:, None])
    lk2 = np.asarray(loglik(p_pred), dtype=np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        pp = np.exp((lk2 - lk1) * gamma_new) * p0
        if log_prior_ratio is not None:
            pp = pp * np.exp(np.where(p0 > 0, log_prior_ratio(p_pred, p_filt), 0.0))
    r = (pp >= U).astype(np.int32)
    sel = r.astype(bool)
    p_new = np.where(sel[:, None], p_pred, p_filt)
    lk_new = np.where(sel, lk2, lk1)
    return p_new, lk_new, r, int(p0.sum())


# --------------------------------------------------------------------------- randomness
class ReferenceStream:
    """Replays the legacy global NumPy stream the way the reference driver draws from it."""

    def __init__(self, seed=20250205):
        self.rs = np.random.RandomState(seed)

    def prior_uniform(self, low, high, N):
        """`sample_prior` (`Micmem_settings.py:69-87`): one uniform(low_j, high_j, N) per parameter."""
        return np.stack([self.rs.uniform(l, h, N) for l, h in zip(low, high)], axis=1)

    def u0(self):
        return self.rs.rand()

    def normals(self, N, d):
        return self.rs.standard_normal(N * d).reshape(N, d)

    def uniforms(self, N):
        return self.rs.uniform(0, 1, N)


class PhiloxStream:
    """The engine's default randomness, restated: u0 from a seeded legacy RandomState on the host,
    normals/uniforms from Philox keyed by (seed, global particle id, stage, sweep)."""

    def __init__(self, seed=20250205):
        from . import philox
        self.philox, self.seed = philox, seed
        self.rs = np.random.RandomState(seed)
        self.stage, self.sweep = 0, 0

    def u0(self):
        self.stage += 1
        self.sweep = 0
        return self.rs.rand()

    def normals(self, N, d):
        return self.philox.normals(self.seed, np.arange(N, dtype=np.uint64), self.stage, self.sweep, d)

    def uniforms(self, N):
        u = self.philox.uniforms(self.seed, np.arange(N, dtype=np.uint64), self.stage, self.sweep)
        self.sweep += 1
        return u


class ArrayStream:
    """Hands out pre-generated arrays (shared verbatim with the device path in tests)."""

    def __init__(self, u0s, Zs, Us):
        self._u0, self._Z, self._U = list(u0s), list(Zs), list(Us)

    def u0(self):
        return self._u0.pop(0)

    def normals(self, N, d):
        return self._Z.pop(0)

    def uniforms(self, N):
        return self._U.pop(0)


# --------------------------------------------------------------------------- the loop
@dataclass
class Trace:
    gamma: list = field(default_factory=list)
    ess: list = field(default_factory=list)
    max_lk: list = field(default_factory=list)
    n_backoff: list = field(default_factory=list)
    n_mh: list = field(default_factory=list)          # sweeps actually run
    moved: list = field(default_factory=list)
    log_evidence: list = field(default_factory=list)  # cumulative
    ancestors: list = field(default_factory=list)
    mean: list = field(default_factory=list)
    n_eval: int = 0


def run(loglik, p_pred, low, high, cfg: Settings, stream, lk0=None, hook=None, resampler=None, early_exit=True,
        log_prior_ratio=None, factor=None):
    """Whole tempered-SMC run.  Returns (particles, lk, Trace).

    resampler: `resample_sequential` (the reference, default) or `resample_fixed` (the engine's
    shard-invariant arithmetic); early_exit=False runs all nMH sweeps (BASELINE config 5)."""
    resampler = resampler or resample_sequential
    factor = factor or proposal_factor     # NumPy's SVD factor (the reference); proposal_factor_eig = the device's
    p_pred = np.array(p_pred, dtype=np.float64)
    N, d = p_pred.shape
    tr = Trace()
    lk = np.asarray(loglik(p_pred) if lk0 is None else lk0, dtype=np.float64)
    tr.n_eval += N
    w_cov = cfg.w_cov(d)
    gamma_old = 0.0
    logZ = 0.0
    for step in range(1, cfg.itr_max):
        t = temper_backoff(lk, gamma_old, cfg)
        gamma_new = t["gamma_new"]
        # log-evidence increment (the reference computes sum_weight and discards it, :127)
        logZ += math.log(t["sum_weight"] / N) + t["gm_used"] * t["max_lk"]
        anc, counts, info = resampler(t["p_weight"], stream.u0())
        anc = fit_ancestors(anc, N)
        p_filt = p_pred[anc]
        lk1 = lk[anc]
        r_ac = np.zeros(N)
        mhstep_ratio = 1.0
        if gamma_new >= 1.0:
            nMH, r_th = cfg.ad_mhstep_num, cfg.r_threshold_f
        else:
            nMH, r_th = cfg.mhstep_num, cfg.r_threshold
        n_run = 0
        for j in range(nMH):
            cov_m = particle_cov(p_filt) * w_cov
            F = factor(cov_m)
            Z = stream.normals(N, d)
            U = stream.uniforms(N)
            if hook is not None:
                hook("sweep", step=step, j=j, p_filt=p_filt, lk1=lk1, cov=cov_m, F=F, Z=Z, U=U,
                     gamma=gamma_new, ratio=mhstep_ratio)
            p_filt, lk1, r, ne = mh_sweep(p_filt, lk1, gamma_new, F, Z, U, mhstep_ratio,
                                          loglik, low, high, log_prior_ratio)
            tr.n_eval += N          # the reference evaluates all N (out-of-box ones at the old point)
            r_ac = np.maximum(r_ac, r)
            n_run += 1
            if early_exit and r_ac.sum() > r_th * N:
                break
            if r_ac.sum() < cfg.r_threshold_min * N:
                mhstep_ratio = mhstep_ratio * 0.5
        p_pred = p_filt.copy()
        lk = lk1.copy()
        tr.gamma.append(gamma_new)
        tr.ess.append(t["ess"])
        tr.max_lk.append(t["max_lk"])
        tr.n_backoff.append(t["n_backoff"])
        tr.n_mh.append(n_run)
        tr.moved.append(int(r_ac.sum()))
        tr.log_evidence.append(logZ)
        tr.ancestors.append(anc)
        tr.mean.append(p_pred.mean(axis=0))
        if hook is not None:
            hook("stage", step=step, gamma=gamma_new, p_pred=p_pred, lk=lk)
        if gamma_new == 1.0:
            break
        gamma_old = gamma_new
    return p_pred, lk, tr


# --------------------------------------------------------------------------- fixed-point scan twin
TWO62 = 1 << 62


def resample_fixed(p_weight, u0, N=None):
    """The engine's FIXED scan arithmetic, restated with exact Python/NumPy integers.

    Same floor counts and residuals as the reference (`trunc(w*N)`, `w - p_is*inv_Np`), but the
    residuals are quantised to 2^-62 fixed point (round-to-nearest-even, negatives clamped to 0),
    summed exactly, and particle j receives `cross(s_j) - cross(s_{j-1})` extra copies where
    `cross(s) = floor((s*N - u0q)/2^62) + 1` counts thresholds (u0+k)/N <= s.  No rounding depends on
    the order of summation, so the result is independent of blocking and sharding.
    Returns (ancestors, counts, info).
    """
    w = np.array(p_weight, dtype=np.float64)
    n = w.shape[0]
    N = n if N is None else N
    inv_Np = 1 / N
    fl = np.trunc(w * N)
    resid = w - fl * inv_Np
    rq = np.maximum(resid * float(TWO62), 0.0)
    q = np.rint(rq).astype(np.uint64).astype(object)       # exact Python ints
    u0q = int(np.rint(u0 * float(TWO62)))
    s = 0
    c_prev = 0          # nothing is crossed before the first particle
    counts = np.zeros(n, dtype=np.int64)
    for j in range(n):
        s += int(q[j])
        x = s * N
        c = 0 if x < u0q else ((x - u0q) >> 62) + 1
        counts[j] = int(fl[j]) + (c - c_prev)
        c_prev = c
    ancestors = np.repeat(np.arange(n, dtype=np.int64), counts)
    return ancestors, counts, dict(n_floor=int(fl.sum()), q_total=s, n_filled=int(counts.sum()))


# --------------------------------------------------------------------------- shard twins (multi-GPU host logic tests)
def resample_fixed_shard(w_shard, u0, N, carry_q, first):
    """One shard of `resample_fixed`: counts of the particles in `w_shard` given the exact fixed-point
    residual prefix `carry_q` of all lower ranks (mirrors smcb_resample_counts in FIXED mode with
    carry_q / id_offset).  Returns (counts int64[n], floor_sum, q_total)."""
    w = np.array(w_shard, dtype=np.float64)
    inv_Np = 1 / N
    fl = np.trunc(w * N)
    resid = w - fl * inv_Np
    q = np.rint(np.maximum(resid * float(TWO62), 0.0)).astype(np.uint64).astype(object)
    u0q = int(np.rint(u0 * float(TWO62)))

    def cross(s):
        x = s * N
        return 0 if x < u0q else ((x - u0q) >> 62) + 1

    s = int(carry_q)
    c_prev = 0 if first else cross(s)
    counts = np.zeros(len(w), dtype=np.int64)
    for j in range(len(w)):
        s += int(q[j])
        c = cross(s)
        counts[j] = int(fl[j]) + (c - c_prev)
        c_prev = c
    return counts, int(fl.sum()), s - int(carry_q)


def resample_sequential_shard(w_shard, carry, N):
    """One shard of `resample_sequential`: carry = (running sum, next threshold) entering the shard.
    Returns (counts, carry_out, floor_sum, crossings)."""
    w = np.array(w_shard, dtype=np.float64)
    inv_Np = 1 / N
    p_is = np.trunc(w * N).astype(np.int64)
    w = w - p_is * inv_Np
    run, wrand = float(carry[0]), float(carry[1])
    cnt = p_is.tolist()
    n_cross = 0
    for j, wj in enumerate(w.tolist()):
        run += wj
        if run >= wrand:
            cnt[j] += 1
            wrand += inv_Np
            n_cross += 1
    return np.array(cnt, dtype=np.int64), (run, wrand), int(p_is.sum()), n_cross
