"""Per-kernel parity: every C-ABI entry point on the B200 against the CPU oracle / the golden run of the
unmodified reference, on the same inputs.  Bars (BASELINE.json north_star): ancestor indices and
all integer work bit-exact; FP64 quantities within 1e-5 relative (the tests assert far tighter bounds
where the arithmetic allows and say so), FP32 within 1e-3."""
import numpy as np
import pytest
import torch

from oracle import kinetic, mm, philox, smc

pytestmark = pytest.mark.gpu

REL_FP64 = 1e-5   # north_star bar
REL_FP32 = 1e-3


@pytest.fixture(scope="module")
def abi():
    from _abi import Abi
    a = Abi(1 << 21, 32)
    yield a
    a.close()


@pytest.fixture(scope="module")
def mm_abi(abi, golden):
    t, P, S0 = (np.ascontiguousarray(golden[k]) for k in ("data_t", "data_P", "data_S0"))
    abi.ck(abi.lib.smcb_set_data_mm_progress(abi.h, t.ctypes.data, P.ctypes.data, S0.ctypes.data, t.shape[0],
                                             t.shape[1]))
    return abi


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    den = np.maximum(np.abs(b), 1e-300)
    r = np.abs(a - b) / den
    r[same_inf] = 0.0
    return r


# ------------------------------------------------------------------------------------ K1 MM progress
@pytest.mark.parametrize("sweep", [0, 1, 5, 20, 33])
def test_mm_progress_matches_reference_sweeps(mm_abi, golden, sweep):
    """Every particle of sweeps the reference itself evaluated (prior cloud ... posterior cloud)."""
    P, want = golden["sweeps_in"][sweep], golden["sweeps_out"][sweep]
    got = mm_abi.loglik(1, P)
    rel = _rel(got, want)
    assert rel.max() < REL_FP64
    assert rel.max() < 1e-9, rel.max()      # same steps as scipy; single operations differ by an ulp or two
    st = mm_abi.stats()
    assert st[3] == 0 and st[0] > 6 * 1000 * 8
    assert st[0] == 2 * 6000 + 6 * (st[1] + st[2])     # RHS evaluations = 2 per set-up + 6 per attempted step


def test_mm_progress_step_counts_match_c_oracle(mm_abi, golden):
    """The device takes exactly the steps scipy takes: accepted / rejected counts of a whole prior sweep
    equal those of the operation-for-operation C twin of scipy's RK45 (oracle/c/mm_dopri5.c)."""
    from oracle import cmm
    P = golden["sweeps_in"][0]
    want, info = cmm.loglik_progress(P, golden["data_t"], golden["data_P"], golden["data_S0"])
    got = mm_abi.loglik(1, P)
    st = mm_abi.stats()
    assert st[1] == info["accepted"] and st[2] == info["rejected"]
    assert _rel(got, want).max() < 1e-9


def test_mm_progress_prior_cloud_of_65536_particles_matches_c_oracle(mm_abi, golden):
    """SURVEY.md 8(d) C2 subset: a 2^16-particle prior cloud (the first 65536 particles of the bench's Philox prior
    sample, 393 216 solves, stiff ones included) against the C twin of scipy's RK45: every log-likelihood to 1e-9
    relative (bar 1e-5) and exactly the same number of accepted and of rejected steps over the whole cloud."""
    from oracle import cmm
    n = 1 << 16
    P = philox.uniform_box(20250205, np.arange(n, dtype=np.uint64), np.zeros(3), np.full(3, 10.0))
    want, info = cmm.loglik_progress(P, golden["data_t"], golden["data_P"], golden["data_S0"])
    got = mm_abi.loglik(1, P)
    st = mm_abi.stats()
    assert info["failed"] == 0 and st[3] == 0
    assert st[1] == info["accepted"] and st[2] == info["rejected"]
    assert st[11] > 0                                   # some solves went through the tail kernel
    # 1e-9 over the reference's own 34 sweeps; the stiffest solves of a 2^16 prior cloud take 1e4 steps and collect
    # the ulp-level differences of the reciprocal / root corrections: 6e-9 observed, the bar is 1e-5
    assert _rel(got, want).max() < 1e-7, _rel(got, want).max()


@pytest.mark.parametrize("budget", [1, 7, 64, 100000])
def test_mm_progress_is_independent_of_the_deferral_budget(mm_abi, golden, budget):
    """Bulk kernel + tail kernel: wherever a solve is finished it takes the same steps and gives the same result up
    to rounding.  (Round 1 had the same bits: both kernels called mmsolve::attempt.  The tail kernel now runs the
    latency spelling mmsolve::solve_lat, whose roundings fall elsewhere; with a fixed budget - the product never
    changes it during a run - which kernel finishes a solve is a function of the solve alone.)"""
    P = np.concatenate([golden["sweeps_in"][0], golden["sweeps_in"][33]])
    mm_abi.ck(mm_abi.lib.smcb_set_param(mm_abi.h, 1, 256.0))
    base = mm_abi.loglik(1, P)
    st0 = mm_abi.stats()
    mm_abi.ck(mm_abi.lib.smcb_set_param(mm_abi.h, 1, float(budget)))
    got = mm_abi.loglik(1, P)
    st = mm_abi.stats()
    mm_abi.ck(mm_abi.lib.smcb_set_param(mm_abi.h, 1, 512.0))          # the default
    assert _rel(got, base).max() < 1e-11, _rel(got, base).max()
    assert st[1] == st0[1] and st[2] == st0[2]            # the same accepted / rejected steps wherever a solve runs
    if budget < 100000:
        assert st[11] > st0[11] and st[13] > 0          # more solves deferred, the tail kernel had work
    else:
        assert st[11] == 0 and st[13] == 0


def test_mm_progress_bounded_sweep_is_exact_or_certainly_below(mm_abi, golden):
    """smcb_loglik_bounded: a particle reports its exact likelihood, or -inf and then its likelihood is
    certainly below the threshold it was given (prior cloud, thresholds spread around the likelihoods)."""
    P = golden["sweeps_in"][0]
    n = len(P)
    rs = np.random.RandomState(4)
    thr = mm_abi.loglik(1, P) + rs.normal(0, 300, n)
    thr[:50] = -np.inf
    thr[50:60] = np.inf
    th = mm_abi.t(np.asarray(P).T)
    lk, tt = mm_abi.zeros(n), mm_abi.t(thr)
    for budget in (256.0, 3.0):
        mm_abi.ck(mm_abi.lib.smcb_set_param(mm_abi.h, 1, budget))
        full = mm_abi.loglik(1, P)              # the unbounded evaluation at the same budget: same kernels, same bits
        mm_abi.ck(mm_abi.lib.smcb_loglik_bounded(mm_abi.h, 1, th.data_ptr(), n, n, 3, None, tt.data_ptr(), lk.data_ptr(), None))
        got = lk.cpu().numpy()
        cut = np.isneginf(got) & ~np.isneginf(full)
        assert np.array_equal(got[~cut], full[~cut])
        assert np.all(full[cut] < thr[cut])
        assert not cut[:50].any() and cut[50:60].all()
        assert cut.sum() >= 10
        assert mm_abi.stats()[8] == np.isneginf(got).sum()
    mm_abi.ck(mm_abi.lib.smcb_set_param(mm_abi.h, 1, 512.0))          # the default


def test_mh_threshold_is_conservative(abi):
    """lkmin = lk1 + log(u)/gamma - margin: any lk2 below it fails exp((lk2-lk1)*gamma) >= u."""
    n = 4096
    rs = np.random.RandomState(8)
    lk1 = rs.normal(100, 300, n)
    u = rs.uniform(0, 1, n)
    u[:4] = [0.0, 1e-300, 0.999999999, 0.5]
    lk1[4:7] = [-np.inf, np.inf, np.nan]
    inbox = (rs.uniform(0, 1, n) < 0.8).astype(np.uint8)
    inbox[:7] = 1
    for gamma in (0.0023263051398720674, 0.37, 1.0):
        out = abi.zeros(n)
        lt, ut, it = abi.t(lk1), abi.t(u), abi.t(inbox, torch.uint8)
        abi.ck(abi.lib.smcb_mh_threshold(abi.h, lt.data_ptr(), it.data_ptr(), n, gamma, ut.data_ptr(), None, 0, 0, 0, 0,
                                         out.data_ptr(), None))
        thr = out.cpu().numpy()
        assert np.all(np.isneginf(thr[inbox == 0])) and np.isneginf(thr[0]) and np.all(np.isneginf(thr[4:7]))
        ok = np.isfinite(thr)
        assert ok.sum() > 0.7 * n
        with np.errstate(over="ignore"):
            lk2 = np.nextafter(thr[ok], -np.inf)
            assert not np.any(np.exp((lk2 - lk1[ok]) * gamma) >= u[ok])         # certain rejection just below
            exact = lk1[ok] + np.log(u[ok]) / gamma
        assert np.all(thr[ok] < exact) and np.all(exact - thr[ok] < 1e-8 * (1 + np.abs(lk1[ok]) + np.abs(exact)))


def test_mm_progress_known_answers_and_edge_cases(mm_abi, golden):
    th = np.array([[1.2, 0.5, 0.02], [1.0, 0.4, 0.05], [1.0, 0.4, 0.0], [1.0, 0.4, -1.0], [0.0, 1.0, 1.0],
                   [5.0, 1e-9, 0.1], [1e-12, 5.0, 0.3], [10.0, 10.0, 10.0]])
    got = mm_abi.loglik(1, th)
    d = (golden["data_t"], golden["data_P"], golden["data_S0"])
    want = np.array([mm.loglik_progress_scipy(p, *d) for p in th])
    assert abs(got[0] - 593.9635684697922) < 1e-7 and abs(got[1] - 424.5676547271564) < 1e-7
    assert got[2] == -np.inf and got[3] == -np.inf            # sigma <= 0 (Micmem_likelihood.py:53-54)
    assert _rel(got, want).max() < 1e-8


def test_mm_progress_active_mask_and_ragged_sizes(mm_abi, golden):
    P = golden["sweeps_in"][0]
    full = mm_abi.loglik(1, P)
    for n in (1, 31, 33, 127, 129, 1000):
        assert np.array_equal(mm_abi.loglik(1, P[:n]), full[:n])      # independent of launch shape
    act = (np.arange(1000) % 3 == 0).astype(np.uint8)
    part = mm_abi.loglik(1, P, active=act)
    assert np.array_equal(part[act == 1], full[act == 1]) and np.all(part[act == 0] == 0.0)


def test_mm_progress_predictions(mm_abi, golden):
    P = golden["prior_particles"][:8]
    th = mm_abi.t(P.T)
    pred = mm_abi.zeros(8, 6, 40)
    mm_abi.ck(mm_abi.lib.smcb_predict_mm_progress(mm_abi.h, th.data_ptr(), 8, 8, pred.data_ptr(), None))
    got = pred.cpu().numpy()
    want = golden["pmodel0"]
    assert np.abs(got - want).max() < 1e-11


def test_mm_progress_exact_integrator_matches_its_oracle(mm_abi, golden):
    """SMCB_MM_EXACT (closed-form progress curves, converged mode) against oracle.mm.loglik_progress_exact over a
    prior cloud (stiff particles included: Vmax/Km up to 1e5), a posterior cloud and edge cases; stated bound 1e-9."""
    a = mm_abi
    d = (golden["data_t"], golden["data_P"], golden["data_S0"])
    rs = np.random.RandomState(12)
    prior = rs.uniform(0, 10, (20000, 3))
    prior[:50, 1] = 10.0 ** rs.uniform(-6, -2, 50)          # very stiff: S0/Km up to 2e6
    prior[50, 2] = 0.0
    post = golden["final_particles"]
    try:
        a.ck(a.lib.smcb_set_param(a.h, 7, 1.0))
        for th, tol in ((prior, 1e-9), (post, 1e-10)):
            want = mm.loglik_progress_exact(th, *d)
            got = a.loglik(1, th)
            assert _rel(got, want).max() < tol, _rel(got, want).max()
        assert a.loglik(1, prior)[50] == -np.inf
        # early rejection: exact value, or -inf and then the value is certainly below the threshold given
        full = a.loglik(1, prior)
        thr = full + rs.normal(0, 300, len(full))
        thr[:100] = -np.inf
        thp, lkb, tt = a.t(prior.T), a.zeros(len(full)), a.t(thr)
        a.ck(a.lib.smcb_loglik_bounded(a.h, 1, thp.data_ptr(), len(full), len(full), 3, None, tt.data_ptr(), lkb.data_ptr(), None))
        got_b = lkb.cpu().numpy()
        cut = np.isneginf(got_b) & ~np.isneginf(full)
        assert np.array_equal(got_b[~cut], full[~cut]) and np.all(full[cut] < thr[cut]) and not cut[:100].any()
        assert cut.sum() > 1000
        # predictions (the reference's C_l_) follow the same curves
        pred = a.zeros(4, 6, 40)
        thp = a.t(post[:4].T)
        a.ck(a.lib.smcb_predict_mm_progress(a.h, thp.data_ptr(), 4, 4, pred.data_ptr(), None))
        S0 = d[2][None, :, None]
        from scipy.special import wrightomega
        z = np.log(S0 / post[:4, 1][:, None, None]) + (S0 - post[:4, 0][:, None, None] * d[0][None]) / post[:4, 1][:, None, None]
        want_pred = S0 - post[:4, 1][:, None, None] * wrightomega(z).real
        assert np.abs(pred.cpu().numpy() - want_pred).max() < 1e-12
        # and the distance to the reference's own (rtol 1e-3) likelihood is what SURVEY.md H1 says it is
        ref = golden["sweeps_out"][-1]
        ex = a.loglik(1, golden["sweeps_in"][-1])
        fin = np.isfinite(ref)
        assert 1e-8 < _rel(ex[fin], ref[fin]).max() < 1e-2
    finally:
        a.ck(a.lib.smcb_set_param(a.h, 7, 0.0))


# ------------------------------------------------------------------------------------ K1 MM rate
@pytest.mark.parametrize("n_obs", [1, 63, 2048, 10000])
def test_mm_rate_fp64_and_fp32(abi, n_obs):
    rs = np.random.RandomState(3)
    S = np.exp(rs.uniform(np.log(0.05), np.log(20.0), n_obs))
    v = 1.2 * S / (0.5 + S) + 0.02 * rs.standard_normal(n_obs)
    th = rs.uniform(0, 10, (777, 3))
    th[5, 2] = 0.0
    want = mm.loglik_rate(th, S, v)
    for prec, tol in ((64, 1e-12), (32, REL_FP32)):
        abi.ck(abi.lib.smcb_set_data_mm_rate(abi.h, S.ctypes.data, v.ctypes.data, n_obs, prec))
        got = abi.loglik(2, th)
        assert got[5] == -np.inf
        assert _rel(got, want).max() < tol, (prec, _rel(got, want).max())


@pytest.mark.parametrize("n_obs", [1, 63, 10000])
def test_mm_rate_sufficient_statistic_form_matches_direct_sum(abi, n_obs):
    """sum v^2 - 2 Vmax A(Km) + Vmax^2 B(Km) with tabulated A, B against the direct sum of the
    oracle, over the prior box, around the data-generating point (where the three terms cancel to 6e-4 of their
    size) and outside the tabulated Km range (direct-sum branch).  Stated bound: 1e-9 relative on the log-likelihood."""
    rs = np.random.RandomState(3)
    S = np.exp(rs.uniform(np.log(0.05), np.log(20.0), n_obs))
    v = 1.2 * S / (0.5 + S) + 0.02 * rs.standard_normal(n_obs)
    abi.ck(abi.lib.smcb_set_data_mm_rate_sufficient(abi.h, S.ctypes.data, v.ctypes.data, n_obs, 0.0, 10.0))
    prior = rs.uniform(0, 10, (5000, 3))
    prior[5, 2] = 0.0                                  # sigma = 0
    prior[6, 1] = 0.0                                  # Km at the lower edge of the table
    prior[7, 1] = 10.0                                 # ... and at the upper edge
    post = np.c_[1.2 + 1e-3 * rs.standard_normal(3000), 0.5 + 2e-3 * rs.standard_normal(3000),
                 0.02 + 2e-4 * rs.standard_normal(3000)]
    outside = np.c_[rs.uniform(0, 10, 200), rs.uniform(10.0001, 50, 200), rs.uniform(0.01, 10, 200)]
    for th, tol in ((prior, 1e-11), (post, 1e-9), (outside, 1e-12)):
        want = mm.loglik_rate(th, S, v)
        got = abi.loglik(2, th)
        assert _rel(got, want).max() < tol, _rel(got, want).max()
    assert abi.loglik(2, prior)[5] == -np.inf
    # masked particles keep their old value
    act = (np.arange(len(post)) % 3 == 0).astype(np.uint8)
    got = abi.loglik(2, post, active=act)
    assert np.all(got[act == 0] == 0.0) and _rel(got[act == 1], mm.loglik_rate(post, S, v)[act == 1]).max() < 1e-9


# ------------------------------------------------------------------------------------ K1' kinetic
@pytest.mark.parametrize("n_pairs,d", [(4, 5), (16, 32)])
def test_kinetic_matches_oracle(abi, n_pairs, d):
    cond = kinetic.synthetic_conditions(30)
    base = kinetic.base_vector(n_pairs)
    obs = kinetic.synthetic_observations(cond, base)
    if n_pairs == 4:
        est = np.array(kinetic.EST_POSITION, dtype=np.int32)
        low, high = kinetic.reference_box()
    else:
        est = np.arange(32, dtype=np.int32)
        low, high = base[:32] * 0.9, base[:32] * 1.1
        low, high = np.minimum(low, high), np.maximum(low, high)
    abi.ck(abi.lib.smcb_set_data_kinetic(abi.h, cond.ctypes.data, obs.ctypes.data, 30, base.ctypes.data, n_pairs,
                                         est.ctypes.data, d, 50))
    rs = np.random.RandomState(1)
    th = rs.uniform(low, high, (300, d))
    th[0] = base[est]
    want = kinetic.loglik(th, cond, obs, base, est, 50)
    got = abi.loglik(3, th)
    assert np.all(np.isfinite(got))
    # A few percent of the wide reference prior box are fast-kinetics particles for which the explicit
    # fixed-step march is numerically unstable: there the result is sensitive to single ulps (the oracle
    # disagrees with *itself* when a parameter moves by one ulp).  Those particles carry likelihoods
    # ~1e6 below the mode (zero weight); parity is asserted on the well-conditioned ones.
    th_ulp = th.copy()
    th_ulp[:, 0] = np.nextafter(th_ulp[:, 0], np.inf)
    sens = _rel(kinetic.loglik(th_ulp, cond, obs, base, est, 50), want)
    good = sens < 1e-12
    # how many are excluded is part of the statement: 10 of 300 in the reference's 5-parameter box, none of the 300
    # drawn from the +-10% box of the 32-parameter family (observed; the bound below leaves room for libm changes)
    assert good[0] and (~good).sum() <= (24 if n_pairs == 4 else 6), (~good).sum()
    assert _rel(got, want)[good].max() < 1e-9, _rel(got, want)[good].max()
    if (~good).any():
        # ... there any ulp-level difference in the arithmetic (NumPy vs CUDA exp(), reciprocal-multiply vs divide)
        # gives a different number; what must hold is that both sides agree the particle is hopeless
        assert np.all(want[~good] < want[0] - 100) and np.all(got[~good] < want[0] - 100)
        ratio = got[~good] / want[~good]
        assert np.all((ratio > 0.05) & (ratio < 20)), (ratio.min(), ratio.max())


def test_kinetic_fixture_known_answers(abi):
    """tests/golden/kinetic_synth.npz: oracle likelihoods committed with their inputs."""
    import os
    kf = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kinetic_synth.npz"))
    for tag, n_pairs, est in (("4", 4, kf["est4"]), ("16", 16, np.arange(32, dtype=np.int32))):
        cond, base, obs = (np.ascontiguousarray(kf[k]) for k in ("cond", "base" + tag, "obs" + tag))
        est = np.ascontiguousarray(est, dtype=np.int32)
        abi.ck(abi.lib.smcb_set_data_kinetic(abi.h, cond.ctypes.data, obs.ctypes.data, 30, base.ctypes.data, n_pairs,
                                             est.ctypes.data, len(est), 50))
        got, want = abi.loglik(3, kf["theta" + tag]), kf["lk" + tag]
        rel = _rel(got, want)
        assert np.median(rel) < 1e-12 and (rel < 1e-9).mean() > 0.9, (np.median(rel), rel.max())


@pytest.mark.parametrize("n_pairs", [4, 16])
def test_kinetic_masked_sweep_is_packed_not_changed(abi, n_pairs):
    """A masked sweep packs the active particles into full warps first; values equal the unmasked sweep's
    bit for bit, inactive slots are left alone, ragged sizes and empty / sparse / full masks included."""
    import os
    kf = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kinetic_synth.npz"))
    tag = str(n_pairs)
    est = np.ascontiguousarray(kf["est4"] if n_pairs == 4 else np.arange(32), dtype=np.int32)
    cond, base, obs = (np.ascontiguousarray(kf[k]) for k in ("cond", "base" + tag, "obs" + tag))
    abi.ck(abi.lib.smcb_set_data_kinetic(abi.h, cond.ctypes.data, obs.ctypes.data, 30, base.ctypes.data, n_pairs,
                                         est.ctypes.data, len(est), 50))
    th = np.tile(kf["theta" + tag], (16, 1))[:1001]
    n = len(th)
    full = abi.loglik(3, th)
    rs = np.random.RandomState(5)
    for n_sub, p_on in ((n, 0.02), (n, 0.5), (n, 1.0), (n, 0.0), (min(n, 131), 0.3), (1, 1.0)):
        act = (rs.uniform(size=n_sub) < p_on).astype(np.uint8)
        part = abi.loglik(3, th[:n_sub], active=act)
        assert np.array_equal(part[act == 1], full[:n_sub][act == 1])
        assert np.all(part[act == 0] == 0.0)


def test_transient_reactor_matches_oracle_march(abi):
    """SURVEY.md 8(f) N3, model KINETIC_DAE: one thread block per (particle, condition) marches the reference's
    357-unknown reactor DAE to 75 s; same implicit-Euler grid and Newton tolerance as oracle/methanation_dae.py."""
    from oracle import methanation_dae as dae
    cond = kinetic.synthetic_conditions(3)
    base = kinetic.base_vector(4)
    est = np.array(kinetic.EST_POSITION, dtype=np.int32)
    obs = kinetic.synthetic_observations(cond, base)
    low, high = kinetic.reference_box()
    abi.ck(abi.lib.smcb_set_data_kinetic(abi.h, cond.ctypes.data, obs.ctypes.data, 3, base.ctypes.data, 4,
                                         est.ctypes.data, 5, 1))
    rs = np.random.RandomState(2)
    th = np.vstack([base[est], base[est] * rs.uniform(0.8, 1.25, (5, 5))])
    want = dae.loglik(th, cond, obs, base, est)
    got = abi.loglik(4, th)
    assert np.all(np.isfinite(got)) and np.all(want > -1e7)       # none of these marches fails
    assert _rel(got, want).max() < 1e-7, _rel(got, want).max()    # Newton tolerance 1e-10 on both sides; bar 1e-5
    # across the reference's wide prior box a third of the particles have kinetics the fixed-grid march cannot
    # follow in some condition (the reference penalises IDA failures the same way): both sides must call them
    # hopeless, and agree to the tolerance on every particle that marches through
    wide = rs.uniform(low, high, (6, 5))
    want_w, got_w = dae.loglik(wide, cond, obs, base, est), abi.loglik(4, wide)
    hopeless = want_w < -1e6
    assert np.array_equal(hopeless, got_w < -1e6)
    if (~hopeless).any():
        assert _rel(got_w[~hopeless], want_w[~hopeless]).max() < 1e-6
    act = np.array([1, 0, 1, 1, 0, 1], dtype=np.uint8)            # masked sweep: packed work list
    part = abi.loglik(4, th, active=act)
    assert np.array_equal(part[act == 1], got[act == 1]) and np.all(part[act == 0] == 0.0)
    # a particle whose kinetics make Newton diverge gets the reference's failure penalty, not an error
    bad = th[:1].copy()
    bad[0, 0] *= 1e12
    lk_bad = abi.loglik(4, bad)
    assert np.isfinite(lk_bad[0]) and lk_bad[0] < got[0] - 1e3


def test_transient_reactor_matches_golden_flows_of_256_particles(abi):
    """SURVEY.md 8(f) N3 at a size the oracle cannot reach inside a test: 256 particles x the 30 operating conditions of
    the bench workload (7680 marches) against outlet flows the CPU oracle produced offline
    (tests/golden/make_dae_flows_fixture.py -> dae_flows_256.npz): 192 particles around the data-generating
    parameters, 64 from the reference's wide prior box, where some marches fail on both sides."""
    import os
    gd = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g, fx = np.load(os.path.join(gd, "dae_synth.npz")), np.load(os.path.join(gd, "dae_flows_256.npz"))
    cond, base, obs = (np.ascontiguousarray(g[k]) for k in ("cond", "base4", "obs"))
    est = np.ascontiguousarray(g["est4"], dtype=np.int32)
    abi.ck(abi.lib.smcb_set_data_kinetic(abi.h, cond.ctypes.data, obs.ctypes.data, cond.shape[0], base.ctypes.data, 4,
                                         est.ctypes.data, len(est), 1))
    th, want, flows = fx["theta"], fx["lk"], fx["flows"]
    got = abi.loglik(4, th)
    n_fail = (flows <= -9999).any(axis=1).sum(axis=1)   # conditions whose march failed in the oracle (-10000 penalty)
    rel = _rel(got, want)
    # around the data-generating parameters no march fails and device = oracle to rounding (2e-15 observed)
    assert n_fail[:192].sum() == 0 and rel[:192].max() < 1e-9, rel[:192].max()
    # across the wide prior box some marches fail (Newton does not converge on the grid even after the retries); which
    # ones is decided at the edge of convergence, where the last bits of the two implementations matter: both sides
    # must agree wherever no march failed in the oracle and the device did not penalise either, and on all but a few
    # of the others (4 of 64 differ in the set of failed conditions on this fixture)
    wide = np.arange(192, 256)
    clean = wide[(n_fail[wide] == 0) & (got[wide] > -1e5)]
    assert len(clean) >= 50 and rel[clean].max() < 1e-6, rel[clean].max()
    agree = rel[wide] < 1e-6
    assert agree.sum() >= 58, (agree.sum(), wide[~agree], got[wide][~agree], want[wide][~agree])
    assert 0 < (n_fail[wide] > 0).sum() < 16


# ------------------------------------------------------------------------------------ K2 tempering
@pytest.mark.parametrize("n", [1, 2, 777, 1000, (1 << 20) + 3])
def test_temper_reductions(abi, n):
    rs = np.random.RandomState(n % 1000)
    lk = rs.normal(300, 80, n)
    if n > 10:
        lk[3] = -np.inf
        lk[7] = -618002.599096893
    t = abi.t(lk)
    out = abi.zeros(40)
    abi.ck(abi.lib.smcb_lk_max(abi.h, t.data_ptr(), n, out.data_ptr(), None))
    assert out[0].item() == lk.max()
    gms = np.array([1.0 * 0.7 ** k for k in range(0, 32, 2)])
    for nc in (1, 2, 3, 8, 16):
        g = np.ascontiguousarray(gms[:nc])
        abi.ck(abi.lib.smcb_temper_sums(abi.h, t.data_ptr(), n, out.data_ptr(), g.ctypes.data, nc,
                                        out[2:].data_ptr(), None))
        got = out[2:2 + 2 * nc].cpu().numpy()
        for k in range(nc):
            w = np.exp((lk - lk.max()) * g[k])
            assert abs(got[2 * k] - w.sum()) <= 1e-12 * w.sum()
            assert abs(got[2 * k + 1] - (w * w).sum()) <= 1e-12 * (w * w).sum()


def test_weights_kernel_matches_numpy(abi):
    rs = np.random.RandomState(0)
    lk = rs.normal(300, 80, 5000)
    t, out, w = abi.t(lk), abi.zeros(4), abi.zeros(5000)
    gm = 0.0023263051398720674
    ref = np.exp((lk - lk.max()) * gm)
    out[0], out[1] = lk.max(), ref.sum()
    abi.ck(abi.lib.smcb_weights(abi.h, t.data_ptr(), 5000, out.data_ptr(), gm, out[1:].data_ptr(), w.data_ptr(), None))
    assert _rel(w.cpu().numpy(), ref / ref.sum()).max() < 4e-16      # exp() within an ulp


# ------------------------------------------------------------------------------------ K3 resampling
def _weights(n, seed, conc=0.3):
    return np.random.RandomState(seed).dirichlet(np.full(n, conc))


def _resample(abi, w, u0, mode, n_total=None, carry_q=0, id_offset=0, m=None):
    n = w.shape[0]
    n_total = n if n_total is None else n_total
    m = n if m is None else m
    wt = abi.t(w)
    counts = abi.zeros(n, dtype=torch.int32)
    tot = abi.zeros(2, dtype=torch.int64)
    abi.ck(abi.lib.smcb_resample_counts(abi.h, wt.data_ptr(), n, n_total, u0, mode, None, carry_q, id_offset,
                                        counts.data_ptr(), tot.data_ptr(), None))
    anc = abi.zeros(m, dtype=torch.int32)
    filled = abi.zeros(1, dtype=torch.int64)
    abi.ck(abi.lib.smcb_ancestors(abi.h, counts.data_ptr(), n, m, anc.data_ptr(), filled.data_ptr(), None))
    torch.cuda.synchronize()
    return counts.cpu().numpy().astype(np.int64), anc.cpu().numpy().astype(np.int64), tot.cpu().numpy(), int(filled.item())


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (7, 2), (1000, 3), (2048, 4), (2049, 5), (5000, 6), (65536, 7)])
@pytest.mark.parametrize("u0", [0.0, 0.37, 0.999999])
def test_resample_sequential_is_bit_exact(abi, n, seed, u0):
    """Same weights, same uniform -> the reference's ancestors, bit for bit (Micmem_SMC_main.py:147-184)."""
    w = _weights(n, seed)
    a_ref, c_ref, info = smc.resample_sequential(w, u0)
    c, a, tot, filled = _resample(abi, w, u0, 0)
    assert np.array_equal(c, c_ref)
    assert filled == info["n_filled"] and tot[0] == info["n_floor"] and tot[1] == info["n_cross"]
    assert np.array_equal(a, smc.fit_ancestors(a_ref, n))


@pytest.mark.parametrize("n,conc,u0", [((1 << 20) + 3, 0.3, 0.37), (1 << 20, 0.02, 0.0), (300007, 5.0, 0.999999),
                                       (1 << 18, 1.0, 0.5)])
def test_resample_sequential_is_bit_exact_at_full_size(abi, n, conc, u0):
    """The block-parallel form of the reference's sequentially rounded running sum (integer scans inside a binade,
    literal loop where a chunk breaks an assumption) against the literal Python loop at BASELINE config 2's size,
    with even, concentrated and very concentrated weights: counts and ancestors bit for bit."""
    w = _weights(n, 21, conc=conc)
    a_ref, c_ref, info = smc.resample_sequential(w, u0)
    c, a, tot, filled = _resample(abi, w, u0, 0)
    assert np.array_equal(c, c_ref)
    assert filled == info["n_filled"] and tot[0] == info["n_floor"] and tot[1] == info["n_cross"]
    assert np.array_equal(a, smc.fit_ancestors(a_ref, n))


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (7, 2), (1000, 3), (2048, 4), (2049, 5), (5000, 6), (65536, 7)])
@pytest.mark.parametrize("u0", [0.0, 0.37, 0.999999])
def test_resample_fixed_matches_integer_twin(abi, n, seed, u0):
    w = _weights(n, seed)
    a_ref, c_ref, info = smc.resample_fixed(w, u0)
    c, a, tot, filled = _resample(abi, w, u0, 1)
    assert np.array_equal(c, c_ref)
    assert tot[0] == info["n_floor"] and tot[1] == info["q_total"]
    assert np.array_equal(a, smc.fit_ancestors(a_ref, n))


def _resample_fused(abi, w, u0, D1=4, lk=None, mx=0.0, gm=0.0, sum_w=1.0):
    """smcb_resample_fused on explicit weights (or on lk / max / gm / sum_w); returns counts, ancestors, filled, dst."""
    n = (w if w is not None else lk).shape[0]
    src = torch.arange(D1 * n, dtype=torch.float64, device=abi.dev).reshape(D1, n) * 0.5 + 1.0
    dst = torch.full_like(src, -1.0)
    counts, anc = abi.zeros(n, dtype=torch.int32), abi.zeros(n, dtype=torch.int32)
    filled = abi.zeros(1, dtype=torch.int64)
    wt = abi.t(w) if w is not None else None
    lkt = abi.t(lk) if lk is not None else None
    sc = abi.t(np.array([mx, sum_w]))
    abi.ck(abi.lib.smcb_resample_fused(abi.h, lkt.data_ptr() if lkt is not None else None,
                                       wt.data_ptr() if wt is not None else None, n, n, 0, 0, n, sc.data_ptr(), gm,
                                       sc[1:].data_ptr(), u0, src.data_ptr(), n, D1, dst.data_ptr(), n, anc.data_ptr(),
                                       counts.data_ptr(), filled.data_ptr(), None))
    torch.cuda.synchronize()
    return (counts.cpu().numpy().astype(np.int64), anc.cpu().numpy().astype(np.int64), int(filled.item()),
            src.cpu().numpy(), dst.cpu().numpy())


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (7, 2), (1000, 3), (2048, 4), (2049, 5), (5000, 6), (65536, 7),
                                    ((1 << 20) + 77, 8)])
@pytest.mark.parametrize("u0", [0.0, 0.37, 0.999999])
def test_fused_resampling_equals_the_kernel_chain(abi, n, seed, u0):
    """One kernel (look-back scan + cooperative expansion + gather) against counts -> ancestors -> gather and the
    exact-integer twin of the oracle: counts, ancestors, filled count and the moved rows, bit for bit."""
    w = _weights(n, seed)
    c_ref, a_ref, _, f_ref = _resample(abi, w, u0, 1)
    c, a, filled, src, dst = _resample_fused(abi, w, u0)
    assert filled == f_ref and np.array_equal(c, c_ref) and np.array_equal(a, a_ref)
    assert np.array_equal(dst, src[:, a])
    if n <= 65536:
        a_o, c_o, info = smc.resample_fixed(w, u0)
        assert np.array_equal(c, c_o) and np.array_equal(a, smc.fit_ancestors(a_o, n)) and filled == info["n_filled"]


def test_fused_resampling_skewed_weights_and_weights_on_the_fly(abi):
    # one particle takes (almost) everything: a single tile expands the whole output
    n = 300001
    w = np.full(n, 1e-9 / n)
    w[123457] = 1.0 - w.sum() + w[123457]
    c_ref, a_ref, _, f_ref = _resample(abi, w, 0.25, 1)
    c, a, filled, src, dst = _resample_fused(abi, w, 0.25)
    assert filled == f_ref and np.array_equal(c, c_ref) and np.array_equal(a, a_ref) and np.array_equal(dst, src[:, a])
    assert c[123457] >= n - 1
    # weights computed inside the kernel = smcb_weights followed by the chain
    rs = np.random.RandomState(1)
    lk = rs.normal(-500, 40, 70001)
    gm = 0.013
    mx = lk.max()
    sum_w = float(np.exp((lk - mx) * gm).sum())
    lkt, sc, wd = abi.t(lk), abi.t(np.array([mx, sum_w])), abi.zeros(len(lk))
    abi.ck(abi.lib.smcb_weights(abi.h, lkt.data_ptr(), len(lk), sc.data_ptr(), gm, sc[1:].data_ptr(), wd.data_ptr(), None))
    c_ref, a_ref, _, f_ref = _resample(abi, wd.cpu().numpy(), 0.7, 1)
    c, a, filled, src, dst = _resample_fused(abi, None, 0.7, lk=lk, mx=mx, gm=gm, sum_w=sum_w)
    assert filled == f_ref and np.array_equal(c, c_ref) and np.array_equal(a, a_ref) and np.array_equal(dst, src[:, a])


def test_fused_resampling_shard_by_shard_equals_unsharded(abi):
    """The sharded form of the single-pass kernel (residual prefix of the lower ranks, global index of the shard's first
    particle, the slots the shard fills): the shards' outputs concatenated are the unsharded result, bit for bit."""
    import smcb200
    n, W, D1 = 1 << 16, 4, 4
    nl = n // W
    for seed, u0 in ((11, 0.6180339887), (12, 0.0)):
        w = _weights(n, seed, conc=0.05)
        c_full, a_full, tot_full, f_full = _resample(abi, w, u0, 1)
        src = torch.arange(D1 * n, dtype=torch.float64, device=abi.dev).reshape(D1, n) * 0.25 - 3.0
        tots = []
        for r in range(W):
            wt, tot = abi.t(w[r * nl:(r + 1) * nl]), abi.zeros(2, dtype=torch.int64)
            abi.ck(abi.lib.smcb_resample_totals(abi.h, wt.data_ptr(), nl, n, tot.data_ptr(), None))
            tots.append(tot.cpu().numpy())
        tots = np.array(tots)
        plan = smcb200.migration_plan(tots[:, 0], tots[:, 1], n, nl, u0, W)
        assert plan["filled"] == f_full
        out_rows, out_anc = [], []
        for r in range(W):
            m_loc = plan["M"][r]
            wt = abi.t(w[r * nl:(r + 1) * nl])
            shard = src[:, r * nl:(r + 1) * nl].contiguous()
            dst = torch.full((D1, max(m_loc, 1)), -1.0, dtype=torch.float64, device=abi.dev)
            anc, cnt = abi.zeros(max(m_loc, 1), dtype=torch.int32), abi.zeros(nl, dtype=torch.int32)
            filled = abi.zeros(1, dtype=torch.int64)
            abi.ck(abi.lib.smcb_resample_fused(abi.h, None, wt.data_ptr(), nl, n, plan["carry_q"][r], r * nl, m_loc, None, 0.0,
                                               None, u0, shard.data_ptr(), nl, D1, dst.data_ptr(), max(m_loc, 1),
                                               anc.data_ptr(), cnt.data_ptr(), filled.data_ptr(), None))
            torch.cuda.synchronize()
            assert np.array_equal(cnt.cpu().numpy(), c_full[r * nl:(r + 1) * nl])
            out_rows.append(dst[:, :m_loc].cpu().numpy())
            out_anc.append(anc[:m_loc].cpu().numpy().astype(np.int64) + r * nl)
        assert np.array_equal(np.concatenate(out_anc), a_full)
        assert np.array_equal(np.concatenate(out_rows, axis=1), src.cpu().numpy()[:, a_full])


def test_resample_golden_stage_weights(abi, golden):
    """Weights of the reference's first stage (from its own likelihoods) and its own rand() draw."""
    lk = golden["sweeps_out"][0]
    t = smc.temper_backoff(lk, 0.0, smc.Settings())
    u0 = float(golden["draws_rand"][0])
    a_ref, c_ref, _ = smc.resample_sequential(t["p_weight"], u0)
    for mode in (0, 1):
        c, a, _, _ = _resample(abi, t["p_weight"], u0, mode)
        assert np.array_equal(c, c_ref) and np.array_equal(a, a_ref)
    assert len(np.unique(a_ref)) == 783           # SURVEY.md 6.2, stage 1 distinct ancestors


def test_resample_degenerate_and_padding(abi):
    w = np.zeros(3000)
    w[1717] = 1.0
    for mode in (0, 1):
        c, a, _, filled = _resample(abi, w, 0.5, mode)
        assert c[1717] == 3000 and c.sum() == 3000 and np.all(a == 1717)
    # under-filled output is padded with the last ancestor, over-filled output is cut
    counts = np.array([2, 0, 1, 0], dtype=np.int32)
    ct = abi.t(counts, torch.int32)
    for m, want in ((6, [0, 0, 2, 2, 2, 2]), (2, [0, 0]), (3, [0, 0, 2])):
        anc = abi.zeros(m, dtype=torch.int32)
        filled = abi.zeros(1, dtype=torch.int64)
        abi.ck(abi.lib.smcb_ancestors(abi.h, ct.data_ptr(), 4, m, anc.data_ptr(), filled.data_ptr(), None))
        assert anc.cpu().tolist() == want and filled.item() == 3


def test_resample_fixed_is_shard_invariant(abi):
    """Counts computed shard by shard with the carried fixed-point prefix equal the unsharded counts."""
    n, W = 1 << 16, 4
    w = _weights(n, 11)
    u0 = 0.6180339887
    c_full, _, tot_full, _ = _resample(abi, w, u0, 1)
    nl = n // W
    carry, got = 0, []
    for r in range(W):
        wt = abi.t(w[r * nl:(r + 1) * nl])
        tot = abi.zeros(2, dtype=torch.int64)
        abi.ck(abi.lib.smcb_resample_totals(abi.h, wt.data_ptr(), nl, n, tot.data_ptr(), None))
        c, _, tot2, _ = _resample(abi, w[r * nl:(r + 1) * nl], u0, 1, n_total=n, carry_q=carry, id_offset=r * nl)
        assert np.array_equal(tot.cpu().numpy(), tot2)
        carry += int(tot2[1])
        got.append(c)
    assert np.array_equal(np.concatenate(got), c_full) and carry == tot_full[1]


def test_resample_full_size_properties(abi):
    """2^20 particles (BASELINE config 2 size): size-independent properties."""
    n = 1 << 20
    rs = np.random.RandomState(5)
    lk = rs.normal(0, 3, n)
    w = np.exp(lk - lk.max())
    w /= w.sum()
    c, a, tot, filled = _resample(abi, w, 0.123456789, 1)
    assert abs(filled - n) <= 1 and c.sum() == filled
    assert np.all(np.diff(a) >= 0) and a.min() >= 0 and a.max() < n
    assert np.array_equal(np.bincount(a, minlength=n)[: n - 1], c[: n - 1])
    fl = np.trunc(w * n)
    assert np.all(c >= fl) and np.all(c <= fl + 1)
    # sequential mode agrees except at ~1e-13 near-ties
    c2, a2, _, _ = _resample(abi, w, 0.123456789, 0)
    assert (c != c2).sum() <= 2
    # gather: checksum of checksums
    D1 = 4
    src = torch.arange(D1 * n, dtype=torch.float64, device=abi.dev).reshape(D1, n)
    dst = torch.zeros_like(src)
    at = abi.t(a, torch.int32)
    abi.ck(abi.lib.smcb_gather(abi.h, src.data_ptr(), n, at.data_ptr(), n, D1, dst.data_ptr(), n, None))
    want = src[:, at.long()]
    assert torch.equal(dst, want)


# ------------------------------------------------------------------------------------ K4 MH
@pytest.mark.parametrize("d,n", [(1, 100), (3, 1000), (5, 4097), (8, 3000), (32, 2500), (32, (1 << 16) + 5),
                                 (26, 1 << 15)])
def test_moments_match_numpy_cov(abi, d, n):
    rs = np.random.RandomState(d)
    X = rs.normal(0, 1, (n, d)) @ rs.normal(0, 1, (d, d)) + rs.uniform(1e3, 1e6, d)   # large offsets (H4)
    th = abi.t(X.T)
    mom = abi.zeros(d + d * d)
    abi.ck(abi.lib.smcb_colsum(abi.h, th.data_ptr(), n, n, d, mom.data_ptr(), None))
    mean = mom[:d] / n
    assert _rel(mean.cpu().numpy(), X.mean(axis=0)).max() < 1e-13
    mom[:d] = mean
    abi.ck(abi.lib.smcb_centered_moments(abi.h, th.data_ptr(), n, n, d, mom.data_ptr(), mom[d:].data_ptr(), None))
    cov = mom[d:].cpu().numpy().reshape(d, d) / n
    want = np.atleast_2d(np.cov(X.T, bias=True))
    assert np.abs(cov - want).max() <= 1e-10 * np.abs(want).max()
    assert np.array_equal(cov, cov.T)


@pytest.mark.parametrize("d,n", [(1, 100), (3, 1000), (3, (1 << 18) + 1), (5, 4097), (32, 70000)])
def test_merged_moments_and_device_factor(abi, d, n):
    """smcb_moments_merged on one shard: counters pass through, mean and M2 are bit-identical to smcb_colsum / n and
    smcb_centered_moments, the device's Jacobi factor satisfies F^T F = cov (*) w_cov and equals the oracle's twin."""
    rs = np.random.RandomState(100 + d)
    scale = 10.0 ** rs.uniform(-2, 6, d)                      # parameters of very different magnitude (methanation)
    X = (rs.normal(0, 1, (n, d)) @ rs.normal(0, 1, (d, d))) * scale + rs.uniform(1e3, 1e6, d)
    th = abi.t(X.T)
    mom = abi.zeros(d + d * d)
    abi.ck(abi.lib.smcb_colsum(abi.h, th.data_ptr(), n, n, d, mom.data_ptr(), None))
    mean = abi.t(mom[:d].cpu().numpy() / n)                  # IEEE division, as np.cov's X.mean() (torch's tensor / scalar
    mom[:d] = mean                                           # on the GPU multiplies by the reciprocal: 1 ulp off)
    abi.ck(abi.lib.smcb_centered_moments(abi.h, th.data_ptr(), n, n, d, mom.data_ptr(), mom[d:].data_ptr(), None))
    cnt = abi.t(np.array([5, 7, 11, 13, 0, 0, 0, 0]), torch.int64)
    w = np.full((d, d), 0.5)
    blk = abi.zeros(4 + d + 2 * d * d)
    abi.ck(abi.lib.smcb_moments_merged(abi.h, th.data_ptr(), n, n, d, n, cnt.data_ptr(), w.ctypes.data, blk.data_ptr(), None))
    b = blk.cpu().numpy()
    assert np.array_equal(b[:4], [5, 7, 11, 13])
    assert np.array_equal(b[4:4 + d], mean.cpu().numpy())
    assert np.array_equal(b[4 + d:4 + d + d * d], mom[d:].cpu().numpy())
    cov = b[4 + d:4 + d + d * d].reshape(d, d) / n * w
    F = b[4 + d + d * d:].reshape(d, d)
    sd = np.sqrt(np.diag(cov))
    assert np.abs((F.T @ F - cov) / np.outer(sd, sd)).max() < 1e-12
    Fo = smc.proposal_factor_eig(cov)
    assert np.abs((F - Fo) / sd[None, :]).max() < 1e-9


def test_merged_moments_degenerate_clouds(abi):
    """One particle, two particles, a constant column: zero or rank-deficient covariance gives zero factor rows, not NaN."""
    for X in (np.array([[1.0, 2.0, 3.0]]), np.array([[1.0, 2.0, 3.0], [2.0, 2.5, 3.0]]),
              np.c_[np.arange(50.0), np.full(50, 7.0), np.arange(50.0) ** 2]):
        n, d = X.shape
        th = abi.t(X.T)
        blk = abi.zeros(4 + d + 2 * d * d)
        abi.ck(abi.lib.smcb_moments_merged(abi.h, th.data_ptr(), n, n, d, n, None, None, blk.data_ptr(), None))
        b = blk.cpu().numpy()
        F = b[4 + d + d * d:].reshape(d, d)
        cov = np.atleast_2d(np.cov(X.T, bias=True))
        assert np.all(np.isfinite(F)) and np.abs(F.T @ F - cov).max() <= 1e-12 * max(np.abs(cov).max(), 1.0)
        assert np.abs(F - smc.proposal_factor_eig(cov)).max() < 1e-9


def test_temper_eval_equals_max_then_sums(abi):
    """smcb_temper_eval without a communicator: bit-identical to smcb_lk_max + smcb_temper_sums, for up to 48 candidates."""
    n = 100003
    lk = np.random.RandomState(9).normal(300, 80, n)
    lk[5] = -np.inf
    t = abi.t(lk)
    gms = np.ascontiguousarray([0.7 ** k for k in range(40)])
    a, b = abi.zeros(2 + 96), abi.zeros(2 + 96)
    abi.ck(abi.lib.smcb_temper_eval(abi.h, t.data_ptr(), n, gms.ctypes.data, 40, a.data_ptr(), None))
    abi.ck(abi.lib.smcb_lk_max(abi.h, t.data_ptr(), n, b.data_ptr(), None))
    for o in range(0, 40, 16):
        g = np.ascontiguousarray(gms[o:o + 16])
        abi.ck(abi.lib.smcb_temper_sums(abi.h, t.data_ptr(), n, b.data_ptr(), g.ctypes.data, len(g),
                                        b[2 + 2 * o:].data_ptr(), None))
    assert a[0].item() == lk.max()
    assert torch.equal(a[2:82], b[2:82])


def test_proposal_with_device_factor_equals_host_factor(abi):
    n, d, seed = 3001, 5, 3
    rs = np.random.RandomState(2)
    X = rs.normal(0, 1, (n, d))
    F = np.ascontiguousarray(rs.normal(0, 0.3, (d, d)))
    low, high = np.full(d, -1.5), np.full(d, 1.5)
    th, Fd = abi.t(X.T), abi.t(F)
    outs = []
    for dev in (False, True):
        prop, inbox = abi.zeros(d, n), abi.zeros(n, dtype=torch.uint8)
        fn = abi.lib.smcb_mh_propose_dev if dev else abi.lib.smcb_mh_propose
        abi.ck(fn(abi.h, th.data_ptr(), n, n, d, Fd.data_ptr() if dev else F.ctypes.data, 0.5, low.ctypes.data,
                  high.ctypes.data, None, seed, 0, 2, 1, prop.data_ptr(), n, inbox.data_ptr(), None))
        outs.append((prop.cpu().numpy(), inbox.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert 0 < outs[0][1].sum() < n


def test_philox_draws_match_numpy_twin(abi):
    n, d, seed = 5000, 5, 20250205
    z, u = abi.zeros(n, d), abi.zeros(n)
    abi.ck(abi.lib.smcb_philox_draws(abi.h, n, d, seed, 1 << 33, 7, 3, z.data_ptr(), u.data_ptr(), None))
    ids = np.arange(n, dtype=np.uint64) + np.uint64(1 << 33)
    assert np.array_equal(u.cpu().numpy(), philox.uniforms(seed, ids, 7, 3))
    assert np.abs(z.cpu().numpy() - philox.normals(seed, ids, 7, 3, d)).max() < 1e-13
    th = abi.zeros(3, n)
    low, high = np.array([0.0, -5.0, 2.0]), np.array([10.0, 5.0, 2.5])
    abi.ck(abi.lib.smcb_sample_uniform_box(abi.h, th.data_ptr(), n, n, 3, low.ctypes.data, high.ctypes.data, seed,
                                           1 << 33, None))
    assert np.array_equal(th.cpu().numpy().T, philox.uniform_box(seed, ids, low, high))


@pytest.mark.parametrize("d", [3, 5, 32])
def test_propose_and_accept_match_oracle(abi, d):
    """Same particles, same normals, same uniforms -> same proposals, same accept decisions."""
    n = 3001
    rs = np.random.RandomState(d)
    X = rs.uniform(0, 10, (n, d))
    lk1 = rs.normal(100, 5, n)
    A = rs.normal(0, 1, (d, d))
    F = smc.proposal_factor(A @ A.T / d * 0.5)
    Z, U = rs.standard_normal((n, d)), rs.uniform(0, 1, n)
    low, high = np.zeros(d), np.full(d, 10.0)
    lk2_all = rs.normal(100, 5, n)
    gamma, ratio = 0.37, 0.5

    def fake_loglik(p):
        return lk2_all

    p_new, lk_new, r, n_in = smc.mh_sweep(X, lk1, gamma, F, Z, U, ratio, fake_loglik, low, high)
    step = np.dot(Z, F)
    prop_ref = X + step * ratio
    inbox_ref = smc.in_box(prop_ref, low, high)
    prop_ref = np.where(inbox_ref[:, None], prop_ref, X)

    th, lk, prop = abi.t(X.T), abi.t(lk1), abi.zeros(d, n)
    inbox, moved = abi.zeros(n, dtype=torch.uint8), abi.zeros(n, dtype=torch.uint8)
    cnt = abi.zeros(4, dtype=torch.int64)
    Fc = np.ascontiguousarray(F)
    Zt, Ut, lk2t = abi.t(Z), abi.t(U), abi.t(lk2_all)      # keep the device copies alive across the launches
    abi.ck(abi.lib.smcb_mh_propose(abi.h, th.data_ptr(), n, n, d, Fc.ctypes.data, ratio, low.ctypes.data,
                                   high.ctypes.data, Zt.data_ptr(), 1, 0, 1, 0, prop.data_ptr(), n,
                                   inbox.data_ptr(), None))
    assert np.array_equal(inbox.cpu().numpy().astype(bool), inbox_ref)
    got_prop = prop.cpu().numpy().T
    # z @ F is accumulated in the same order as np.dot for d<=8 up to FMA contraction
    assert np.abs(got_prop - prop_ref).max() < 1e-13 * 10
    abi.ck(abi.lib.smcb_mh_accept(abi.h, th.data_ptr(), n, lk.data_ptr(), prop.data_ptr(), n, lk2t.data_ptr(),
                                  inbox.data_ptr(), n, d, gamma, Ut.data_ptr(), None, 1, 0, 1, 0, moved.data_ptr(),
                                  cnt.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(moved.cpu().numpy(), r.astype(np.uint8))
    c = cnt.cpu().numpy()
    assert c[0] == r.sum() and c[1] == r.sum() and c[2] == n_in
    assert np.array_equal(lk.cpu().numpy(), lk_new)
    assert np.abs(th.cpu().numpy().T - p_new).max() < 1e-12


# ------------------------------------------------------------------------------------ error behaviour
def test_error_codes_and_messages(mm_abi):
    """Every entry point returns a negative SMCB_ERR_* code with a message instead of raising or crashing
    (the counterpart of the reference's blanket try/except, Micmem_SMC_main.py:305-314)."""
    import ctypes as C
    lib, h = mm_abi.lib, mm_abi.h
    n = 64
    th, lk = mm_abi.zeros(3, n), mm_abi.zeros(n)
    th += 1.0

    def msg():
        return lib.smcb_last_error(h).decode()

    assert lib.smcb_loglik(h, 99, th.data_ptr(), n, n, 3, None, lk.data_ptr(), None) == -1 and "unknown model" in msg()
    assert lib.smcb_loglik(h, 1, th.data_ptr(), n, n, 5, None, lk.data_ptr(), None) == -1 and "d=3" in msg()
    assert lib.smcb_loglik(h, 1, None, n, n, 3, None, lk.data_ptr(), None) == -1 and "null" in msg()
    assert lib.smcb_loglik(h, 1, th.data_ptr(), n - 1, n, 3, None, lk.data_ptr(), None) == -1      # ld < n
    assert lib.smcb_loglik(h, 1, th.data_ptr(), 1 << 22, 1 << 22, 3, None, lk.data_ptr(), None) == -3 \
        and "smcb_reserve" in msg()                                                              # beyond the reserve
    assert lib.smcb_set_param(h, 12345, 1.0) == -1 and "unknown key" in msg()
    assert lib.smcb_set_param(h, 1, 0.0) == -1                                                    # budget < 1
    t = np.array([[0.0, 1.0, 1.0]])
    assert lib.smcb_set_data_mm_progress(h, t.ctypes.data, t.ctypes.data, t.ctypes.data, 1, 3) == -1 \
        and "strictly increasing" in msg()
    out = mm_abi.zeros(4)
    assert lib.smcb_temper_sums(h, lk.data_ptr(), n, out.data_ptr(), t.ctypes.data, 99, out.data_ptr(), None) == -1
    bad = C.c_void_p()
    assert lib.smcb_create(4096, C.byref(bad)) == -1 and not bad.value                            # no such device
    assert b"out of range" in lib.smcb_last_error(None)
    # the handle is still usable after errors
    assert lib.smcb_loglik(h, 1, th.data_ptr(), n, n, 3, None, lk.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert np.all(np.isfinite(lk.cpu().numpy()))
