"""The DEVICE solver arithmetic, compiled for the host, against scipy (CPU only, no GPU needed).

`csrc/mm_solver.cuh` is the header the CUDA kernels call for every attempted RK45 step.  It also compiles
with plain g++ (`tests/host_twin.cpp`): same FMA chains, same reciprocal / root corrections, only the
hardware seeds (MUFU.RCP64H, lg2/ex2) are replaced by truncated host values which the corrections make
irrelevant.  These tests pin that arithmetic against the likelihoods the unmodified reference computed
(golden fixture) and against the operation-for-operation C twin of scipy's RK45."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def twin(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("twin") / "libhost_twin.so")
    subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                    os.path.join(HERE, "host_twin.cpp")], check=True)
    lib = C.CDLL(so)

    def loglik(theta, t, P, S0, fn="twin_loglik"):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        t, P, S0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (t, P, S0))
        n, (n_ex, n_t) = len(th), t.shape
        lk, cnt, steps = np.empty(n), np.zeros(4, dtype=np.int64), np.zeros((n, n_ex), dtype=np.int32)
        getattr(lib, fn)(C.c_void_p(th.ctypes.data), C.c_int64(n), C.c_void_p(t.ctypes.data), C.c_void_p(P.ctypes.data),
                        C.c_void_p(S0.ctypes.data), C.c_int(n_ex), C.c_int(n_t), C.c_void_p(lk.ctypes.data),
                        C.c_void_p(cnt.ctypes.data), C.c_void_p(steps.ctypes.data))
        return lk, cnt, steps

    def predict(theta, t, S0):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        t, S0 = np.ascontiguousarray(t, dtype=np.float64), np.ascontiguousarray(S0, dtype=np.float64)
        n, (n_ex, n_t) = len(th), t.shape
        out = np.zeros((n, n_ex, n_t))
        lib.twin_predict(C.c_void_p(th.ctypes.data), C.c_int64(n), C.c_void_p(t.ctypes.data), C.c_void_p(S0.ctypes.data),
                         C.c_int(n_ex), C.c_int(n_t), C.c_void_p(out.ctypes.data))
        return out

    def loglik_parked(theta, t, P, S0, budget):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        t, P, S0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (t, P, S0))
        n, (n_ex, n_t) = len(th), t.shape
        lk, cnt, steps = np.empty(n), np.zeros(4, dtype=np.int64), np.zeros((n, n_ex), dtype=np.int32)
        lib.twin_loglik_parked(C.c_void_p(th.ctypes.data), C.c_int64(n), C.c_void_p(t.ctypes.data), C.c_void_p(P.ctypes.data),
                               C.c_void_p(S0.ctypes.data), C.c_int(n_ex), C.c_int(n_t), C.c_int(budget),
                               C.c_void_p(lk.ctypes.data), C.c_void_p(cnt.ctypes.data), C.c_void_p(steps.ctypes.data))
        return lk, cnt, steps

    loglik.parked = loglik_parked
    return loglik, predict


def test_device_arithmetic_matches_every_reference_sweep(twin, golden):
    """All 34 sweeps x 1000 particles the reference itself evaluated (prior cloud ... posterior)."""
    loglik, _ = twin
    d = (golden["data_t"], golden["data_P"], golden["data_S0"])
    worst = 0.0
    for sweep in range(golden["sweeps_in"].shape[0]):
        got, cnt, _ = loglik(golden["sweeps_in"][sweep], *d)
        want = golden["sweeps_out"][sweep]
        worst = max(worst, float(np.max(np.abs(got - want) / np.abs(want))))
        assert cnt[3] == 0
    assert worst < 1e-9, worst            # bar: 1e-5 (BASELINE.json north_star)


def test_device_arithmetic_takes_scipys_steps(twin, golden):
    """Same accepted/rejected step counts as the strict-order C twin of scipy's RK45 on a prior cloud
    (heavy-tailed: 10 ... >1e4 attempts per solve), solve by solve."""
    from oracle import cmm
    loglik, _ = twin
    d = (golden["data_t"], golden["data_P"], golden["data_S0"])
    th = np.random.RandomState(1).uniform(0, 10, (1 << 14, 3))
    got, cnt, steps = loglik(th, *d)
    want, info = cmm.loglik_progress(th, *d, want_steps=True)
    assert np.array_equal(steps, info["steps"])
    assert cnt[1] == info["accepted"] and cnt[2] == info["rejected"] and cnt[0] == info["nfev"]
    assert steps.max() > 2000                                  # the cloud does contain stiff solves
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-9


def test_tail_kernel_spelling_takes_scipys_steps(twin, golden):
    """mmsolve::solve_lat (the tail kernel: plain steps in the latency spelling, steps that touch an observation time
    through attempt()) against the C twin of scipy's RK45 and against attempt() alone: the same accepted / rejected
    steps solve by solve - the two stiffest solves of the bench's 2^20-particle prior cloud included (83 133 and
    62 913 attempts) - and log-likelihoods that differ by rounding only."""
    from oracle import cmm
    loglik, _ = twin
    d = (golden["data_t"], golden["data_P"], golden["data_S0"])
    th = np.concatenate([np.random.RandomState(1).uniform(0, 10, (1 << 14, 3)),
                         [[9.88869508e+00, 4.28267643e-04, 4.53281765e+00], [7.30644405, 4.35543917e-04, 1.0]],
                         golden["sweeps_in"][33][:64]])
    std, cnt0, steps0 = loglik(th, *d)
    got, cnt, steps = loglik(th, *d, fn="twin_loglik_lat")
    want, info = cmm.loglik_progress(th, *d, want_steps=True)
    assert np.array_equal(steps, info["steps"]) and np.array_equal(steps, steps0)
    assert cnt[1] == info["accepted"] and cnt[2] == info["rejected"] and cnt[3] == 0
    assert steps.max() == 83133
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-7         # 6e-9 observed on the stiffest solves; bar 1e-5
    assert np.max(np.abs(got - std) / np.abs(std)) < 1e-11          # the two spellings: rounding only


@pytest.mark.parametrize("budget", [1, 7, 64, 512])
def test_parked_solves_resume_where_they_stopped(twin, golden, budget):
    """The hand-over between the two kernels (mm_bulk_kernel parks a solve that exceeds its budget, mm_tail_kernel
    resumes it: mmsolve::park_store / park_load): whatever the budget, a solve takes the steps of the C twin of scipy's
    RK45 - before, across and after the hand-over - and the log-likelihood moves by rounding only."""
    from oracle import cmm
    loglik, _ = twin
    d = (golden["data_t"], golden["data_P"], golden["data_S0"])
    th = np.concatenate([np.random.RandomState(2).uniform(0, 10, (4096, 3)), golden["sweeps_in"][33][:64],
                         [[7.30644405, 4.35543917e-04, 1.0]]])
    th[:, 2] = np.maximum(th[:, 2], 1e-3)
    base, _, steps0 = loglik(th, *d)
    got, cnt, steps = loglik.parked(th, *d, budget)
    want, info = cmm.loglik_progress(th, *d, want_steps=True)
    assert cnt[3] == 0
    assert np.array_equal(steps, info["steps"]) and np.array_equal(steps, steps0)
    assert cnt[1] == info["accepted"] and cnt[2] == info["rejected"]
    assert (steps > budget).any()                                    # some solves were handed over
    assert np.max(np.abs(got - base) / np.abs(base)) < 1e-11
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-7


def test_device_predictions_match_reference(twin, golden):
    _, predict = twin
    got = predict(golden["prior_particles"][:8], golden["data_t"], golden["data_S0"])
    assert np.abs(got - golden["pmodel0"]).max() < 1e-11


def test_known_answers_and_degenerate_parameters(twin, golden):
    loglik, _ = twin
    d = (golden["data_t"], golden["data_P"], golden["data_S0"])
    th = np.array([[1.2, 0.5, 0.02], [1.0, 0.4, 0.05], [1.0, 0.4, 0.0], [1.0, 0.4, -1.0], [1.0, -2.0, 1.0]])
    got, _, _ = loglik(th, *d)
    assert abs(got[0] - 593.9635684697922) < 1e-7 and abs(got[1] - 424.5676547271564) < 1e-7
    assert got[2] == -np.inf and got[3] == -np.inf             # sigma <= 0 (Micmem_likelihood.py:53-54)
    assert got[4] == -np.inf                                   # Km + S0 = 0 for the S0 = 2 curves: no valid first step


def test_arrhenius_exp_is_within_an_ulp_of_libm(tmp_path):
    """csrc/exp_table.cuh (table + degree-5 polynomial, used by the kinetic reactor march) against NumPy's exp."""
    so = str(tmp_path / "libhost_exp.so")
    subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                    os.path.join(HERE, "host_exp.cpp")], check=True)
    lib = C.CDLL(so)
    rs = np.random.RandomState(0)
    x = np.concatenate([rs.uniform(-700, 700, 400000), rs.uniform(-60, 5, 400000), rs.normal(0, 1e-3, 1000),
                        np.array([0.0, -0.0, 1e-300, -1e-300, 699.999, -699.999, 707.999, -707.999, 708.0, -708.0, 709.7, -745.0, 800.0, 1e300, -1e300,
                                  -800.0, np.inf, -np.inf, np.nan]),
                        np.arange(-3000, 3000) * (np.log(2) / 128)])          # the reduction's break points
    y = np.empty_like(x)
    lib.exp_fast_host(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_long(len(x)))
    with np.errstate(over="ignore", under="ignore"):
        want = np.exp(x)
    want[x >= 708.0] = np.inf                      # the device function's documented range
    want[x <= -708.0] = 0.0
    assert np.array_equal(np.isnan(y), np.isnan(want))
    fin = np.isfinite(want) & (want > 0)
    assert np.array_equal(y[~fin & ~np.isnan(want)], want[~fin & ~np.isnan(want)])
    ulp = np.abs(y[fin] - want[fin]) / np.spacing(want[fin])
    assert ulp.max() <= 1.0, ulp.max()             # never more than one ulp away from libm
    assert np.mean(ulp == 0) > 0.7                 # and identical to it three times out of four
