"""A likelihood the library was not compiled with (the reference's plug-in is a user-written
`sim_particle`, SMC_example/Micmem_likelihood.py:79-92): a user-compiled CUDA kernel behind `smcb_set_user_likelihood`
and a Python callable on device tensors, each driving the unchanged tempering / resampling / MH kernels, against the
oracle loop with the same model in NumPy."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import smc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "user_gauss.cu")


def _data(m=60, seed=4):
    rs = np.random.RandomState(seed)
    x = np.linspace(0.0, 4.0, m)
    y = 1.5 + 0.7 * x + 0.3 * rs.standard_normal(m)
    return x, y


def _numpy_loglik(x, y):
    def ll(theta):
        mu, slope, sigma = theta[:, 0:1], theta[:, 1:2], theta[:, 2]
        r = y[None, :] - (mu + slope * x[None, :])
        with np.errstate(divide="ignore", invalid="ignore"):
            out = -0.5 * len(x) * np.log(2 * np.pi * sigma ** 2) - (r * r).sum(1) / (2 * sigma ** 2)
        return np.where(sigma > 0, out, -np.inf)
    return ll


def test_user_example_source_uses_only_the_public_headers():
    """CPU: the example includes nothing but include/smcb_user.cuh (which includes include/smcb200.h)."""
    src = open(SRC).read()
    incs = [ln.split('"')[1] for ln in src.splitlines() if ln.startswith('#include "')]
    assert incs == ["smcb_user.cuh"]
    hdr = open(os.path.join(ROOT, "include", "smcb_user.cuh")).read()
    assert '#include "smcb200.h"' in hdr and "smcb_user_loglik_fn" in open(os.path.join(ROOT, "include", "smcb200.h")).read()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["kernel", "callable"])
def test_user_likelihood_runs_the_sampler_and_matches_the_oracle_loop(pkg, kind):
    import torch
    x, y = _data()
    N, seed = 4096, 17
    prior = pkg.UniformBox([-5, -5, 0], [5, 5, 5], names=["mu", "slope", "sigma"])
    if kind == "kernel":
        so = pkg.build_user_library(SRC)
        dll = C.CDLL(so)
        xc, yc = np.ascontiguousarray(x), np.ascontiguousarray(y)
        assert dll.user_gauss_set_data(C.c_void_p(xc.ctypes.data), C.c_void_p(yc.ctypes.data), len(x)) == 0
        lik = pkg.UserKernelLikelihood(dll, "user_gauss_loglik", d=3, n_obs=len(x))
    else:
        xd, yd = torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda")

        def fn(theta, active, lk_out):
            mu, slope, sigma = theta[0], theta[1], theta[2]
            r = yd[None, :] - (mu[:, None] + slope[:, None] * xd[None, :])
            ll = -0.5 * len(x) * torch.log(2 * torch.pi * sigma ** 2) - (r * r).sum(1) / (2 * sigma ** 2)
            return torch.where(sigma > 0, ll, torch.full_like(ll, -float("inf")))

        lik = pkg.CallableLikelihood(fn, d=3, n_obs=len(x))
    eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=N, seed=seed))
    eng.sample_prior()
    p0 = eng.particles().cpu().numpy()
    lk0 = eng.sim_particle().cpu().numpy().copy()
    ll = _numpy_loglik(x, y)
    assert np.abs(lk0 / ll(p0) - 1).max() < 1e-11
    res = eng.run(keep_ancestors=True)
    p, lk, tr = smc.run(ll, p0, prior.low, prior.high, smc.Settings(n_particle=N), smc.PhiloxStream(seed),
                        resampler=smc.resample_fixed, factor=smc.proposal_factor_eig)
    assert res.reached_one and np.array_equal(np.array(res.betas), np.array(tr.gamma))
    assert res.n_mh == tr.n_mh and res.n_moved == tr.moved
    for a, b in zip(res.ancestors, tr.ancestors):
        assert np.array_equal(a, b)
    assert np.abs(res.particles - p).max() < 1e-9 and np.abs(res.lk / lk - 1).max() < 1e-9
    assert abs(res.log_evidence / tr.log_evidence[-1] - 1) < 1e-9
    m = res.particles.mean(0)
    assert abs(m[0] - 1.5) < 0.3 and abs(m[1] - 0.7) < 0.15 and abs(m[2] - 0.3) < 0.1
    eng.close()


@pytest.mark.gpu
def test_user_callback_errors_surface_as_exceptions(pkg):
    def bad(theta, active, lk_out):
        raise ValueError("boom")

    lik = pkg.CallableLikelihood(bad, d=3)
    eng = pkg.Engine(lik, pkg.UniformBox([0, 0, 0], [1, 1, 1]), pkg.Settings(n_particle=64))
    eng.sample_prior()
    with pytest.raises(pkg._lib.SmcbError) as ei:
        eng.sim_particle()
    assert ei.value.code == -6 and isinstance(lik.error, ValueError)
    eng.close()
