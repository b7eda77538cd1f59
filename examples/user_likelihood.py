#!/usr/bin/env python
"""A likelihood the library was not compiled with, two ways.

The reference's plug-in is a user-written `sim_particle` (`SMC_example/Micmem_likelihood.py:79-92`): to change the
model one edits that function.  Here the model is either

  1. a CUDA functor in a file of your own (`examples/user_gauss.cu`, 60 lines against `include/smcb_user.cuh`),
     compiled with nvcc into a shared library and registered through `smcb_set_user_likelihood`, or
  2. any Python callable on device tensors (torch operations), registered through the same entry point.

Tempering, resampling and the MH mutation run unchanged on the library's kernels either way.

    python examples/user_likelihood.py [kernel|callable]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200  # noqa: E402


def main(kind="kernel"):
    import torch
    rs = np.random.RandomState(4)
    x = np.linspace(0.0, 4.0, 60)
    y = 1.5 + 0.7 * x + 0.3 * rs.standard_normal(60)
    prior = smcb200.UniformBox([-5, -5, 0], [5, 5, 5], names=["mu", "slope", "sigma"])
    if kind == "kernel":
        so = smcb200.build_user_library(os.path.join(ROOT, "examples", "user_gauss.cu"))     # nvcc, sm_100a
        dll = C.CDLL(so)
        assert dll.user_gauss_set_data(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), len(x)) == 0
        lik = smcb200.UserKernelLikelihood(dll, "user_gauss_loglik", d=3, n_obs=len(x), names=prior.names)
    else:
        xd, yd = torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda")

        def loglik(theta, active, lk_out):          # theta [3, n] on the device
            r = yd[None, :] - (theta[0][:, None] + theta[1][:, None] * xd[None, :])
            ll = -0.5 * len(x) * torch.log(2 * torch.pi * theta[2] ** 2) - (r * r).sum(1) / (2 * theta[2] ** 2)
            return torch.where(theta[2] > 0, ll, torch.full_like(ll, -float("inf")))

        lik = smcb200.CallableLikelihood(loglik, d=3, n_obs=len(x), names=prior.names)
    res = smcb200.run(lik, prior, settings=smcb200.Settings(n_particle=1 << 16))
    print(f"{kind}: {len(res.betas)} stages, log-evidence {res.log_evidence:.4f}, posterior mean "
          f"{dict(zip(prior.names, np.round(res.particles.mean(0), 4)))} (truth: 1.5, 0.7, 0.3), {res.seconds * 1e3:.1f} ms")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "kernel")
