"""Likelihood specifications (the data a model needs on the device).

The reference's "plugin" is `sim_particle(particle) -> (llk, C_l_)` plus module globals holding the
data (`SMC_example/Micmem_likelihood.py:79-92`, `Micmem_settings.py:103-115`;
`SMC_methanation/methanation_functions.py:70-92`).  Here a likelihood is a small object that
uploads its data once; the engine then evaluates all particles with one kernel launch.
"""
import os

import numpy as np

from . import _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class MMProgress:
    """Michaelis-Menten progress curves, scipy-RK45 arithmetic (Micmem_likelihood.py:14-77)."""
    model_id = _lib.MODEL_MM_PROGRESS
    d = 3
    names = ("Vmax", "Km", "sigma")

    def __init__(self, t, P_obs, S0, integrator="rk45_scipy"):
        """integrator = "rk45_scipy": scipy's adaptive RK45 step for step (the reference's likelihood; parity mode).
        integrator = "exact": the closed form S(t) = Km * wrightomega(ln(S0/Km) + (S0 - Vmax t)/Km), the converged
        solution of the same ODE (converged mode: cost independent of stiffness; NOT the reference's rtol-1e-3
        numbers)."""
        self.t, self.P_obs, self.S0 = _f64(t), _f64(P_obs), _f64(S0)
        if self.t.ndim != 2 or self.t.shape != self.P_obs.shape or self.S0.shape != (self.t.shape[0],):
            raise ValueError("t, P_obs must be [n_ex, n_t] and S0 [n_ex]")
        if integrator not in ("rk45_scipy", "exact"):
            raise ValueError("integrator must be 'rk45_scipy' or 'exact'")
        self.integrator = integrator

    @classmethod
    def from_csv(cls, base_path="data/mm_pseudo_data", n_ex=6):
        """Loads `{base_path}_{i}.csv` with columns t,S_true,P_true,P_obs exactly as
        Micmem_settings.py:103-115 does (S0 = first S_true)."""
        import pandas as pd
        t, P, S0 = [], [], []
        for i in range(n_ex):
            df = pd.read_csv(f"{base_path}_{i}.csv")
            t.append(df["t"].values)
            P.append(df["P_obs"].values)
            S0.append(df["S_true"].iloc[0])
        return cls(np.array(t), np.array(P), np.array(S0))

    @property
    def n_obs(self):
        return self.t.size

    def upload(self, lib, handle):
        _lib.check(handle, lib.smcb_set_data_mm_progress(
            handle, self.t.ctypes.data, self.P_obs.ctypes.data, self.S0.ctypes.data,
            self.t.shape[0], self.t.shape[1]))
        _lib.check(handle, lib.smcb_set_param(handle, _lib.PARAM_MM_INTEGRATOR,
                                              float(_lib.MM_EXACT if self.integrator == "exact" else _lib.MM_RK45_SCIPY)))


class MMRate:
    """Rate-law observations (S_i, v_i), v ~ N(Vmax*S/(Km+S), sigma^2)."""
    model_id = _lib.MODEL_MM_RATE
    d = 3
    names = ("Vmax", "Km", "sigma")

    def __init__(self, S, v, precision=64, form="direct", km_range=(0.0, 10.0)):
        """form = "direct": every likelihood sums over the observations (FP64 or FP32 terms, `precision`);
        form = "sufficient": sum v^2, A(Km) = sum v S/(Km+S) and B(Km) = sum S^2/(Km+S)^2 are tabulated once over
        km_range (the prior's Km interval) and a likelihood costs O(1) (smcb_set_data_mm_rate_sufficient)."""
        self.S, self.v, self.precision = _f64(S), _f64(v), int(precision)
        if self.S.ndim != 1 or self.S.shape != self.v.shape:
            raise ValueError("S and v must be 1-D arrays of equal length")
        if form not in ("direct", "sufficient"):
            raise ValueError("form must be 'direct' or 'sufficient'")
        self.form, self.km_range = form, (float(km_range[0]), float(km_range[1]))

    @classmethod
    def synthetic(cls, n_obs=10000, Vmax=1.2, Km=0.5, sigma=0.02, seed=20250205, precision=64, form="direct",
                  km_range=(0.0, 10.0)):
        rs = np.random.RandomState(seed)
        S = np.exp(rs.uniform(np.log(0.05), np.log(20.0), n_obs))
        v = Vmax * S / (Km + S) + sigma * rs.standard_normal(n_obs)
        return cls(S, v, precision, form, km_range)

    @property
    def n_obs(self):
        return self.S.size

    def upload(self, lib, handle):
        if self.form == "sufficient":
            _lib.check(handle, lib.smcb_set_data_mm_rate_sufficient(handle, self.S.ctypes.data, self.v.ctypes.data,
                                                                    self.S.size, self.km_range[0], self.km_range[1]))
        else:
            _lib.check(handle, lib.smcb_set_data_mm_rate(handle, self.S.ctypes.data, self.v.ctypes.data,
                                                         self.S.size, self.precision))


class KineticRK:
    """Methanation-style reactor, fixed-step RK4 (physics: methanation_set_likelihood.py:44-66,
    204-208,289-298; reactor definition: DESIGN.md)."""
    model_id = _lib.MODEL_KINETIC_RK

    def __init__(self, cond, obs, base, est_pos, n_steps=50, names=None):
        self.cond, self.obs, self.base = _f64(cond), _f64(obs), _f64(base)
        self.est_pos = np.ascontiguousarray(est_pos, dtype=np.int32)
        self.n_steps = int(n_steps)
        if self.cond.ndim != 2 or self.cond.shape[1] != _lib.KIN_NCOND_FIELDS:
            raise ValueError(f"cond must be [n_cond, {_lib.KIN_NCOND_FIELDS}]")
        if self.obs.shape != (5, self.cond.shape[0]):
            raise ValueError("obs must be [5, n_cond]")
        if (self.base.size - 1) % 8 != 0:
            raise ValueError("base must hold 8*M kinetic parameters followed by sigma")
        self.n_pairs = (self.base.size - 1) // 2
        self.d = int(self.est_pos.size)
        self.names = tuple(names) if names is not None else tuple(f"p{i}" for i in self.est_pos)

    @property
    def n_obs(self):
        return self.obs.size

    def upload(self, lib, handle):
        _lib.check(handle, lib.smcb_set_data_kinetic(
            handle, self.cond.ctypes.data, self.obs.ctypes.data, self.cond.shape[0], self.base.ctypes.data,
            self.n_pairs, self.est_pos.ctypes.data, self.d, self.n_steps))


class KineticDAE(KineticRK):
    """The reference's transient fixed-bed reactor (`reaction`, methanation_set_likelihood.py:69-139: 7 unknowns on 51
    axial nodes, start-up from a 400 K bed to 75 s) instead of the plug-flow march: like-for-like physics for
    reference-sized particle counts.  Same operating-condition rows, observations and parameter
    vector as `KineticRK` with the reference's 8 kinetic parameters; implicit Euler on a fixed time grid, one thread
    block per (particle, condition) (`csrc/dae.cu`).  Fused sweeps are not available for this model."""
    model_id = _lib.MODEL_KINETIC_DAE

    def __init__(self, cond, obs, base, est_pos, names=None):
        super().__init__(cond, obs, base, est_pos, n_steps=1, names=names)
        if self.n_pairs != 4:
            raise ValueError("the transient reactor model takes the reference's 8 kinetic parameters + sigma")


class UserKernelLikelihood:
    """A likelihood compiled by the user (SMCB_MODEL_USER): a shared library exporting a host function of type
    `smcb_user_loglik_fn` (include/smcb200.h) that enqueues the user's own kernels; `include/smcb_user.cuh` holds the
    few lines such a kernel needs and `examples/user_gauss.cu` is a complete one.  Counterpart of writing a new
    `sim_particle` for the reference (`SMC_example/Micmem_likelihood.py:79-92`,
    `SMC_methanation/methanation_functions.py:70-92`)."""
    model_id = _lib.MODEL_USER

    def __init__(self, library, symbol, d, n_obs=0, user_data=None, names=None):
        import ctypes as C
        self.dll = C.CDLL(library) if isinstance(library, (str, bytes, os.PathLike)) else library
        self._fn = C.cast(getattr(self.dll, symbol), C.c_void_p)
        self._user_data = user_data
        self.d, self.n_obs = int(d), int(n_obs)
        self.names = tuple(names) if names is not None else tuple(f"p{i}" for i in range(self.d))

    def upload(self, lib, handle):
        _lib.check(handle, lib.smcb_set_user_likelihood(handle, self._fn, self._user_data))


class _DevView:
    """Zero-copy torch view of device memory the library owns (through __cuda_array_interface__)."""

    def __init__(self, ptr, shape, typestr, strides=None):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": strides}


class CallableLikelihood:
    """Any Python callable on device tensors as the likelihood (SMCB_MODEL_USER through a ctypes trampoline):

        fn(theta, active, lk_out)      theta  [d, n] float64 CUDA tensor (SoA view, row stride >= n)
                                       active [n] uint8 CUDA tensor or None: particles to evaluate
                                       lk_out [n] float64 CUDA tensor

    `fn` either writes lk_out in place (entries of inactive particles are ignored and restored) or returns a [n]
    tensor.  It runs on the sampler's CUDA stream (torch's current stream inside the call) and must not synchronise.
    The counterpart of handing the reference a new `sim_particle` (`SMC_example/Micmem_likelihood.py:79-92`)."""
    model_id = _lib.MODEL_USER

    def __init__(self, fn, d, n_obs=0, names=None):
        self.fn, self.d, self.n_obs = fn, int(d), int(n_obs)
        self.names = tuple(names) if names is not None else tuple(f"p{i}" for i in range(self.d))
        self.error = None
        self._cb = _lib.USER_LOGLIK_FN(self._trampoline)     # keeps the C thunk alive as long as this object

    def _trampoline(self, user, theta_p, ld, n, d, active_p, lk_p, stream):
        import contextlib
        import torch
        try:
            dev = torch.device("cuda", torch.cuda.current_device())
            theta = torch.as_tensor(_DevView(theta_p, (d, n), "<f8", (ld * 8, 8)), device=dev)
            lk = torch.as_tensor(_DevView(lk_p, (n,), "<f8"), device=dev)
            active = torch.as_tensor(_DevView(active_p, (n,), "|u1"), device=dev) if active_p else None
            cur = torch.cuda.current_stream(dev)
            ctx = contextlib.nullcontext() if (stream or 0) == cur.cuda_stream else \
                torch.cuda.stream(torch.cuda.ExternalStream(stream or 0, device=dev))
            with ctx:
                keep = lk.clone() if active is not None else None
                out = self.fn(theta, active, lk)
                if out is not None and out.data_ptr() != lk.data_ptr():
                    lk.copy_(out)
                if active is not None:                     # masked particles keep their old value
                    torch.where(active.bool(), lk, keep, out=lk)
            return 0
        except Exception as e:                             # an exception must not unwind through the C frames
            self.error = e
            return 1

    def upload(self, lib, handle):
        import ctypes as C
        _lib.check(handle, lib.smcb_set_user_likelihood(handle, C.cast(self._cb, C.c_void_p), None))
