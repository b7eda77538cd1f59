"""ctypes front of the plain-C oracle twin `oracle/c/mm_dopri5.c` (test infrastructure only).

Same arithmetic as `oracle.dopri5` / `oracle.mm.loglik_progress_twin`, fast enough to follow the device
at 2^16-2^20 particles and to report the per-solve step counts of a particle cloud."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_DIR, "libmm_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_DIR, "mm_dopri5.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _DIR, "-B"], check=True, capture_output=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.mm_progress_loglik.restype = None
        _lib.mm_progress_loglik.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def loglik_progress(theta, data_t, data_P, data_S0, want_steps=False, want_pred=False):
    """theta [n,3] -> lk [n] (+ dict with counters / steps_per_solve [n,n_ex] / pred [n,n_ex,n_t])."""
    lib = _load()
    th = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, 3)
    t, P, S0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (data_t, data_P, data_S0))
    n, (n_ex, n_t) = th.shape[0], t.shape
    lk = np.empty(n)
    counters = np.zeros(4, dtype=np.int64)
    steps = np.zeros((n, n_ex), dtype=np.int32) if want_steps else None
    pred = np.zeros((n, n_ex, n_t)) if want_pred else None
    lib.mm_progress_loglik(th.ctypes.data, n, t.ctypes.data, P.ctypes.data, S0.ctypes.data, n_ex, n_t,
                           lk.ctypes.data, counters.ctypes.data, steps.ctypes.data if want_steps else None,
                           pred.ctypes.data if want_pred else None)
    return lk, dict(nfev=int(counters[0]), accepted=int(counters[1]), rejected=int(counters[2]),
                    failed=int(counters[3]), steps=steps, pred=pred)


def _chunk(args):
    return loglik_progress(*args)[0]


def loglik_progress_parallel(theta, data_t, data_P, data_S0, pool, n_chunks):
    parts = pool.map(_chunk, [(c, data_t, data_P, data_S0) for c in np.array_split(np.asarray(theta), n_chunks) if len(c)])
    return np.concatenate(parts)
