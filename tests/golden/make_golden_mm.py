#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference Michaelis-Menten
SMC script in this container.

The reference (`/root/reference/SMC_example/Micmem_SMC_main.py`) cannot be
imported as-is because it imports `ray`, `assimulo`, `matplotlib`, `seaborn`
and `memory_profiler` at module top, none of which is installed here.  None of
those packages does arithmetic on the sampler path: `ray` only fans the
per-particle likelihood out to worker processes, the others are unused or
plotting.  This script therefore installs *stub modules* for them
(`ray.remote`/`ray.get` evaluate the very same reference function in a local
fork pool) and then executes the reference sources where they lie with
`runpy`.  Every number stored in the fixture is produced by reference code +
numpy/scipy; nothing from this repository is involved.

What is recorded (-> tests/golden/mm_reference_run.npz):
  * the data set the reference loaded (six CSVs: t, P_obs, S0),
  * the prior particles drawn at import (`Micmem_settings.py:69-87`),
  * every `sim_particle` sweep: the particle matrix that went in and the
    log-likelihood vector that came out (`Micmem_likelihood.py:79-92`),
  * every draw the driver took from the global NumPy stream, in order:
    `rand()` (`Micmem_SMC_main.py:156`), `multivariate_normal` (`:220`),
    `uniform(0,1,N)` (`:235`),
  * the per-stage line the reference prints (`:254`): nMH index, ESS, max
    log-likelihood, gamma, moved count,
  * final particles / log-likelihoods.

Run:  python tests/golden/make_golden_mm.py   (about 1-5 min on 8 cores)
"""
import io
import os
import re
import runpy
import sys
import types
import contextlib
import multiprocessing as mp

import numpy as np

REF_DIR = "/root/reference/SMC_example"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mm_reference_run.npz")

# ----------------------------------------------------------------------------
# stub modules
# ----------------------------------------------------------------------------
_REGISTRY = []          # remote functions, index = id
_POOL = None
_TRACE = {"sweeps_in": [], "sweeps_out": [], "pmodel0": None}


def _call(task):
    fid, args = task
    return _REGISTRY[fid](*args)


class _Remote:
    def __init__(self, fn):
        self.fn = fn
        _REGISTRY.append(fn)
        self.fid = len(_REGISTRY) - 1

    def remote(self, *args):
        return (self.fid, args)

    def __call__(self, *a, **k):  # direct call (the reference's broken smoke block)
        return self.fn(*a, **k)


def _ray_remote(*a, **k):
    if len(a) == 1 and callable(a[0]) and not k:
        return _Remote(a[0])
    return lambda fn: _Remote(fn)


def _ray_get(tasks):
    global _POOL
    if _POOL is None:
        _POOL = mp.get_context("fork").Pool(os.cpu_count())
    res = _POOL.map(_call, tasks, chunksize=16)
    _TRACE["sweeps_in"].append(np.array([t[1][0] for t in tasks], dtype=np.float64))
    _TRACE["sweeps_out"].append(np.array([r[0] for r in res], dtype=np.float64))
    if _TRACE["pmodel0"] is None:
        # model predictions of the first 8 prior particles (6 x 40 each)
        _TRACE["pmodel0"] = np.array([np.array(r[1]) for r in res[:8]])
    return res


def _mk(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Anything:
    """Absorbs any attribute access / call (plotting stubs)."""
    def __getattr__(self, k):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def install_stubs():
    _mk("ray", remote=_ray_remote, get=_ray_get, init=lambda *a, **k: None,
        shutdown=lambda *a, **k: None)
    _mk("memory_profiler", profile=lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f)))
    asm = _mk("assimulo")
    asm.solvers = _mk("assimulo.solvers", Radau5DAE=_Anything(), IDA=_Anything())
    asm.problem = _mk("assimulo.problem", Implicit_Problem=_Anything())
    mpl = _mk("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = _mk("matplotlib.pyplot")
    mpl.pyplot.__getattr__ = lambda k: _Anything()
    _mk("seaborn").__getattr__ = lambda k: _Anything()
    _mk("pylab").__getattr__ = lambda k: _Anything()


def main():
    install_stubs()
    os.chdir(REF_DIR)               # the reference opens data/... relative to cwd
    sys.path.insert(0, REF_DIR)

    draws = {"rand": [], "mvn": [], "mvn_cov": [], "unif": []}
    import numpy.random as npr
    _rand, _mvn, _unif = npr.rand, npr.multivariate_normal, npr.uniform

    def rand(*a):
        v = _rand(*a)
        draws["rand"].append(v)
        return v

    def mvn(mean, cov, size=None, **k):
        v = _mvn(mean, cov, size, **k)
        draws["mvn"].append(np.array(v))
        draws["mvn_cov"].append(np.array(cov))
        return v

    in_settings = {"on": True}

    def unif(low=0.0, high=1.0, size=None):
        v = _unif(low, high, size)
        if not in_settings["on"]:
            draws["unif"].append(np.array(v))
        return v

    npr.rand, npr.multivariate_normal, npr.uniform = rand, mvn, unif
    np.random.rand, np.random.multivariate_normal, np.random.uniform = rand, mvn, unif

    import Micmem_settings as S       # seeds, samples the prior, loads the CSVs
    in_settings["on"] = False
    prior_particles = S.p_pred.copy()
    data_t = np.array([d["t"] for d in S.dataset])
    data_P = np.array([d["P_obs"] for d in S.dataset])
    data_S0 = np.array([d["S0"] for d in S.dataset])

    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        g = runpy.run_path(os.path.join(REF_DIR, "Micmem_SMC_main.py"), run_name="ref_main")
    text = buf.getvalue()
    sys.stderr.write(text[-2000:])

    pat = re.compile(r"iteration:(\d+), nMH:(\d+), Calculation time:[^,]+, ESS:([^,]+), "
                     r"Max Likelihood:([^,]+), New Gamma:([^,]+), Number of Adoption:([^\s]+)")
    rows = [(int(m[1]), int(m[2]), float(m[3]), float(m[4]), float(m[5]), float(m[6]))
            for m in pat.finditer(text)]
    stage = np.array(rows, dtype=np.float64)
    assert len(rows) > 0 and rows[-1][4] == 1.0, "reference did not reach gamma=1"

    # known answers of the reference likelihood at two parameter vectors
    import Micmem_likelihood as L
    ka_in = np.array([[1.2, 0.5, 0.02], [1.0, 0.4, 0.05]])
    ka_out = np.array([L.log_likelihood_mm_multi.fn(p)[0] for p in ka_in])

    np.savez_compressed(
        OUT,
        data_t=data_t, data_P=data_P, data_S0=data_S0,
        prior_particles=prior_particles,
        sweeps_in=np.array(_TRACE["sweeps_in"]), sweeps_out=np.array(_TRACE["sweeps_out"]),
        pmodel0=_TRACE["pmodel0"],
        draws_rand=np.array(draws["rand"]), draws_mvn=np.array(draws["mvn"]),
        draws_mvn_cov=np.array(draws["mvn_cov"]), draws_unif=np.array(draws["unif"]),
        stage_table=stage,  # columns: step, nMH index j, ESS, max lk, gamma_new, moved
        final_particles=np.array(g["p_pred"]), final_lk=np.array(g["lk"], dtype=np.float64),
        ka_in=ka_in, ka_out=ka_out,
        seed=np.array(20250205),
        versions=np.array([np.__version__, __import__("scipy").__version__]),
    )
    print("wrote", OUT, "stages:", len(rows), "sweeps:", len(_TRACE["sweeps_in"]), file=sys.stderr)


if __name__ == "__main__":
    main()
