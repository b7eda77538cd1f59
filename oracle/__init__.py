"""CPU oracle for the likelihood-tempered SMC hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU and in plain NumPy / SciPy / C, the
arithmetic of the reference sampler loop
(`/root/reference/SMC_example/Micmem_SMC_main.py:105-262` and its twins in
`SMC_methanation/`).  It exists so that the CUDA path can be checked; it is
never the thing that is shipped or measured.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.  The product package never does, and raises if its CUDA
library is missing.

Parity status
-------------
* Michaelis-Menten path (likelihood + tempering + resampling + MH): PINNED.
  `tests/golden/mm_reference_run.npz` was produced by executing the unmodified
  reference sources in the build container (`tests/golden/make_golden_mm.py`);
  `tests/test_oracle_golden.py` replays the oracle against every sweep, every
  stage and the final particles of that run.
* The MM likelihood itself lives in third-party `scipy.integrate.solve_ivp`
  (RK45; scipy is un-pinned by the reference, 1.18.1 here).  `oracle.mm` can
  call scipy directly (that *is* the reference arithmetic) and also carries an
  operation-for-operation scalar twin (`oracle.dopri5`, and `oracle/c/`) that
  the device kernel mirrors.
* Methanation forward model: PARITY UNPINNED.  `assimulo`/SUNDIALS IDA and the
  operating-conditions file `methanation_data/information.csv` are absent, so
  the reference's DAE cannot be run.  `oracle.kinetic` restates the rate law,
  gas density, outlet-flow and log-likelihood formulas exactly
  (`methanation_set_likelihood.py:44-66,204-208,289-298`) and integrates a
  builder-defined steady plug-flow reactor with fixed-step RK4.
"""
