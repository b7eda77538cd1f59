#!/usr/bin/env python
"""Transient reactor model (KINETIC_DAE): sweep timing and a reference-sized tempered run.

    python profiles/prof_dae.py [n_particle=1000] [run=1]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200 as pkg  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
do_run = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = np.load(os.path.join(ROOT, "tests", "golden", "dae_synth.npz"))
kf = np.load(os.path.join(ROOT, "tests", "golden", "kinetic_synth.npz"))
lik = pkg.KineticDAE(g["cond"], g["obs"], g["base4"], g["est4"])
prior = pkg.UniformBox(kf["low4"], kf["high4"], names=["Af", "Eaf", "Ar", "Ear", "sigma"])
eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=N))
eng.sample_prior()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    e0.record(); eng.sim_particle(); e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    lk = eng.lk.cpu().numpy()
    print(f"prior sweep N={N} x 30 conditions: {ms:.1f} ms, {N * 30 / ms * 1e3:.0f} marches/s, "
          f"{N / ms * 1e3:.0f} evals/s; lk finite {np.isfinite(lk).mean():.3f}, median {np.median(lk):.1f}, "
          f"particles with a failed march {np.mean(lk < -1e6):.3f}", flush=True)
if do_run:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = eng.run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"run to beta=1: {dt:.2f} s, {len(res.betas)} stages, {sum(res.n_mh)} sweeps, n_eval {res.n_eval}, "
          f"log-evidence {res.log_evidence:.3f}")
    print("posterior mean", res.particles.mean(0))
    print("truth         ", g["base4"][g["est4"]])
