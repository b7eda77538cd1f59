#!/usr/bin/env python
"""Likelihood-kernel micro-run for profiling (ncu) and for quick timing while tuning K1.

    python profiles/prof_loglik.py [log2_particles=20] [reps=3] [model=mm_progress|mm_rate32|mm_rate64|kinetic|kinetic32] [budget] [refill_min] [patience]

Evaluates (a) a prior cloud (Philox uniform box) and (b) a posterior-like cloud (for MM: a Gaussian
around the reference posterior) and prints the CUDA-event time of each sweep and the device work
counters.  Inputs are > L2-resident-irrelevant: the kernel is compute bound (32 B per particle).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200 as pkg  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model = sys.argv[3] if len(sys.argv) > 3 else "mm_progress"
N = 1 << lg
g = np.load(os.path.join(ROOT, "tests", "golden", "mm_reference_run.npz"))
kf = np.load(os.path.join(ROOT, "tests", "golden", "kinetic_synth.npz"))
if model == "mm_progress":
    lik = pkg.MMProgress(g["data_t"], g["data_P"], g["data_S0"])
    prior = pkg.UniformBox([0, 0, 0], [10, 10, 10])
elif model.startswith("mm_rate"):
    lik = pkg.MMRate.synthetic(10000, precision=int(model[-2:]))
    prior = pkg.UniformBox([0, 0, 0], [10, 10, 10])
elif model == "kinetic":
    lik = pkg.KineticRK(kf["cond"], kf["obs4"], kf["base4"], kf["est4"], n_steps=50)
    prior = pkg.UniformBox(kf["low4"], kf["high4"])
else:
    b = kf["base16"]
    lik = pkg.KineticRK(kf["cond"], kf["obs16"], b, np.arange(32, dtype=np.int32), n_steps=50)
    prior = pkg.UniformBox(np.minimum(b[:32] * 0.8, b[:32] * 1.2), np.maximum(b[:32] * 0.8, b[:32] * 1.2))
eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=N))
if len(sys.argv) > 4:   # deferral budget of the bulk MM_PROGRESS kernel
    eng._ck(eng.lib.smcb_set_param(eng.h, 1, float(sys.argv[4])))
if len(sys.argv) > 5:   # lanes a bulk warp waits for before refilling
    eng._ck(eng.lib.smcb_set_param(eng.h, 2, float(sys.argv[5])))
if len(sys.argv) > 6:   # ... for how many attempted steps
    eng._ck(eng.lib.smcb_set_param(eng.h, 3, float(sys.argv[6])))


def sweep(tag):
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.sim_particle()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        extra = ""
        if model == "mm_progress":
            st = eng.loglik_stats()
            extra = (f" rhs/particle={st[0] / N:.1f} acc={st[1] / N:.1f} rej={st[2] / N:.1f} fail={st[3]}"
                     f" max_attempts={st[10]} deferred_solves={st[11]} tail_particles={st[13]}"
                     f" longest_tail_solve={int(st[16]) >> 32}x{int(st[16]) & 0xffffffff}cyc")
        print(f"{model} {tag} N=2^{lg} rep{r}: {ms:.3f} ms  {N / ms * 1e3:.4g} evals/s{extra}", flush=True)


eng.sample_prior()
sweep("prior")
if model.startswith("mm"):
    rs = np.random.RandomState(0)
    fp = g["final_particles"]
    mu, L = fp.mean(0), np.linalg.cholesky(np.cov(fp.T))
    th = mu + rs.standard_normal((N, 3)) @ L.T
    eng.set_particles(th)
    sweep("posterior")
print("lk checksum", float(eng.lk.sum().item()))
