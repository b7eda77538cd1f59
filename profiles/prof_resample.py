#!/usr/bin/env python
"""Tempering / resampling / moment kernels on one stage of a 2^22-particle cloud (for ncu: HBM traffic of the
bandwidth-bound kernels K2, K3, K4a).

    python profiles/prof_resample.py [log2_particles=22] [reps=3]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200 as pkg  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = 1 << lg
lik = pkg.MMRate.synthetic(64, precision=32)          # a cheap likelihood: the stage kernels are what is measured
prior = pkg.UniformBox([0, 0, 0], [10, 10, 10])
eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=N))
eng.sample_prior()
eng.sim_particle()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
eng.enable_profiling(True)
for r in range(reps):
    flush.fill_(r)                                     # > 126 MB L2: the stage reads its inputs from HBM
    t = eng.temper(0.0)
    eng.resample(t["gm"], 0.37)
    eng.proposal_factor()
    eng.sim_particle()
s = eng.profile_summary()
print(f"N=2^{lg}, {reps} stages: " + ", ".join(f"{k} {v[1] / reps:.3f} ms" for k, v in s.items()))
D1 = eng.d + 1
print(f"gather: {N * (4 + 2 * D1 * 8) / 1e6:.1f} MB algorithmic per launch; weights/temper pass: {N * 8 / 1e6:.1f} MB read")
