#!/usr/bin/env python
"""Host-clock trace of a kinetic32 run (config 5 shard): seconds at every stage boundary.
    python profiles/trace_kinetic32.py [log2_particles=21]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200 as pkg  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 21
N = 1 << lg
kf = np.load(os.path.join(ROOT, "tests", "golden", "kinetic_synth.npz"))
b = kf["base16"]
lik = pkg.KineticRK(kf["cond"], kf["obs16"], b, np.arange(32, dtype=np.int32), n_steps=50)
prior = pkg.UniformBox(np.minimum(b[:32] * 0.8, b[:32] * 1.2), np.maximum(b[:32] * 0.8, b[:32] * 1.2))
eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=N, fused_sweeps=10, mhstep_num=10, ad_mhstep_num=10,
                                          early_exit=False))
eng.sample_prior()
p0 = eng.state[: eng.d].clone()
for rep in range(3):
    eng.state[: eng.d].copy_(p0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    marks = []

    def hook(kind, **kw):
        if kind == "stage":
            torch.cuda.synchronize()
            marks.append((kw["step"], kw["gamma"], time.perf_counter() - t0))

    res = eng.run(hook=hook)
    torch.cuda.synchronize()
    print(f"rep {rep}: total {time.perf_counter() - t0:.4f} s (device {res.seconds:.4f} s), n_eval {res.n_eval}")
    for m in marks:
        print(f"   stage {m[0]} gamma {m[1]:.5f} at {m[2] * 1e3:.1f} ms")
