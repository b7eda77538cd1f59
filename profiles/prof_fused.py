#!/usr/bin/env python
"""Fused multi-sweep MH kernel (kinetic32) timing split: proposals only (ratio so large that nothing survives the
box test) against the normal launch.   python profiles/prof_fused.py [log2_particles=21] [sweeps=10]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200 as pkg  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 21
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
N = 1 << lg
kf = np.load(os.path.join(ROOT, "tests", "golden", "kinetic_synth.npz"))
b = kf["base16"]
lik = pkg.KineticRK(kf["cond"], kf["obs16"], b, np.arange(32, dtype=np.int32), n_steps=50)
prior = pkg.UniformBox(np.minimum(b[:32] * 0.8, b[:32] * 1.2), np.maximum(b[:32] * 0.8, b[:32] * 1.2))
eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=N))
eng.sample_prior()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.sim_particle(); e1.record(); e1.synchronize()
print(f"first sweep N=2^{lg}: {e0.elapsed_time(e1):.2f} ms")
F, _ = eng.proposal_factor()
for tag, ratio in (("proposals only", 1e3), ("normal", 1.0), ("normal", 1.0), ("ratio 0.25", 0.25)):
    eng.moved.zero_(); eng.icnt.zero_()
    e0.record(); eng.mh_fused(0.01, F, ratio, 3, 0, k); e1.record(); e1.synchronize()
    c = eng.icnt.cpu().numpy()
    ms = e0.elapsed_time(e1)
    print(f"{tag}: {k} sweeps {ms:.2f} ms, in-box evals {c[2]}, moved {c[1]}, "
          f"{(ms / max(1, c[2])) * 1e3:.3f} us per eval", flush=True)
