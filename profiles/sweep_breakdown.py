"""Per-sweep device time of the bulk and the tail kernel of one 2^20-particle run (SMCB_PARAM_PROFILE read after every
sweep), with the sweep's deferred-solve count, longest solve and early rejections.

profiles/sweeps_r02_baseline.log is its output for the shipped kernels.  profiles/sweeps_r02_direct_tail_experiment_
ratio64.log is the same run with an experiment of round 2 that was NOT kept: the solves of particles with Vmax/Km >= 64
were started in the tail kernel on a second stream next to the bulk kernel instead of after it.  It lost by 2.6x:
those solves then run without what the finished experiments of their particle say about it (the particle-level bound
of mm_finalize_kernel), so proposals that are rejected after ~1e3 attempts today ran for up to 6e4."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200 as pkg  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "mm_reference_run.npz"))
lik = pkg.MMProgress(g["data_t"], g["data_P"], g["data_S0"])
prior = pkg.UniformBox([0, 0, 0], [10, 10, 10])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
budget = int(sys.argv[2]) if len(sys.argv) > 2 else 0         # attempted steps before a solve moves to the tail kernel
eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=n, mm_budget=budget))
eng.sample_prior()
eng.run()                     # warm
eng.sample_prior()
rows = []
orig = eng.loglik_into


def timed(theta, lk_out, active=None, lkmin=None):
    eng.kernel_profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(theta, lk_out, active=active, lkmin=lkmin)
    e1.record()
    e1.synchronize()
    b, t, _ = eng.kernel_profile()
    st = eng.loglik_stats()
    rows.append((e0.elapsed_time(e1), b, t, int(st[11]), int(st[10]), int(st[8])))


eng.loglik_into = timed
res = eng.run()
print(f"n={n} stages={len(res.betas)} sweeps={len(rows)} device {res.seconds * 1e3:.1f} ms (with per-sweep syncs)")
print("sweep  group_ms  bulk_ms  tail_ms  other_ms  deferred_solves  longest_attempts  cut_particles")
for i, (gms, b, t, nd, mx, cut) in enumerate(rows):
    print(f"{i:5d} {gms:9.3f} {b:8.3f} {t:8.3f} {gms - b - t:9.3f} {nd:16d} {mx:17d} {cut:14d}")
a = np.array([r[:3] for r in rows])
print("sum", a.sum(0), "hideable min(bulk, tail) =", np.minimum(a[:, 1], a[:, 2]).sum())
