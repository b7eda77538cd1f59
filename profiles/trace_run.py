#!/usr/bin/env python
"""Per-sweep trace of one tempered-SMC run (MM progress curves): device time of every likelihood sweep
with the work counters of the MM_PROGRESS kernels (smcb_loglik_stats).

    python profiles/trace_run.py [log2_particles=20] [budget=256] [early_reject=1] [chunk=32]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200 as pkg  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 256.0
er = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
N = 1 << lg
g = np.load(os.path.join(ROOT, "tests", "golden", "mm_reference_run.npz"))
lik = pkg.MMProgress(g["data_t"], g["data_P"], g["data_S0"])
prior = pkg.UniformBox([0, 0, 0], [10, 10, 10])
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 32
tw = int(sys.argv[5]) if len(sys.argv) > 5 else 32
eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=N, early_reject=er, mm_budget=int(budget), mm_chunk=chunk, mm_tail_warps=tw))
orig = eng.loglik_into
rows = []


def traced(theta, lk_out, active=None, lkmin=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(theta, lk_out, active=active, lkmin=lkmin)
    e1.record()
    e1.synchronize()
    st = eng.loglik_stats()
    bulk_ms, tail_ms, _ = eng.kernel_profile()
    n_act = int(active.sum().item()) if active is not None else theta.shape[1]
    rows.append((e0.elapsed_time(e1), n_act, st[0], st[1] + st[2], st[8], st[10], st[11], st[13], bulk_ms, tail_ms))


eng.loglik_into = traced
eng.kernel_profile(True)
for rep in range(2):
    rows.clear()
    eng.sample_prior()
    res = eng.run()
print(f"N=2^{lg} budget={budget:.0f} chunk={chunk} tail_warps={tw} early_reject={er}: {res.seconds * 1e3:.1f} ms to beta=1, {len(res.betas)} stages, "
      f"sweeps {res.n_mh}, logZ {res.log_evidence:.4f}")
print(" sweep      ms  bulk_ms  tail_ms   evaluated  rhs/eval  attempts/eval  cut_particles  max_attempts  deferred_solves  tail_particles")
for i, r in enumerate(rows):
    print(f"{i:6d} {r[0]:7.3f} {r[8]:8.3f} {r[9]:8.3f} {r[1]:10d} {r[2] / max(r[1], 1):9.1f} {r[3] / max(r[1], 1):13.1f} {r[4]:13d} {r[5]:13d} {r[6]:15d} {r[7]:14d}")
print(f"sum: sweeps {sum(r[0] for r in rows):.2f} ms, bulk {sum(r[8] for r in rows):.2f} ms, tail {sum(r[9] for r in rows):.2f} ms")
print("betas", [round(b, 5) for b in res.betas])
