#!/usr/bin/env python
"""Where the end-to-end call smcb200.run(...) spends its time (host clock, synchronised)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200 as pkg

N = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
g = np.load(os.path.join(ROOT, "tests", "golden", "mm_reference_run.npz"))
lik = pkg.MMProgress(g["data_t"], g["data_P"], g["data_S0"])
prior = pkg.UniformBox([0, 0, 0], [10, 10, 10])
cfg = pkg.Settings(n_particle=N)
host_p = torch.empty((N, 3), dtype=torch.float64).pin_memory()
host_p.copy_(torch.from_numpy(np.random.RandomState(0).uniform(0, 10, (N, 3))))


def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"  {label:28s} {1e3 * (t1 - t0):8.2f} ms")
    return t1


for it in range(3):
    print("iteration", it)
    t0 = time.perf_counter()
    t = t0
    eng = pkg.Engine(lik, prior, cfg)
    t = tick("Engine()", t)
    eng.set_particles(host_p)
    t = tick("set_particles (H2D)", t)
    res = eng.run()
    t = tick("run (incl. result D2H)", t)
    print(f"    device time inside run       {res.seconds * 1e3:8.2f} ms")
    eng.close()
    t = tick("close", t)
    print(f"  total {1e3 * (t - t0):.2f} ms")
