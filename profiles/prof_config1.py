"""Where the reference-sized run (BASELINE config 1: N = 1000, 34 sweeps) spends its host time: cProfile of ten runs."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200 as pkg  # noqa: E402
from importlib import import_module  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "mm_reference_run.npz"))
lik = pkg.MMProgress(g["data_t"], g["data_P"], g["data_S0"])
prior = pkg.UniformBox([0, 0, 0], [10, 10, 10])
Stream = import_module(pkg.__name__ + ".reference_api").LegacyNumpyStream
mode = sys.argv[1] if len(sys.argv) > 1 else "parity"
eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=1000, scan_mode="sequential" if mode == "parity" else "fixed"))


def once():
    if mode == "parity":
        st = Stream(int(g["seed"]))
        p0 = st.prior_uniform([0, 0, 0], [10, 10, 10], 1000)
        return eng.run(p0, stream=st)
    eng.sample_prior()
    return eng.run()


for _ in range(3):
    r = once()
t0 = time.perf_counter()
for _ in range(10):
    r = once()
print(f"{mode}: wall per run {(time.perf_counter() - t0) * 100:.2f} ms, device {r.seconds * 1e3:.2f} ms, launches per run "
      f"{eng.launch_count() / 13:.0f}, sweeps {sum(r.n_mh) + 1}")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    once()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
