"""Time of smcb_resample_counts in SEQUENTIAL mode (the reference's sequentially rounded sum, bit-exact) at 2^20
particles, first-stage weights of a prior cloud."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200 as pkg  # noqa: E402

for logn in (16, 20, 22):
    n = 1 << logn
    eng = pkg.Engine(pkg.MMRate.synthetic(200), pkg.UniformBox([0, 0, 0], [10, 10, 10]),
                     pkg.Settings(n_particle=n, scan_mode="sequential"))
    eng.sample_prior()
    eng.sim_particle()
    t = eng.temper(0.0)
    eng._ck(eng.lib.smcb_weights(eng.h, eng.lk.data_ptr(), n, eng.scal.data_ptr(), t["gm"], eng.scal[1:].data_ptr(),
                                 eng.w.data_ptr(), eng._stream))
    for mode, name in ((0, "sequential"), (1, "fixed")):
        ms = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng._ck(eng.lib.smcb_resample_counts(eng.h, eng.w.data_ptr(), n, n, 0.375, mode, None, 0, 0,
                                                 eng.counts.data_ptr(), eng.icnt[4:6].data_ptr(), eng._stream))
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        print(f"n=2^{logn} smcb_resample_counts {name}: {np.median(ms) * 1e3:.1f} us")
    eng.close()
