"""Timing of the single-pass resampling kernel on first-stage weights of a prior cloud at 2^20, 2^22 and 2^23 particles:
median of 7 launches, L2 flushed.  profiles/resample_variants_r02.log is its output from the build that still had an
experiment switch (variant 0: unconstrained registers, 4 blocks per SM; 1: __launch_bounds__(256, 6); 2: 6 blocks +
unrolled rows; 3: unrolled rows; 4: 5 blocks): the shipped kernel is variant 1 above 2^21 particles, variant 0 below."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200 as pkg  # noqa: E402

flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for logn in (20, 22, 23):
    n = 1 << logn
    lik = pkg.MMRate.synthetic(10000, form="sufficient")
    eng = pkg.Engine(lik, pkg.UniformBox([0, 0, 0], [10, 10, 10]), pkg.Settings(n_particle=n))
    eng.sample_prior()
    eng.sim_particle()
    t = eng.temper(0.0)
    D1 = eng.d + 1
    nbytes = n * (8 + 4 + 4 + 2 * D1 * 8)
    for var in (0,):
        ms = []
        for _ in range(7):
            flush_buf.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng._ck(eng.lib.smcb_resample_fused(eng.h, eng.lk.data_ptr(), None, n, n, 0, 0, n, eng.scal.data_ptr(), t["gm"],
                                                eng.scal[1:].data_ptr(), 0.375, eng.state.data_ptr(), n, D1,
                                                eng.state2.data_ptr(), n, eng.anc.data_ptr(), None,
                                                eng.icnt[6:7].data_ptr(), eng._stream))
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        m = float(np.median(ms))
        print(f"n=2^{logn} variant {var}: {m * 1e3:8.1f} us  {nbytes / m / 1e6:7.0f} GB/s  ({nbytes / m / 1e6 / 6555.2:.3f} of HBM peak)")
    eng.close()
