"""The tail kernel alone, for ncu: 32 copies of the stiffest particle of the bench's 2^20-particle prior cloud (its six
solves take up to 83 133 attempted steps), so that mm_tail_kernel's duration is that of one strictly serial chain and
the per-instruction stall samples of the source page show where a lone lane waits.

    ncu --set full --import-source on -k regex:mm_tail_kernel -s 2 -c 1 -o gpurun_out/prof_tail python profiles/prof_tail.py
"""
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=n))
eng.kernel_profile(True)
th = np.tile(np.array([9.88869508e+00, 4.28267643e-04, 4.53281765e+00]), (n, 1))
for rep in range(2):
    lk = eng.sim_particle(th)
    torch.cuda.synchronize()
    b, t, _ = eng.kernel_profile()
    st = eng.loglik_stats()
    print(f"rep {rep}: bulk {b:.3f} ms, tail {t:.3f} ms, deferred solves {int(st[11])}, longest {int(st[16]) >> 32} attempts x "
          f"{int(st[16]) & 0xffffffff} cycles, lk[0] = {lk[0][0] if isinstance(lk, tuple) else lk[0]!r}")
eng.close()
