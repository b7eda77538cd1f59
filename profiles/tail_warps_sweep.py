import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import smcb200 as pkg
g = np.load('/root/repo/tests/golden/mm_reference_run.npz')
lik = pkg.MMProgress(g["data_t"], g["data_P"], g["data_S0"]); prior = pkg.UniformBox([0,0,0],[10,10,10])
N = 1 << 20
for tw in (1, 2, 4, 8, 12):
    eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=N, mm_tail_warps=tw))
    eng.kernel_profile(True)
    eng.sample_prior()
    for rep in range(2):
        eng.sim_particle(); torch.cuda.synchronize()
        b, t, _ = eng.kernel_profile()
    st = eng.loglik_stats()
    print(f"tail_warps={tw:2d}: bulk {b:.2f} ms tail {t:.2f} ms; longest solve {int(st[16])>>32} attempts x {int(st[16]) & 0xffffffff} cycles")
    eng.close()
