import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import smcb200 as pkg
g = np.load('/root/repo/tests/golden/mm_reference_run.npz')
lik = pkg.MMProgress(g["data_t"], g["data_P"], g["data_S0"]); prior = pkg.UniformBox([0,0,0],[10,10,10])
worst = np.array([9.88869508e+00, 4.28267643e-04, 4.53281765e+00])
for n, fill in ((1, None), (32, "same"), (32, "light"), (1024, "light"), (65536, "prior")):
    eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=n))
    eng.kernel_profile(True)
    th = np.tile(worst, (n, 1))
    if fill == "light":
        th[1:] = [1.2, 0.5, 0.02]
    if fill == "prior":
        th = np.random.RandomState(0).uniform(0, 10, (n, 3)); th[0] = worst
    for rep in range(2):
        eng.sim_particle(th); torch.cuda.synchronize()
        b, t, _ = eng.kernel_profile()
    st = eng.loglik_stats()
    print(f"n={n:6d} fill={fill}: bulk {b:.2f} ms tail {t:.2f} ms; deferred {st[11]}; longest {int(st[16])>>32} x {int(st[16]) & 0xffffffff} cycles")
    eng.close()
