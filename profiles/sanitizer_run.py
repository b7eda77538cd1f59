import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import smcb200 as pkg
g = np.load('/root/repo/tests/golden/mm_reference_run.npz')
lik = pkg.MMProgress(g["data_t"], g["data_P"], g["data_S0"]); prior = pkg.UniformBox([0,0,0],[10,10,10])
for n in (1, 37, 1000, 4099):
    eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=n, mm_budget=16))
    eng.sample_prior()
    res = eng.run()
    print("mm_progress", n, len(res.betas), res.log_evidence, res.n_eval, res.n_eval_cut)
    eng.close()
lik2 = pkg.MMRate.synthetic(333, precision=32)
for n in (1, 5, 1023, 4100):
    eng = pkg.Engine(lik2, prior, pkg.Settings(n_particle=n))
    eng.sample_prior(); res = eng.run(); print("mm_rate32", n, len(res.betas), res.log_evidence); eng.close()
lik3 = pkg.MMRate.synthetic(333, precision=64)
eng = pkg.Engine(lik3, prior, pkg.Settings(n_particle=777)); eng.sample_prior(); res = eng.run(); print("mm_rate64", len(res.betas)); eng.close()
kf = np.load('/root/repo/tests/golden/kinetic_synth.npz')
lik4 = pkg.KineticRK(kf["cond"], kf["obs4"], kf["base4"], kf["est4"], n_steps=10)
eng = pkg.Engine(lik4, pkg.UniformBox(kf["low4"], kf["high4"]), pkg.Settings(n_particle=513)); eng.sample_prior(); res = eng.run(); print("kinetic", len(res.betas)); eng.close()
print("done")
