#!/usr/bin/env python
"""The reference's Michaelis-Menten driver (`SMC_example/Micmem_SMC_main.py`) written against this engine.

Run it from the reference's `SMC_example/` directory (it reads `data/mm_pseudo_data_{0..5}.csv` exactly as
`Micmem_settings.py:103-115` does), or pass the directory:

    python examples/mm_main.py /path/to/SMC_example [n_particle]

Two shapes are shown:
  (A) the one-call surface: likelihood + prior + settings in, posterior particles / beta schedule /
      log-evidence out;
  (B) the reference's own script shape - `sim_particle`, tempering, resampling and the MH sweeps called
      stage by stage with the reference's variable names - for users who keep their driver.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200  # noqa: E402

data_dir = sys.argv[1] if len(sys.argv) > 1 else "."
n_particle = int(sys.argv[2]) if len(sys.argv) > 2 else 1000

# ---- Micmem_settings.py ---------------------------------------------------------------------------
priors = {"Vmax": {"dist": "uniform", "low": 0.0, "high": 10.0},
          "Km": {"dist": "uniform", "low": 0.0, "high": 10.0},
          "sigma": {"dist": "uniform", "low": 0.0, "high": 10.0}}
settings = smcb200.Settings(n_particle=n_particle)          # same names and defaults as Micmem_settings.py:15-31
prior = smcb200.UniformBox.from_priors(priors)
if os.path.exists(os.path.join(data_dir, "data", "mm_pseudo_data_0.csv")):
    likelihood = smcb200.MMProgress.from_csv(os.path.join(data_dir, "data", "mm_pseudo_data"), n_ex=6)
else:   # the same six curves as recorded in this repository's golden fixture
    g = np.load(os.path.join(ROOT, "tests", "golden", "mm_reference_run.npz"))
    likelihood = smcb200.MMProgress(g["data_t"], g["data_P"], g["data_S0"])
np.random.seed(20250205)                                     # Micmem_settings.py:47
p_pred = np.stack([np.random.uniform(c["low"], c["high"], n_particle) for c in priors.values()], axis=1)

# ---- (A) one call ----------------------------------------------------------------------------------
res = smcb200.run(likelihood, prior, particles=p_pred, settings=settings)
print(f"(A) {len(res.betas)} stages, {sum(res.n_mh)} MH sweeps, {res.seconds * 1e3:.1f} ms on the device, "
      f"log-evidence {res.log_evidence:.4f}")
print("    posterior mean", res.particles.mean(0), "std", res.particles.std(0))

# ---- (B) the reference's loop, stage by stage -------------------------------------------------------
eng = smcb200.Engine(likelihood, prior, settings)
lk = eng.sim_particle(p_pred)                                # lk, _ = sim_particle(p_pred)        main:98
gamma_old = 0.0
rng = np.random.RandomState(1)
for step in range(1, settings.itr_max):                      # main:109
    t = eng.temper(gamma_old)                                # back-off on ESS                     main:111-144
    gamma_new = t["gamma_new"]
    eng.resample(t["gm"], rng.rand())                        # residual-systematic resampling      main:147-184
    eng.moved.zero_()
    eng.icnt[:4].zero_()
    nMH, r_th = (settings.ad_mhstep_num, settings.r_threshold_f) if gamma_new >= 1.0 else \
                (settings.mhstep_num, settings.r_threshold)  # main:193-208
    mhstep_ratio = 1.0
    for j in range(nMH):                                     # main:209
        F, cov = eng.proposal_factor()                       # np.cov(...)*w_cov, SVD factor       main:212-215
        eng.mh_sweep(gamma_new, F, mhstep_ratio, step, j)    # propose, box prior, likelihood, accept  main:220-241
        moved = int(eng.icnt[1].item())
        if moved > r_th * n_particle:                        # main:243
            break
        if moved < settings.r_threshold_min * n_particle:    # main:247
            mhstep_ratio *= 0.5
    print(f"(B) step {step:2d} nMH {j + 1:2d} ESS {t['ess']:.4f} max lk {t['max_lk']:.3f} gamma {gamma_new:.6f} moved {moved}")
    if gamma_new == 1.0:
        break
    gamma_old = gamma_new
post = eng.particles().cpu().numpy()
print("    posterior mean", post.mean(0))
eng.close()
