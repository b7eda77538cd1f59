"""Small runs through every kernel family for `compute-sanitizer --tool memcheck` (one tool per gpurun call):

    python profiles/sanitizer_run.py && compute-sanitizer --tool memcheck python profiles/sanitizer_run.py

Round 2 covers what round 1's advisor found by reading (the d = 32 moment scratch, the fused-sweep lists after a
growing reserve) and the kernels that are new this round: single-pass resampling (ragged sizes, several tiles),
merged moments + device factor (d = 3, 5, 32), sufficient-statistic and closed-form likelihoods, a user kernel."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smcb200 as pkg  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "mm_reference_run.npz"))
kf = np.load(os.path.join(ROOT, "tests", "golden", "kinetic_synth.npz"))
box3 = pkg.UniformBox([0, 0, 0], [10, 10, 10])


def run(tag, lik, prior, n, **kw):
    eng = pkg.Engine(lik, prior, pkg.Settings(n_particle=n, **kw))
    eng.sample_prior()
    res = eng.run()
    print(tag, n, len(res.betas), f"{res.log_evidence:.6f}", res.n_eval, res.n_eval_cut, flush=True)
    eng.close()


d = (g["data_t"], g["data_P"], g["data_S0"])
for n in (1, 37, 1000, 4099, 20481):                      # 20481: eleven resampling tiles, the last one ragged
    run("mm_progress", pkg.MMProgress(*d), box3, n, mm_budget=16)
run("mm_progress_exact", pkg.MMProgress(*d, integrator="exact"), box3, 3001)
for n in (1, 5, 1023, 4100):
    run("mm_rate32", pkg.MMRate.synthetic(333, precision=32), box3, n)
run("mm_rate64", pkg.MMRate.synthetic(333, precision=64), box3, 777)
run("mm_rate_suff", pkg.MMRate.synthetic(333, form="sufficient"), box3, 5000)
run("mm_rate_host_factor", pkg.MMRate.synthetic(100), box3, 900, factor="host")
run("kinetic d=5", pkg.KineticRK(kf["cond"], kf["obs4"], kf["base4"], kf["est4"], n_steps=10),
    pkg.UniformBox(kf["low4"], kf["high4"]), 513)
# fused sweeps: a small run first, then a larger one on the same (pooled) handle - the case of ADVICE r1 (medium)
base = kf["base16"]
lo, hi = np.minimum(base[:32] * 0.8, base[:32] * 1.2), np.maximum(base[:32] * 0.8, base[:32] * 1.2)
lik32 = pkg.KineticRK(kf["cond"][:4], kf["obs16"][:, :4], base, np.arange(32, dtype=np.int32), n_steps=5)
for n in (300, 70001):                                     # 70001 x d = 32: the moment scratch of ADVICE r1 (high)
    run("kinetic32 fused", lik32, pkg.UniformBox(lo, hi), n, fused_sweeps=3, mhstep_num=3, ad_mhstep_num=3,
        early_exit=False)
dg = np.load(os.path.join(ROOT, "tests", "golden", "dae_synth.npz"))
run("kinetic_dae", pkg.KineticDAE(dg["cond"][:2], dg["obs"][:, :2], dg["base4"], dg["est4"]),
    pkg.UniformBox(dg["base4"][dg["est4"]] * 0.9, dg["base4"][dg["est4"]] * 1.1), 24, mhstep_num=1, ad_mhstep_num=2)
so = pkg.build_user_library(os.path.join(ROOT, "examples", "user_gauss.cu"))
dll = C.CDLL(so)
x = np.linspace(0.0, 4.0, 60)
y = 1.5 + 0.7 * x + 0.3 * np.random.RandomState(4).standard_normal(60)
assert dll.user_gauss_set_data(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), 60) == 0
run("user kernel", pkg.UserKernelLikelihood(dll, "user_gauss_loglik", d=3, n_obs=60),
    pkg.UniformBox([-5, -5, 0], [5, 5, 5]), 2500)
print("done")
