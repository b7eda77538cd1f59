"""The reference's function-level surface on top of the engine.

`sim_particle(particle) -> (llk, C_l_)` and `cal_prior(theta, priors)` are the two calls the
reference driver makes into its likelihood/prior modules (`SMC_example/Micmem_likelihood.py:79-92`,
`SMC_example/Micmem_SMC_main.py:60-90`).  These shims keep that script shape working while the
work runs on the GPU; `examples/mm_main.py` shows the whole reference driver written against them.
"""
import numpy as np
import torch


class ReferenceSurface:
    def __init__(self, engine):
        self.eng = engine

    def sim_particle(self, particle, with_predictions=False):
        """particle: [N, d] array -> (llk tuple-like ndarray[N], C_l_ or None).

        The reference returns per-particle model predictions C_l_ (used only for plots); they are
        produced on request for the MM progress-curve model."""
        eng = self.eng
        lk = eng.sim_particle(np.asarray(particle, dtype=np.float64))
        llk = lk.cpu().numpy().copy()
        C_l_ = None
        if with_predictions:
            n_ex, n_t = eng.lik.t.shape
            pred = torch.empty((eng.n, n_ex, n_t), dtype=torch.float64, device=eng.device)
            eng._ck(eng.lib.smcb_predict_mm_progress(eng.h, eng.state.data_ptr(), eng.n, eng.n, pred.data_ptr(),
                                                     eng._stream))
            C_l_ = pred.cpu().numpy()
        return llk, C_l_

    def cal_prior(self, theta, priors=None):
        """Product of independent uniform pdfs (only `> 0` is used by the sampler)."""
        pr = self.eng.prior
        inside = pr.contains(np.asarray(theta))
        dens = 1.0 / np.prod(np.where(pr.high > pr.low, pr.high - pr.low, 1.0))
        return inside * dens


class LegacyNumpyStream:
    """The random inputs of a run drawn from NumPy's legacy global generator in the order the reference driver
    consumes it: `np.random.seed(seed)` and the prior draws at import (`Micmem_settings.py:47,69-87`), then per
    stage `rand()` (`Micmem_SMC_main.py:156`) and per sweep `multivariate_normal(...)` - i.e.
    `standard_normal(N*d).reshape(N, d)` through the SVD factor - followed by `uniform(0, 1, N)` (`:220,235`).
    Pass it as `Engine.run(particles, stream=...)` to reproduce a run of the reference draw for draw."""

    def __init__(self, seed=20250205):
        self.rs = np.random.RandomState(seed)

    def prior_uniform(self, low, high, N):
        return np.stack([self.rs.uniform(l, h, N) for l, h in zip(low, high)], axis=1)

    def u0(self):
        return self.rs.rand()

    def normals(self, N, d):
        return self.rs.standard_normal(N * d).reshape(N, d)

    def uniforms(self, N):
        return self.rs.uniform(0, 1, N)
