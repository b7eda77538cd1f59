"""Sampler settings with the reference's names and defaults.

Mirrors the hyper-parameter block of `SMC_example/Micmem_settings.py:15-31,90` (identical in
`SMC_methanation/methanation_set_conditon.py:107-125`).  Fields below the divider are additions of
this engine; their defaults reproduce the reference's behaviour.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class Settings:
    n_particle: int = 1000
    ess_limit: float = 0.5
    mhstep_factor: float = 0.5          # diagonal of w_cov
    mhstep_factor_cov: float = 0.5      # off-diagonal of w_cov
    ad_mhstep_num: int = 20             # max MH sweeps once gamma == 1
    mhstep_num: int = 5                 # max MH sweeps while gamma < 1
    mhstep_ratio: float = 1.0           # reset to 1.0 at every stage (Micmem_SMC_main.py:190)
    r_threshold: float = 0.5
    r_threshold_f: float = 0.7
    r_threshold_min: float = 0.1
    d_gamma_max: float = 1
    gm_reduction_itr: int = 80
    gm_reduction_rate: float = 0.7
    itr_max: int = 50
    n_cores: int = 30                   # accepted for compatibility; unused (no ray pool)
    # ------------------------------------------------------------------ engine additions
    seed: int = 20250205
    temper_rule: str = "backoff"        # "backoff" (reference) | "bisect" (ESS bisection)
    scan_mode: str = "fixed"            # "fixed" (parallel, shard-invariant) | "sequential" (bit-exact reference sum)
    cand_batch: int = 8                 # tempering candidates evaluated per pass (<=16)
    early_exit: bool = True             # stop sweeps when moved fraction exceeds r_threshold (reference)
    early_reject: bool = True           # let the likelihood stop once the MH rejection is certain (exact decisions)
    mm_budget: int = 0                  # MM_PROGRESS: attempted RK steps before a solve moves to the tail kernel;
                                        # 0 = by size (Engine.mm_budget: 512 when the solves outnumber the lanes of the
                                        # bulk kernel many times over, down to 32 when every solve has a lane to itself)
    mm_refill_min: int = 8              # MM_PROGRESS: free lanes a warp waits for before setting up new solves ...
    mm_tail_warps: int = 32             # MM_PROGRESS: one-warp blocks per SM of the tail kernel
    mm_chunk: int = 32                  # MM_PROGRESS: particles per work-queue item of the bulk kernel
    mm_patience: int = 3                # ... and for how many steps (results do not depend on these three)
    fused_sweeps: int = 0               # >0: this many sweeps per library call (smcb_mh_sweeps: covariance refreshed every
                                        # sweep; kinetic model: smcb_mh_fused with a frozen factor); the early-exit and
                                        # step-halving rules then act between batches
    bisect_iters: int = 60
    factor: str = "auto"                # proposal factor: "host" = NumPy's SVD factor on the host (reproduces the reference's
                                        # proposals for given normals), "device" = Jacobi eigen-factor on the device (same
                                        # proposal distribution, no host round trip), "auto" = host when the random inputs
                                        # are supplied (parity mode), device otherwise

    @property
    def inv_Np(self):
        return 1 / self.n_particle

    def w_cov(self, d):
        """`w_cov` of Micmem_settings.py:93-96."""
        w = np.full((d, d), self.mhstep_factor_cov, dtype=np.float64)
        np.fill_diagonal(w, self.mhstep_factor)
        return w

    def validate(self):
        if self.n_particle < 1:
            raise ValueError("n_particle must be positive")
        if not (1 <= self.cand_batch <= 16):
            raise ValueError("cand_batch must be in 1..16")
        if self.temper_rule not in ("backoff", "bisect"):
            raise ValueError("temper_rule must be 'backoff' or 'bisect'")
        if self.scan_mode not in ("fixed", "sequential"):
            raise ValueError("scan_mode must be 'fixed' or 'sequential'")
        if self.factor not in ("auto", "host", "device"):
            raise ValueError("factor must be 'auto', 'host' or 'device'")
        if self.gm_reduction_itr < 1:
            raise ValueError("gm_reduction_itr must be at least 1 (the back-off loop tests at least one increment)")
        if not (0.0 < self.gm_reduction_rate < 1.0):
            raise ValueError("gm_reduction_rate must lie in (0, 1)")
        if self.itr_max < 2:
            raise ValueError("itr_max must be at least 2 (stages are numbered 1 .. itr_max-1)")
        if not (0.0 < self.ess_limit < 1.0):
            raise ValueError("ess_limit must lie in (0, 1)")
        if not (0.0 < self.d_gamma_max):
            raise ValueError("d_gamma_max must be positive")
        if self.fused_sweeps < 0 or self.mhstep_num < 0 or self.ad_mhstep_num < 0:
            raise ValueError("sweep counts must not be negative")
        return self
