"""Priors.  The reference's active path only uses independent uniform priors as an in-support
indicator (`cal_prior(...) > 0`, SMC_example/Micmem_SMC_main.py:60-90,224-226;
SMC_methanation/methanation_functions.py:130-133), i.e. a closed box."""
import numpy as np


class UniformBox:
    def __init__(self, low, high, names=None):
        self.low = np.ascontiguousarray(low, dtype=np.float64)
        self.high = np.ascontiguousarray(high, dtype=np.float64)
        if self.low.shape != self.high.shape or self.low.ndim != 1:
            raise ValueError("low/high must be 1-D arrays of equal length")
        if np.any(self.high < self.low):
            raise ValueError("high < low")
        self.names = list(names) if names is not None else [f"p{i}" for i in range(self.d)]

    @property
    def d(self):
        return self.low.shape[0]

    @classmethod
    def from_priors(cls, priors):
        """From the reference's PyMC-like dict (`Micmem_settings.py:62-66`)."""
        low, high = [], []
        for name, cfg in priors.items():
            if cfg["dist"] != "uniform":
                raise NotImplementedError(
                    f"prior '{cfg['dist']}' for {name}: only uniform priors are on the accelerated path "
                    "(the reference's normal-prior branch is dead code, SURVEY.md 8(f) N2)")
            low.append(cfg["low"])
            high.append(cfg["high"])
        return cls(low, high, names=list(priors.keys()))

    def contains(self, theta):
        theta = np.asarray(theta)
        return np.all((theta >= self.low) & (theta <= self.high), axis=1)
