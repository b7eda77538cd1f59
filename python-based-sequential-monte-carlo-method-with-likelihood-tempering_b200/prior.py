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

    # uniform priors only: no density ratio enters the MH test (it is 1 inside the box)
    has_normal = False


class IndependentPrior(UniformBox):
    """Independent normal / uniform components: the reference's PyMC-like `priors` dict with
    both branches of `cal_prior` (`SMC_example/Micmem_SMC_main.py:60-90`) and the density-ratio MH test
    `pp = exp(px*gamma) * (p0_2/p0_1)` (`SMC_methanation/SMC_methanation_main.py:359-375`), which the shipped
    configuration never reaches.

    dists: list of ("uniform", low, high) or ("normal", mu, sigma), one per parameter.  For the engine a uniform
    component is a closed interval in the box test, a normal one is unbounded and contributes
    -(x-mu)^2/(2 sigma^2) to the log prior."""

    def __init__(self, dists, names=None):
        low, high, mu, inv2var = [], [], [], []
        for kind, a, b in dists:
            if kind == "uniform":
                low.append(a); high.append(b); mu.append(0.0); inv2var.append(0.0)
            elif kind == "normal":
                if not b > 0:
                    raise ValueError("normal prior needs sigma > 0")
                low.append(-np.inf); high.append(np.inf); mu.append(a); inv2var.append(1.0 / (2.0 * b * b))
            else:
                raise ValueError(f"Unknown prior: {kind}")
        super().__init__(low, high, names)
        self.dists = [tuple(x) for x in dists]
        self.mu = np.ascontiguousarray(mu, dtype=np.float64)
        self.inv2var = np.ascontiguousarray(inv2var, dtype=np.float64)
        self.has_normal = bool(np.any(self.inv2var > 0))

    @classmethod
    def from_priors(cls, priors):
        """From the reference's dict (`Micmem_settings.py:54-66`)."""
        dists = []
        for name, cfg in priors.items():
            if cfg["dist"] == "uniform":
                dists.append(("uniform", cfg["low"], cfg["high"]))
            elif cfg["dist"] == "normal":
                dists.append(("normal", cfg["mu"], cfg["sigma"]))
            else:
                raise ValueError(f"Unknown prior: {cfg['dist']}")
        return cls(dists, names=list(priors.keys()))

    def log_ratio(self, theta_new, theta_old):
        """log p(theta_new) - log p(theta_old) over the normal components (uniform ones: 0 inside the box)."""
        a, b = np.asarray(theta_old) - self.mu, np.asarray(theta_new) - self.mu
        return np.sum(self.inv2var * (a * a - b * b), axis=1)

    def pdf(self, theta):
        """Product of the component densities, as `cal_prior` returns it."""
        import scipy.stats
        theta = np.asarray(theta, dtype=np.float64)
        out = np.ones(theta.shape[0])
        for j, (kind, a, b) in enumerate(self.dists):
            out *= (scipy.stats.norm.pdf(theta[:, j], loc=a, scale=b) if kind == "normal"
                    else scipy.stats.uniform.pdf(theta[:, j], loc=a, scale=b - a))
        return out
