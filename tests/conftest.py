import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Run of the unmodified reference (tests/golden/make_golden_mm.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "mm_reference_run.npz"))


@pytest.fixture(scope="session")
def pkg():
    import smcb200
    return smcb200


@pytest.fixture(scope="session")
def mm_engine_factory(golden, pkg):
    """Engine on the reference MM data with the reference prior box."""
    def make(n, **kw):
        lik = pkg.MMProgress(golden["data_t"], golden["data_P"], golden["data_S0"])
        prior = pkg.UniformBox([0, 0, 0], [10, 10, 10], names=["Vmax", "Km", "sigma"])
        cfg = pkg.Settings(n_particle=n, **kw)
        return pkg.Engine(lik, prior, cfg)
    return make
