"""The C-ABI library: loads, exports every symbol include/smcb200.h declares, fails loudly without a GPU.
CPU only - no kernel is launched here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "smcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smcb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(pkg):
    lib = pkg._lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/smcb200.h but not exported"
    # the ctypes binding covers exactly the header
    assert sorted(pkg._lib.EXPORTS) == names


def test_version_and_no_silent_cpu_path(pkg):
    lib = pkg._lib.load()
    assert lib.smcb_version() >= 100
    import torch
    h = C.c_void_p()
    rc = lib.smcb_create(0, C.byref(h))
    if torch.cuda.is_available():
        assert rc == 0
        lib.smcb_destroy(h)
    else:
        assert rc < 0 and not h.value
        assert b"no CPU fallback" in lib.smcb_last_error(None)
        lik = pkg.MMRate.synthetic(16)
        prior = pkg.UniformBox([0, 0, 0], [10, 10, 10])
        with pytest.raises(RuntimeError):
            pkg.Engine(lik, prior, pkg.Settings(n_particle=8))


def test_product_package_never_imports_the_oracle():
    pkgdir = os.path.join(ROOT, "python-based-sequential-monte-carlo-method-with-likelihood-tempering_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_model_ids_and_likelihood_specs_match_the_header(pkg):
    """The Python likelihood classes carry the header's model ids; the transient reactor model accepts only the
    reference's 8 kinetic parameters."""
    import numpy as np
    text = open(os.path.join(ROOT, "include", "smcb200.h")).read()
    ids = {k: int(v) for k, v in re.findall(r"#define\s+SMCB_MODEL_([A-Z_]+)\s+(\d+)", text)}
    assert ids == {"MM_PROGRESS": 1, "MM_RATE": 2, "KINETIC_RK": 3, "KINETIC_DAE": 4, "USER": 5}
    assert pkg.UserKernelLikelihood.model_id == pkg.CallableLikelihood.model_id == 5
    assert (pkg.MMProgress.model_id, pkg.MMRate.model_id, pkg.KineticRK.model_id, pkg.KineticDAE.model_id) == (1, 2, 3, 4)
    cond = np.ones((3, pkg._lib.KIN_NCOND_FIELDS))
    obs = np.zeros((5, 3))
    base9 = np.arange(1.0, 10.0)
    lik = pkg.KineticDAE(cond, obs, base9, [0, 1, 2, 3, 8])
    assert (lik.n_pairs, lik.d, lik.n_obs) == (4, 5, 15)
    with pytest.raises(ValueError):
        pkg.KineticDAE(cond, obs, np.arange(1.0, 34.0), np.arange(32))       # 32-parameter family: plug-flow model only
    with pytest.raises(ValueError):
        pkg.KineticRK(cond, np.zeros((4, 3)), base9, [0, 1, 2, 3, 8])         # observations must be [5, n_cond]
