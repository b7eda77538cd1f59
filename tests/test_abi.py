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
