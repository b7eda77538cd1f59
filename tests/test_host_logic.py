"""Host-side rules that need no GPU."""
import importlib
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@pytest.fixture(scope="module")
def engine_mod():
    import smcb200  # noqa: F401  (registers the hyphenated package directory under this name)
    return importlib.import_module(smcb200.__name__ + ".engine") if hasattr(smcb200, "__name__") else None


def test_deferral_budget_follows_the_solves_per_lane(engine_mod):
    """Settings.mm_budget = 0: 32 when every solve has a lane of the bulk kernel to itself, 512 when the solves outnumber
    the lanes many times over (profiles/budget_by_size_r02.log); monotone in between."""
    f = engine_mod.auto_mm_budget
    sm = 148
    assert f(1, 6, sm) == 32 and f(1000, 6, sm) == 32 and f(16384, 6, sm) == 32     # <= 113 664 lanes
    assert f(1 << 17, 6, sm) == 128 and f(1 << 18, 6, sm) == 128
    assert f(1 << 19, 6, sm) == 256
    assert f(1 << 20, 6, sm) == 512 and f(1 << 23, 6, sm) == 512
    assert f(1 << 20, 1, sm) == 128                                                   # one experiment: a sixth of the solves
    prev = 0
    for lg in range(0, 27):
        b = f(1 << lg, 6, sm)
        assert b >= prev
        prev = b


def test_settings_default_budget_is_by_size():
    import smcb200
    assert smcb200.Settings().mm_budget == 0
