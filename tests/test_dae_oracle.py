"""Transient reactor DAE: the oracle's restatement of the reference's residual against
residuals computed by the reference's own function text (tests/golden/make_dae_fixture.py), and the oracle's
implicit-Euler march against physical invariants.  CPU only."""
import os

import numpy as np
import pytest

from oracle import kinetic
from oracle import methanation_dae as dae

HERE = os.path.dirname(os.path.abspath(__file__))


def test_residual_matches_the_reference_function():
    g = np.load(os.path.join(HERE, "golden", "methanation_dae_residual.npz"))
    assert g["RES"].shape == (12, 7 * dae.NX)
    for x, dx, p, want in zip(g["X"], g["dX"], g["P"], g["RES"]):
        got = dae.residual(dae.from_reference_order(x), dae.from_reference_order(dx), p[:10], p[10:18])
        got = dae.to_reference_order(got)
        assert np.max(np.abs(got - want) / (np.abs(want) + 1e-300)) < 1e-12


def test_march_reaches_a_steady_state_that_conserves_atoms():
    cond = kinetic.synthetic_conditions(2)
    k8 = kinetic.BASEPARAMS
    for row in cond:
        Y, ok, hist = dae.integrate(row, k8, return_history=True)
        assert ok and len(hist) == len(dae.time_grid())
        # steady at 75 s: the last 5 s step changes nothing any more
        assert np.max(np.abs(hist[-1] - hist[-3]) / (np.abs(hist[-1]) + dae.FLOOR[:, None])) < 1e-6
        # residual of the steady equations
        F = dae.residual(Y, np.zeros_like(Y), row, k8)
        assert np.max(np.abs(F[:5, 1:-1])) < 1e-5 * np.max(np.abs(row[7] * row[:5] / (row[9] / 50)))
        # outlet fluxes u*C conserve C, H and O atoms and argon (CO2 + 4 H2 -> CH4 + 2 H2O)
        fin, fout = row[7] * row[:5], Y[6, -1] * Y[:5, -1]
        carbon = lambda f: f[1] + f[2]
        hydrogen = lambda f: 2 * f[0] + 4 * f[2] + 2 * f[3]
        oxygen = lambda f: 2 * f[1] + f[3]
        for atoms in (carbon, hydrogen, oxygen, lambda f: f[4]):
            assert abs(atoms(fout) / atoms(fin) - 1) < 2e-3      # dispersion at the inlet node is one-sided
        assert fout[2] > 0 and Y[5].max() > row[6]                # methane is made, the bed runs hotter than the jacket


def test_transient_and_plug_flow_models_agree_on_conversion():
    """Two discretisations of the same balances: the outlet methane flows differ by the dispersion / grid terms only."""
    cond = kinetic.synthetic_conditions(3)
    base = kinetic.base_vector(4)
    Fd = dae.outlet_flows(base[None, :], cond)[0]
    Fp = kinetic.outlet_flows(base[None, :], cond)[0]
    assert np.all(Fd > -1)
    assert np.max(np.abs(Fd[2] / Fp[2] - 1)) < 0.05
    assert np.max(np.abs(Fd[4] / Fp[4] - 1)) < 1e-6               # argon passes through


def test_a_failed_grid_step_is_retried_in_pieces():
    """Fast kinetics (A_f 14x the data-generating value, E_f halved): Newton diverges on the bare grid during ignition;
    with the retry rule the march goes through and lands on a steady state of the same balances."""
    g = np.load(os.path.join(HERE, "golden", "dae_synth.npz"))
    row = g["cond"][0]
    k8 = kinetic.BASEPARAMS.copy()
    k8[:4] = [1.87749225e+02, 2.77515159e+04, 7.66117997e+05, 7.35659385e+04]
    Y, fac, bare_ok = dae.start_state(row), {}, True
    with np.errstate(all="ignore"):
        for H in dae.time_grid():
            Y, ok, _ = dae._attempt(Y.copy(), Y, H, row, k8, fac)
            if not ok:
                bare_ok = False
                break
    assert not bare_ok
    Y, ok = dae.integrate(row, k8)
    assert ok and np.all(np.isfinite(Y))
    F = dae.residual(Y, np.zeros_like(Y), row, k8)
    assert np.max(np.abs(F[:5, 1:-1])) < 1e-5 * np.max(np.abs(row[7] * row[:5] / (row[9] / 50)))
