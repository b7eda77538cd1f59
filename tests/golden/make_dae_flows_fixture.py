#!/usr/bin/env python
"""Golden outlet flows of the transient reactor model from the CPU oracle, for the GPU parity test
at a size the oracle cannot reach inside a test run: 256 particles x the 30 operating conditions of dae_synth.npz
(7680 marches of oracle/methanation_dae.py, ~1.4 s each; ~25 minutes on 8 cores).  192 particles lie around the
data-generating parameters (x U[0.8, 1.25] per component), 64 are drawn from the reference's wide prior box
(methanation_set_conditon.py:64-70), where some marches fail and get the reference's -10000 penalty
(methanation_set_likelihood.py:244).  Writes tests/golden/dae_flows_256.npz (inputs and outputs only).

Run:  python tests/golden/make_dae_flows_fixture.py [n_workers]
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import kinetic, methanation_dae as dae  # noqa: E402


def _one(args):
    full_row, cond = args
    return dae.outlet_flows(full_row[None, :], cond)[0]


def main(workers):
    g = np.load(os.path.join(HERE, "dae_synth.npz"))
    kf = np.load(os.path.join(HERE, "kinetic_synth.npz"))
    cond, base, est, obs = g["cond"], g["base4"], g["est4"], g["obs"]
    rs = np.random.RandomState(20250205)
    near = base[est] * rs.uniform(0.8, 1.25, (192, len(est)))
    wide = rs.uniform(kf["low4"], kf["high4"], (64, len(est)))
    theta = np.vstack([near, wide])
    full = kinetic.assemble(theta, base, est)
    with mp.get_context("fork").Pool(workers) as pool:
        flows = np.array(pool.map(_one, [(row, cond) for row in full], chunksize=1))
    sigma = full[:, -1]
    ssr = np.sum((flows - obs[None, :, :]) ** 2, axis=(1, 2))
    lk = -(0.5 / sigma ** 2) * ssr - 5.0 * cond.shape[0] * np.log(sigma)
    np.savez_compressed(os.path.join(HERE, "dae_flows_256.npz"), theta=theta, flows=flows, lk=lk)
    print("wrote dae_flows_256.npz;", int((flows <= -9999).any(axis=(1, 2)).sum()), "particles with a failed march")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1))
