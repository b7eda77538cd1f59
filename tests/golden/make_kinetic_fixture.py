#!/usr/bin/env python
"""Synthetic inputs for the methanation-style kinetic workloads (BASELINE configs 3 and 5).

The reference's operating-conditions file (`methanation_data/information.csv`,
`methanation_set_conditon.py:137`) is not in the upstream repository, so the conditions are
builder-chosen (`oracle.kinetic.synthetic_conditions`) and the observations follow the reference's own
recipe, data = model(baseparams) + N(0, sigma^2) (`SMC_methanation_main.py:89-95`), using the oracle's
CPU forward model.  Writes tests/golden/kinetic_synth.npz; bench.py and the tests read it.

Run:  python tests/golden/make_kinetic_fixture.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import kinetic  # noqa: E402

cond = kinetic.synthetic_conditions(30)
base4, base16 = kinetic.base_vector(4), kinetic.base_vector(16)
low4, high4 = kinetic.reference_box()
rs = np.random.RandomState(3)
theta4 = rs.uniform(low4, high4, (64, 5))
theta16 = rs.uniform(np.minimum(base16[:32] * 0.8, base16[:32] * 1.2), np.maximum(base16[:32] * 0.8, base16[:32] * 1.2),
                     (64, 32))
obs4, obs16 = kinetic.synthetic_observations(cond, base4), kinetic.synthetic_observations(cond, base16)
np.savez_compressed(
    os.path.join(os.path.dirname(os.path.abspath(__file__)), "kinetic_synth.npz"),
    cond=cond, base4=base4, base16=base16, obs4=obs4, obs16=obs16, low4=low4, high4=high4,
    est4=np.array(kinetic.EST_POSITION, dtype=np.int32),
    # known answers of the oracle model (64 particles each) for the GPU parity tests
    theta4=theta4, lk4=kinetic.loglik(theta4, cond, obs4, base4, kinetic.EST_POSITION, 50),
    theta16=theta16, lk16=kinetic.loglik(theta16, cond, obs16, base16, np.arange(32), 50))
print("wrote kinetic_synth.npz")
