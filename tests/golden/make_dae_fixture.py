#!/usr/bin/env python
"""Golden residuals of the reference's transient reactor DAE.

`SMC_methanation/methanation_set_likelihood.py` cannot be imported here (it imports assimulo, ray, ... at module
top), so this script lifts the *source text* of the three pure functions `func_rCH4`, `func_rohg` and `reaction`
out of the reference file with `ast` (numba decorators dropped) and the constants they use out of
`methanation_set_conditon.py:74-89`, executes them unmodified in a scratch namespace and evaluates `reaction` on
seeded random states.  Nothing of the reference is copied into the repository: only inputs and outputs are
stored (tests/golden/methanation_dae_residual.npz).  Runs only where /root/reference exists.

Run:  python tests/golden/make_dae_fixture.py
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import kinetic  # noqa: E402

REF = "/root/reference/SMC_methanation"


def lift(path, names):
    tree = ast.parse(open(path, encoding="utf-8").read())
    out = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []
            out.append(node)
    return ast.Module(body=out, type_ignores=[])


def lift_constants(path, names):
    tree = ast.parse(open(path, encoding="utf-8").read())
    out = []
    for node in tree.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) \
                and node.targets[0].id in names:
            out.append(node)
    return ast.Module(body=out, type_ignores=[])


ns = {"np": np}
exec(compile(lift_constants(os.path.join(REF, "methanation_set_conditon.py"),
                            {"NX", "pi", "sc", "Dz", "rhos", "Hr", "R", "Rr", "S", "Cpg", "Cps", "keff", "dint", "U"}),
             "conditon", "exec"), ns)
exec(compile(lift(os.path.join(REF, "methanation_set_likelihood.py"), {"func_rCH4", "func_rohg", "reaction"}),
             "likelihood", "exec"), ns)
NX = ns["NX"]
cond = kinetic.synthetic_conditions(6)
rs = np.random.RandomState(11)
X, dX, P, RES = [], [], [], []
for case in range(12):
    row = cond[case % len(cond)]
    k8 = kinetic.BASEPARAMS * rs.uniform(0.7, 1.3, 8)
    x = np.empty(7 * NX)
    for k in range(5):
        x[k * NX:(k + 1) * NX] = max(row[k], 1.0) * rs.uniform(0.2, 1.5, NX)
    x[5 * NX:6 * NX] = rs.uniform(400.0, 700.0, NX)
    x[6 * NX:7 * NX] = row[7] * rs.uniform(0.5, 2.0, NX)
    dx = rs.normal(0.0, 1.0, 7 * NX) * np.repeat([5, 5, 5, 5, 5, 20, 0.1], NX)
    p0 = np.concatenate([row[:9], [row[9] / (NX - 1)], k8])     # my_model :164
    res = ns["reaction"](0.0, x, dx, p0)
    X.append(x), dX.append(dx), P.append(np.concatenate([row, k8])), RES.append(np.asarray(res, dtype=np.float64))
np.savez_compressed(os.path.join(HERE, "methanation_dae_residual.npz"), X=np.array(X), dX=np.array(dX), P=np.array(P),
                    RES=np.array(RES))
print("wrote methanation_dae_residual.npz", np.array(RES).shape)

# ---- synthetic observations of the transient model for the 30 builder-chosen operating conditions:
# data = model(baseparams) + N(0, sigma^2) as in SMC_methanation_main.py:89-95, with the oracle's implicit-Euler march
from oracle import methanation_dae as dae  # noqa: E402

cond30 = kinetic.synthetic_conditions(30)
base4 = kinetic.base_vector(4)
F = dae.outlet_flows(base4[None, :], cond30)[0]
obs = F + kinetic.SIGMA_TRUE * np.random.RandomState(20250206).standard_normal(F.shape)
np.savez_compressed(os.path.join(HERE, "dae_synth.npz"), cond=cond30, base4=base4, flows=F, obs=obs,
                    est4=np.array(kinetic.EST_POSITION, dtype=np.int32))
print("wrote dae_synth.npz", F.shape, "failed marches:", int(np.sum(F == kinetic.FAIL_FLOW)))
