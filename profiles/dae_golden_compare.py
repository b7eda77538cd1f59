"""Device (KINETIC_DAE) against the oracle's offline golden flows (tests/golden/dae_flows_256.npz): per-particle table."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _abi import Abi  # noqa: E402

gd = os.path.join(ROOT, "tests", "golden")
g, fx = np.load(os.path.join(gd, "dae_synth.npz")), np.load(os.path.join(gd, "dae_flows_256.npz"))
cond, base, obs = (np.ascontiguousarray(g[k]) for k in ("cond", "base4", "obs"))
est = np.ascontiguousarray(g["est4"], dtype=np.int32)
a = Abi(4096, 8)
a.ck(a.lib.smcb_set_data_kinetic(a.h, cond.ctypes.data, obs.ctypes.data, cond.shape[0], base.ctypes.data, 4,
                                 est.ctypes.data, len(est), 1))
th, want, flows = fx["theta"], fx["lk"], fx["flows"]
got = a.loglik(4, th)
nfail = (flows <= -9999).any(axis=1).sum(axis=1)      # failed conditions per particle in the oracle
rel = np.abs(got - want) / np.abs(want)
print("near (192): max rel", rel[:192].max())
print("idx  oracle_failed_conditions  oracle_lk  device_lk  rel")
for i in range(192, 256):
    print(f"{i:4d} {nfail[i]:4d} {want[i]:16.6e} {got[i]:16.6e} {rel[i]:10.3e}")
np.save(os.path.join(ROOT, "gpurun_out", "dae_golden_device_lk.npy"), got)
