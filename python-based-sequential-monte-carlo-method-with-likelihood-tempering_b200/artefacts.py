"""Run artefacts in the reference's file formats, so that its plotting scripts can read runs
of this engine.

The reference writes, per run directory (`SMC_methanation/SMC_methanation_main.py:35-44,180-181,422,432-438`,
`methanation_functions.py:223-234`):

    pred/first_p_pred.csv          prior particles           np.savetxt(..., delimiter=',')   ('%.18e', no header)
    pred/{step}_p_pred.csv         particles after stage     same
    pred/last_p_pred.csv           posterior particles       same
    Posterior_Distribution.csv     posterior particles       pandas.DataFrame.to_csv(index=False), header = names

`RunWriter` produces exactly those; `RunWriter.hook` plugs into `Engine.run(hook=...)`.
"""
import os

import numpy as np


class RunWriter:
    def __init__(self, dirname, names=None):
        self.dir = dirname
        self.pred = os.path.join(dirname, "pred")
        os.makedirs(self.pred, exist_ok=True)
        self.names = list(names) if names is not None else None

    @staticmethod
    def _rows(p):
        return np.asarray(p.detach().cpu().numpy() if hasattr(p, "detach") else p, dtype=np.float64)

    def first(self, p_pred):
        """`np.savetxt(firstpred, p_pred, delimiter=',')` (main:180-181)."""
        np.savetxt(os.path.join(self.pred, "first_p_pred.csv"), self._rows(p_pred), delimiter=",")

    def stage(self, step, p_pred):
        """`np.savetxt(f'{dirnamepred}{step}_p_pred.csv', p_pred, delimiter=',')` (main:422)."""
        np.savetxt(os.path.join(self.pred, f"{int(step)}_p_pred.csv"), self._rows(p_pred), delimiter=",")

    def last(self, p_filt):
        """`SavePosteriorcsv` (functions:223-234): Posterior_Distribution.csv with a header + pred/last_p_pred.csv."""
        rows = self._rows(p_filt)
        names = self.names or [f"p{i}" for i in range(rows.shape[1])]
        with open(os.path.join(self.dir, "Posterior_Distribution.csv"), "w") as f:
            f.write(",".join(names) + "\n")
            for r in rows:
                f.write(",".join(repr(float(x)) for x in r) + "\n")   # what pandas writes for float64
        np.savetxt(os.path.join(self.pred, "last_p_pred.csv"), rows, delimiter=",")

    def hook(self, kind, **kw):
        """For `Engine.run(hook=writer.hook)`: dumps the particle cloud after every stage."""
        if kind == "stage":
            self.stage(kw["step"], kw["engine"].particles())
