"""Run-artefact writer: files in the reference's formats (CPU only)."""
import os

import numpy as np
import pytest

REF_RUN = "/root/reference/SMC_methanation/methanation_SMC/20251124_183100_30"


def _reference_prior(n=1000, sigma_true_draws=30):
    """The prior cloud of the reference's saved 1000-particle methanation run, regenerated from its RNG stream
    (seed 20250205, 5 x standard_normal(30), then 5 x uniform(lo_i, hi_i, 1000))."""
    low = [-39.12, 0.0, -3.441e5, 0.0, 0.5]
    high = [339.04, 1.044e5, 3.5557e6, 2.901e5, 15.0]
    rs = np.random.RandomState(20250205)
    for _ in range(5):
        rs.standard_normal(sigma_true_draws)
    return np.stack([rs.uniform(l, h, n) for l, h in zip(low, high)], axis=1)


def test_writer_produces_the_reference_layout(pkg, tmp_path):
    w = pkg.RunWriter(str(tmp_path / "run"), names=["Af", "Eaf", "Ar", "Ear", "sigma"])
    p = _reference_prior()
    w.first(p)
    w.stage(3, p * 0.5)
    w.last(p[::-1])
    base = tmp_path / "run"
    assert sorted(os.listdir(base)) == ["Posterior_Distribution.csv", "pred"]
    assert sorted(os.listdir(base / "pred")) == ["3_p_pred.csv", "first_p_pred.csv", "last_p_pred.csv"]
    assert np.array_equal(np.loadtxt(base / "pred" / "first_p_pred.csv", delimiter=","), p)      # %.18e round-trips
    assert np.array_equal(np.loadtxt(base / "pred" / "3_p_pred.csv", delimiter=","), p * 0.5)
    lines = open(base / "Posterior_Distribution.csv").read().splitlines()
    assert lines[0] == "Af,Eaf,Ar,Ear,sigma" and len(lines) == 1001
    assert np.array_equal(np.array([[float(x) for x in ln.split(",")] for ln in lines[1:]]), p[::-1])
    import pandas as pd
    assert np.array_equal(pd.read_csv(base / "Posterior_Distribution.csv", float_precision="round_trip").values, p[::-1])


@pytest.mark.skipif(not os.path.exists(REF_RUN), reason="the reference checkout is only present in the build container")
def test_first_p_pred_is_byte_identical_to_the_reference_file(pkg, tmp_path):
    """The reference's saved `pred/first_p_pred.csv`, reproduced byte for byte from its RNG stream."""
    w = pkg.RunWriter(str(tmp_path / "run"))
    w.first(_reference_prior())
    ours = open(tmp_path / "run" / "pred" / "first_p_pred.csv", "rb").read()
    ref = open(os.path.join(REF_RUN, "pred", "first_p_pred.csv"), "rb").read()
    assert ours.replace(b"\r\n", b"\n") == ref.replace(b"\r\n", b"\n")
    # and the header + float formatting of Posterior_Distribution.csv
    ref_post = open(os.path.join(REF_RUN, "Posterior_Distribution.csv")).read().splitlines()
    vals = np.array([[float(x) for x in ln.split(",")] for ln in ref_post[1:]])
    w2 = pkg.RunWriter(str(tmp_path / "run2"), names=ref_post[0].split(","))
    w2.last(vals)
    assert open(tmp_path / "run2" / "Posterior_Distribution.csv").read().splitlines() == ref_post
