"""The oracle against the golden run of the unmodified reference (CPU only)."""
import numpy as np

from oracle import mm, smc


def test_prior_stream_is_bit_exact(golden):
    st = smc.ReferenceStream(int(golden["seed"]))
    p = st.prior_uniform([0, 0, 0], [10, 10, 10], 1000)
    assert np.array_equal(p, golden["prior_particles"])


def test_known_answers_scipy_and_twin(golden):
    d = (golden["data_t"], golden["data_P"], golden["data_S0"])
    for theta, want in zip(golden["ka_in"], golden["ka_out"]):
        assert mm.loglik_progress_scipy(theta, *d) == want            # same scipy, same bits
        assert abs(mm.loglik_progress_twin(theta, *d) - want) <= 1e-11 * abs(want)
    # SURVEY.md section 4 known answers
    assert abs(golden["ka_out"][0] - 593.9635684697922) < 1e-9
    assert abs(golden["ka_out"][1] - 424.5676547271564) < 1e-9


def test_twin_matches_reference_on_prior_and_posterior(golden):
    """Scalar DOPRI5 twin vs the likelihoods the reference computed: first sweep (prior, heavy
    tailed step counts) and last sweep (near posterior)."""
    d = (golden["data_t"], golden["data_P"], golden["data_S0"])
    for sweep in (0, 33):
        P = golden["sweeps_in"][sweep][:250]
        want = golden["sweeps_out"][sweep][:250]
        got = mm.sweep_progress(P, *d, which="twin")
        rel = np.abs(got - want) / np.abs(want)
        assert rel.max() < 1e-10, rel.max()


def test_predictions_match(golden):
    d = (golden["data_t"], golden["data_P"], golden["data_S0"])
    for i in range(3):
        _, pred = mm.loglik_progress_scipy(golden["prior_particles"][i], *d, return_pred=True)
        assert np.array_equal(pred, golden["pmodel0"][i])


def test_loop_replays_reference_bit_for_bit(golden):
    """Oracle sampler loop driven by the recorded likelihoods reproduces every proposal, every
    stage scalar and the final particles of the reference run."""
    st = smc.ReferenceStream(int(golden["seed"]))
    pp = st.prior_uniform([0, 0, 0], [10, 10, 10], 1000)
    k = [0]

    def ll(p):
        i = k[0]
        k[0] += 1
        assert np.array_equal(p, golden["sweeps_in"][i]), f"proposals of sweep {i} differ"
        return golden["sweeps_out"][i]

    p, lk, tr = smc.run(ll, pp, np.zeros(3), np.full(3, 10.0), smc.Settings(), st)
    T = golden["stage_table"]
    assert k[0] == 34
    assert np.array_equal(np.array(tr.gamma), T[:, 4])
    assert np.array_equal(np.array(tr.ess), T[:, 2])
    assert np.array_equal(np.array(tr.max_lk), T[:, 3])
    assert np.array_equal(np.array(tr.moved), T[:, 5])
    assert np.array_equal(np.array(tr.n_mh) - 1, T[:, 1])
    assert np.array_equal(p, golden["final_particles"])
    assert np.array_equal(lk, golden["final_lk"])
    assert tr.n_backoff == [17, 17, 16, 14, 13, 13, 13, 13, 12, 10, 8, 6, 3, 0]
    assert abs(tr.log_evidence[-1] - 567.031312) < 1e-5
    assert all(len(a) == 1000 and np.all(np.diff(a) >= 0) for a in tr.ancestors)


def test_mvn_factor_matches_numpy_legacy(golden):
    """x = Z @ (sqrt(s)[:,None]*Vt) reproduces np.random.multivariate_normal draws."""
    rs = np.random.RandomState(123)
    cov = golden["draws_mvn_cov"][5]
    state = rs.get_state()
    x = rs.multivariate_normal(np.zeros(3), cov, 50)
    rs.set_state(state)
    Z = rs.standard_normal(150).reshape(50, 3)
    assert np.array_equal(np.dot(Z, smc.proposal_factor(cov)), x)
