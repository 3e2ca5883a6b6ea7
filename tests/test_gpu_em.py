"""GPU parity at the EM level: committed golden vectors, the Rcpp-named API, and whole EM
trajectories with the backend swapped under one driver (SURVEY.md 8c protocol T2).

Tolerances (north_star): Viterbi states and called peaks identical; posteriors <= 1e-8 absolute;
log-likelihood (BIC), Q and parameters <= 1e-9 relative at every EM iteration.
"""
import os

import numpy as np
import pytest

from oracle_backend import OracleBackend

from epigrahmm_b200 import em, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MINZERO = 2.2250738585072014e-308


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def test_golden_consensus(ehmm):
    g = np.load(os.path.join(GOLD, "golden_small.npz"))
    th = synth.THETA_CONSENSUS
    M = g["c_counts"].shape[0]
    with ehmm.Session("gold_c", M, 2) as s:
        s.set_data(g["c_counts"], g["c_offsets"])
        s.set_groups(g["c_group"], g["c_u_chip"], g["c_u_offsets"])
        s.loglik_consensus(th["psi"], MINZERO)
        assert _rel(s.fetch("logLikelihood"), g["c_logf"]) <= 1e-11
        s.expstep(th["pi"], th["gamma"])
        for n in ("logFP", "logBP"):
            assert _rel(s.fetch(n), g["c_" + n]) <= 1e-9
        for n in ("logProb1", "logProb2"):
            assert np.max(np.abs(np.exp(s.fetch(n)) - np.exp(g["c_" + n]))) <= 1e-8
        m = s.maxstepprob()
        assert _rel(m["pi"], g["c_pi"]) <= 1e-9 and _rel(m["gamma"], g["c_gamma"]) <= 1e-9
        assert _rel(s.qfunction(th["pi"], th["gamma"]), g["c_q"]) <= 1e-9
        assert _rel(s.bic(5, 2), g["c_bic"]) <= 1e-9
        assert np.array_equal(s.viterbi(th["pi"], th["gamma"]), g["c_viterbi"])
        assert s.rc_consensus(0.05, uniforms=g["c_unif"]) == int(g["c_rc_n"])
        for k in range(2):
            assert np.allclose(s.rc_buckets(k), g["c_rc"][k], rtol=1e-9, atol=1e-12)
        v, gr = s.glm_nb(1, [th["psi"][1][0], 1 / th["psi"][1][1]])
        assert _rel(v, g["c_glm_v"]) <= 1e-9 and np.allclose(gr, g["c_glm_g"], rtol=1e-7, atol=1e-7)


def test_golden_differential(ehmm):
    g = np.load(os.path.join(GOLD, "golden_small.npz"))
    th = synth.THETA_DIFFERENTIAL
    M = g["d_counts"].shape[0]
    delta = np.full(6, 1 / 6)
    with ehmm.Session("gold_d", M, 3) as s:
        s.set_data(g["d_counts"], g["d_offsets"])
        s.set_design(g["d_design"])
        s.set_groups(g["d_group"], g["d_u_chip"], g["d_u_offsets"], None, g["d_u_dsg"])
        s.loglik_differential_hmm(delta, th["psi"], MINZERO)
        assert _rel(s.fetch("logLikelihood"), g["d_logf"]) <= 1e-11
        s.expstep(th["pi"], th["gamma"])
        for n in ("logProb1", "logProb2"):
            assert np.max(np.abs(np.exp(s.fetch(n)) - np.exp(g["d_" + n]))) <= 1e-8
        s.inner_expstep(delta, th["psi"], MINZERO)
        assert np.max(np.abs(s.fetch("mixtureProb") - g["d_eta"])) <= 1e-10
        assert _rel(s.inner_maxstepprob(), g["d_delta"]) <= 1e-9
        m = s.maxstepprob()
        assert _rel(m["pi"], g["d_pi"]) <= 1e-9 and _rel(m["gamma"], g["d_gamma"]) <= 1e-9
        assert _rel(s.qfunction(th["pi"], th["gamma"]), g["d_q"]) <= 1e-9
        assert _rel(s.bic(9, 6), g["d_bic"]) <= 1e-9
        assert np.array_equal(s.viterbi(th["pi"], th["gamma"]), g["d_viterbi"])
        assert s.rc_differential(0.05, uniforms=g["d_unif"]) == int(g["d_rc_n"])
        for c in range(8):
            assert np.allclose(s.rc_buckets(c), g["d_rc"][c], rtol=1e-9, atol=1e-12)


def test_rcpp_named_api(ehmm, oracle):
    """the exported functions by their reference names, keyed by the hdf5 path"""
    from epigrahmm_b200 import api
    M, K = 3001, 2
    d = synth.make_consensus(M, 2, seed=77)
    g = synth.make_groups(d["counts"], d["offsets"])
    th = synth.THETA_CONSENSUS
    logf = oracle.loglikConsensusHMM(d["counts"], d["offsets"], th["psi"], MINZERO)
    h = oracle.expStep(th["pi"], th["gamma"], logf)
    path = "/tmp/epigraHMM_test.h5"
    api.expStep(pi=th["pi"], gamma=th["gamma"], logf=logf, hdf5=path)
    s = api.session(path)
    assert sorted(s.datasets()) == ["logBP", "logFP", "logLikelihood", "logProb1", "logProb2"]
    m, mo = api.maxStepProb(path), oracle.maxStepProb(h)
    assert _rel(m["pi"], mo["pi"]) <= 1e-9 and _rel(m["gamma"], mo["gamma"]) <= 1e-9
    assert _rel(api.computeQFunction(path, th["pi"], th["gamma"]), oracle.computeQFunction(h, th["pi"], th["gamma"])) <= 1e-9
    assert _rel(api.computeBIC(path, 5, 2), oracle.computeBIC(h, 5, 2)) <= 1e-9
    assert np.max(np.abs(api.getMarginalProbability(path) - np.exp(h["logProb1"]))) <= 1e-8
    assert np.array_equal(api.computeViterbiSequence(path, th["pi"], th["gamma"]),
                          oracle.computeViterbiSequence(h, th["pi"], th["gamma"]))
    assert "viterbi" in s.datasets()
    # rejection control consumes the module-level R stream, like RNGScope
    s.set_data(d["counts"], d["offsets"])
    api.set_seed(210423)
    rng = oracle.RNG.set_seed(210423)
    tabs = api.consensusRejectionControlled(path, g["group"], 0.05)
    ref, _ = oracle.consensusRejectionControlled(h, g["group"], 0.05, rng=rng)
    for a, b in zip(tabs, ref):
        assert np.array_equal(a[:, 1], b[:, 1]) and np.allclose(a[:, 0], b[:, 0], rtol=1e-12)
    # the stand-alone helpers continue the same stream
    x = np.random.default_rng(0).random(1000) ** 3
    assert np.array_equal(api.reweight(x, 0.2), oracle.reweight(x, 0.2, rng=rng)[0])
    pr = np.random.default_rng(1).random(500)
    assert np.array_equal(api.rbinomVectorized(pr), oracle.reweight(pr, 1.0, rng=rng)[0])
    f = np.random.default_rng(2).integers(1, 50, 1000)
    assert np.allclose(api.aggregate(x, f), oracle.aggregate(x, f), rtol=1e-12)
    api.close_session(path)
    with pytest.raises(ehmm.EhmmError):
        api.maxStepProb(path)


def _compare_fits(fg, fo, par_tol):
    assert len(fg["parHist"]) == len(fo["parHist"])
    worst = 0.0
    for it, (pg, po) in enumerate(zip(fg["parHist"], fo["parHist"])):
        r = _rel(em.unlist(pg), em.unlist(po))
        worst = max(worst, r)
        assert r <= par_tol, (it + 1, r)
        hg, ho = fg["controlHist"][it], fo["controlHist"][it]
        assert _rel(hg["bic"], ho["bic"]) <= 1e-9 and _rel(hg["q"], ho["q"]) <= 1e-9, it + 1
        assert hg["convergence"] == ho["convergence"]
    return worst


def test_consensus_em_trajectory(ehmm, oracle):
    M, N = 30011, 2
    d = synth.make_consensus(M, N, seed=20261002)
    g = synth.make_groups(d["counts"], d["offsets"])
    th = synth.THETA_CONSENSUS
    theta = dict(pi=th["pi"], gamma=th["gamma"], psi=[np.array([0.5, 2.0]), np.array([2.0, 2.0])])
    ctl = em.controlEM(maxIterEM=8)
    rng = oracle.RNG.set_seed(210423)
    be = OracleBackend(d["counts"], d["offsets"], 2, groups=g, rng=rng)
    fo = em.emConsensus(theta, be, N, ctl)
    with ehmm.Session("traj_c", M, 2) as s:
        s.set_data(d["counts"], d["offsets"])
        s.set_groups(g["group"], g["u_chip"], g["u_offsets"])
        s.rng_set_state(oracle.RNG.set_seed(210423).state())
        fg = em.emConsensus(theta, s, N, ctl)
        worst = _compare_fits(fg, fo, 1e-9)
        print("consensus trajectory: worst relative parameter difference %.3g over %d iterations" % (worst, len(fg["parHist"])))
        assert np.array_equal(s.rng_get_state(), rng.state())       # same number of uniforms consumed
        tg, to = fg["theta_new"], fo["theta_new"]
        vg, vo = s.viterbi(tg["pi"], tg["gamma"]), be.viterbi(to["pi"], to["gamma"])
        assert np.array_equal(vg, vo)
        assert np.array_equal(em.callPeaks(s, 0.05), em.callPeaks(be, 0.05))
        assert np.max(np.abs(np.exp(s.fetch("logProb1")) - np.exp(be.fetch("logProb1")))) <= 1e-8


def test_differential_em_trajectory(ehmm, oracle):
    M = 12007
    d = synth.make_differential(M, 3, 2, seed=20261003)
    g = synth.make_groups(d["counts"], d["offsets"], design=d["design"])
    th = synth.THETA_DIFFERENTIAL
    theta = dict(pi=th["pi"], gamma=th["gamma"], delta=np.full(d["B"], 1.0 / d["B"]),
                 psi=[np.array([0.5, 0.8]), np.array([1.5, 0.2])])
    ctl = em.controlEM(maxIterEM=4)
    rng = oracle.RNG.set_seed(99)
    be = OracleBackend(d["counts"], d["offsets"], 3, design=d["design"], groups=g, rng=rng)
    fo = em.outerEMDifferential(theta, be, 6, ctl)
    with ehmm.Session("traj_d", M, 3) as s:
        s.set_data(d["counts"], d["offsets"])
        s.set_design(d["design"])
        s.set_groups(g["group"], g["u_chip"], g["u_offsets"], None, g["u_dsg"])
        s.rng_set_state(oracle.RNG.set_seed(99).state())
        fg = em.outerEMDifferential(theta, s, 6, ctl)
        worst = _compare_fits(fg, fo, 1e-9)
        print("differential trajectory: worst relative parameter difference %.3g" % worst)
        assert np.array_equal(s.rng_get_state(), rng.state())
        tg, to = fg["theta_new"], fo["theta_new"]
        assert np.array_equal(s.viterbi(tg["pi"], tg["gamma"]), be.viterbi(to["pi"], to["gamma"]))


def test_full_flow_like_reference_callpeaks(ehmm):
    """tests/testthat/test-callPeaks.R:1-39 through initializer() + epigraHMM() + callPeaks()"""
    from epigrahmm_b200 import core, rng
    from test_em_host import _callpeaks_data
    counts, offsets, truth = _callpeaks_data()
    rng.set_seed(210423)
    ctl = em.controlEM()
    peaks = core.initializer(counts, offsets, ctl)
    assert peaks.shape == counts.shape and set(np.unique(peaks)) <= {0.0, 1.0}
    fit = core.epigraHMM(counts, offsets, peaks, ["A", "A"], ctl, type="consensus", dist="nb", key="/tmp/flow.h5")
    s = fit["session"]
    try:
        assert sorted(s.datasets()) == ["logBP", "logFP", "logLikelihood", "logProb1", "logProb2", "viterbi"]
        for n, a in s.datasets().items():
            assert a.shape[0] == 4000 and np.isfinite(a).all(), n
        for method in ("viterbi", 0.05):
            call = em.callPeaks(s, method)
            assert (truth & ~call).sum() / truth.sum() <= 0.05
    finally:
        s.close()
