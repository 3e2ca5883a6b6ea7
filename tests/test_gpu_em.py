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


# Free-running trajectories (same driver, same R stream, backend swapped): north_star's gate, parameters
# within 1e-9 relative at EVERY iteration.  psi comes out of L-BFGS-B, whose stopping rule only determines
# the optimum to ~1e-6, so the gate holds only while both backends take the same optimiser path; with the
# rejection tables accumulated exactly (csrc/rc.cu: order-independent integer limbs) the GPU's fn / gr are
# the same bits on every run and agree with the oracle's to ~1e-14, which keeps the paths together.
FREE_TOL = 1e-9


def _compare_fits(fg, fo, par_tol):
    assert len(fg["parHist"]) == len(fo["parHist"])
    worst = 0.0
    for it, (pg, po) in enumerate(zip(fg["parHist"], fo["parHist"])):
        r = _rel(em.unlist(pg), em.unlist(po))
        worst = max(worst, r)
        assert r <= min(par_tol, FREE_TOL), (it + 1, r)
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
        if worst <= 1e-9:
            assert np.array_equal(s.rng_get_state(), rng.state())   # same number of uniforms consumed
        tg, to = fg["theta_new"], fo["theta_new"]
        vg, vo = s.viterbi(tg["pi"], tg["gamma"]), be.viterbi(to["pi"], to["gamma"])
        assert np.mean(vg != vo) <= 1e-4
        assert np.mean(em.callPeaks(s, 0.05) != em.callPeaks(be, 0.05)) <= 1e-4
        assert np.max(np.abs(np.exp(s.fetch("logProb1")) - np.exp(be.fetch("logProb1")))) <= max(1e-8, 10 * worst)
        if worst <= 1e-9:   # same optimiser path throughout: the strict gates
            assert np.array_equal(vg, vo) and np.array_equal(em.callPeaks(s, 0.05), em.callPeaks(be, 0.05))


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
        if worst <= 1e-9:
            assert np.array_equal(s.rng_get_state(), rng.state())
        tg, to = fg["theta_new"], fo["theta_new"]
        assert np.mean(s.viterbi(tg["pi"], tg["gamma"]) != be.viterbi(to["pi"], to["gamma"])) <= (0 if worst <= 1e-9 else 1e-4)


def _capture(run, theta0, rng):
    """oracle trajectory with (theta_old, RNG state) at the start of every iteration"""
    starts = [(theta0, rng.state().copy())]

    def cb(hist, theta_new):
        import copy
        starts.append((copy.deepcopy(theta_new), rng.state().copy()))
    fit = run(cb)
    return fit, starts[:-1]


def _psi_equivalent(s, column, pg, po):
    """psi agrees to 1e-9, or the two points are indistinguishable under R's own stopping rule
    (|f(a) - f(b)| <= factr * eps * |f|, factr = 1e7) and close (1e-5)"""
    r = _rel(pg, po)
    if r <= 1e-9:
        return True
    fa = s.glm_nb(column, em.invertDispersion(pg))[0]
    fb = s.glm_nb(column, em.invertDispersion(po))[0]
    return r <= 1e-5 and abs(fa - fb) <= 1e7 * np.finfo(float).eps * max(abs(fa), abs(fb), 1.0)


@pytest.mark.parametrize("data", ["helas3", "synthetic"])
def test_teacher_forced_iterations(ehmm, oracle, data):
    """Every EM iteration of an oracle trajectory replayed on the GPU from the SAME inputs (theta.old
    and .Random.seed): pi, gamma, Q, BIC <= 1e-9 relative, identical uniform consumption and next
    RNG state, rejection tables with identical support and sums <= 1e-9, psi <= 1e-9 or
    equivalent under the optimiser's own stopping rule."""
    import copy
    from epigrahmm_b200 import core
    if data == "helas3":
        a = np.load(os.path.join(GOLD, "helas3_counts.npz"))["counts"]
        counts = np.asfortranarray(a[:, 2:4])
        offsets = np.zeros(counts.shape, order="F")
        score = np.log1p(counts)
        z = 1 + ((score > np.quantile(score, 0.75, axis=0)).sum(1) > 0)
        ctl = em.controlEM(maxIterEM=10)
        theta0 = dict(pi=np.array([0.999, 0.001]), gamma=em.estimateTransitionProb(z, 2),
                      psi=core.estimateCoefficients(z, counts, offsets, ctl, "consensus"))
    else:
        d = synth.make_consensus(40009, 3, seed=20261009)
        counts, offsets = d["counts"], d["offsets"]
        ctl = em.controlEM(maxIterEM=6)
        th = synth.THETA_CONSENSUS
        theta0 = dict(pi=th["pi"], gamma=th["gamma"], psi=[np.array([0.5, 2.0]), np.array([2.0, 2.0])])
    M, N = counts.shape
    g = synth.make_groups(counts, offsets)
    rng = oracle.RNG.set_seed(210423)
    be = OracleBackend(counts, offsets, 2, groups=g, rng=rng)
    fo, starts = _capture(lambda cb: em.emConsensus(theta0, be, N, ctl, callback=cb), theta0, rng)
    strict = 0
    with ehmm.Session("forced", M, 2) as s:
        s.set_data(counts, offsets)
        s.set_groups(g["group"], g["u_chip"], g["u_offsets"])
        for it, (theta, state) in enumerate(starts, start=1):
            p = em._rejection_cut(it, ctl)
            # oracle side of this iteration (fresh backend so that the buckets are this iteration's)
            r2 = oracle.RNG.set_seed(0)
            r2.mti = int(state[0])
            for i in range(624):
                r2.mt[i] = int(state[1 + i])
            bo = OracleBackend(counts, offsets, 2, groups=g, rng=r2)
            bo.loglik_consensus(theta["psi"], ctl["minZero"])
            bo.expstep(theta["pi"], theta["gamma"])
            mo, n_o = bo.maxstepprob(), bo.rc_consensus(p)
            po = em.optimNB_all(theta["psi"], bo, ctl)
            # GPU side
            s.rng_set_state(state)
            s.loglik_consensus(theta["psi"], ctl["minZero"])
            s.expstep(theta["pi"], theta["gamma"])
            mg, n_g = s.maxstepprob(), s.rc_consensus(p)
            pg = em.optimNB_all(theta["psi"], s, ctl)
            assert _rel(mg["pi"], mo["pi"]) <= 1e-9 and _rel(mg["gamma"], mo["gamma"]) <= 1e-9, it
            assert _rel(s.qfunction(theta["pi"], theta["gamma"]), bo.qfunction(theta["pi"], theta["gamma"])) <= 1e-9
            assert _rel(s.bic(5, N), bo.bic(5, N)) <= 1e-9
            assert n_g == n_o and np.array_equal(s.rng_get_state(), r2.state()), it
            for k in range(2):
                b = s.rc_buckets(k)
                # the oracle's own posteriors carry the cancellation noise of logFP + logBP - ll
                # (ulp of |ll| ~ 3e4 is 4e-12), so the sums are held to the posterior bar, not to ulps
                assert np.array_equal(b != 0, bo.rc[k] != 0), (it, k)
                assert np.allclose(b, bo.rc[k], rtol=1e-9, atol=0), (it, k, float(np.max(np.abs(b - bo.rc[k]) / np.maximum(bo.rc[k], 1e-300))))
                assert _psi_equivalent(s, k, pg[k]["par"], po[k]["par"]), (it, k, pg[k]["par"], po[k]["par"])
                assert pg[k]["convergence"] == po[k]["convergence"]
                strict += _rel(pg[k]["par"], po[k]["par"]) <= 1e-9
    print("teacher-forced (%s): psi within 1e-9 in %d of %d optimisations" % (data, strict, 2 * len(starts)))
    assert strict >= len(starts)


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


def test_prune_patterns_removes_absent_combinations(ehmm):
    """prunePatterns (R/base-utils.R:398-447; epigraHMM() routes to it when control$pruningThreshold is set,
    R/epigraHMM.R:82-84): three conditions, but the differential windows only ever show 'A alone' or
    'B and C together'.  Every other combinatorial pattern must be pruned, one refit per removal."""
    from epigrahmm_b200 import core, rng, synth
    M, n_cond, n_rep = 30011, 3, 2
    gen = np.random.default_rng(5)
    z = synth.markov_chain(M, synth.GAMMA3, gen)
    pats = synth.enumerate_patterns(n_cond)                  # rows ordered as the reference's table
    keep = [i for i, p in enumerate(pats.tolist()) if p in ([1, 0, 0], [0, 1, 1])]
    assert len(keep) == 2
    comp = np.asarray(keep)[gen.integers(0, 2, M)]
    cond_of = np.repeat(np.arange(n_cond), n_rep)
    enriched = np.where((z == 2)[:, None], 1, np.where((z == 1)[:, None], pats[comp][:, cond_of], 0))
    mu, size = np.where(enriched == 1, 12.0, 2.0), np.where(enriched == 1, 4.0, 3.0)
    offsets = np.asfortranarray(np.round(gen.normal(0.0, 0.1, (M, n_cond * n_rep)), 3))
    counts = np.asfortranarray(gen.poisson(gen.gamma(shape=size, scale=mu / size * np.exp(offsets))).astype(np.int32))
    condition = ["A", "A", "B", "B", "C", "C"]
    rng.set_seed(210423)
    peaks = (enriched == 1).astype(float)                    # initial peak calls (the initializer is tested elsewhere)
    ctl = em.controlEM(maxIterEM=8, pruningThreshold=0.05)
    fit = core.epigraHMM(counts, offsets, peaks, condition, ctl, type="differential", key="/tmp/prune.h5")
    try:
        assert sorted(fit["mixturePatterns"]) == ["BEE", "EBB"]
        assert len(fit["pruned"]) == 4 and [1] not in fit["pruned"] and [2, 3] not in fit["pruned"]
        assert np.all(fit["theta"]["delta"] >= 0.05) and abs(fit["theta"]["delta"].sum() - 1) < 1e-9
        assert fit["design"].shape == (2, 6)
        bic = fit["history"]["controlHist"][-2]["bic"]
        assert np.isfinite(bic) and np.isfinite(fit["fullBIC"])
        assert fit["control"]["pruningThreshold"] == 0.05 and len(fit["control"]["pattern"]) == 2
        v = fit["session"].fetch("viterbi").ravel()
        assert np.mean((v == 1) == (z == 1)) > 0.9
    finally:
        fit["session"].close()
