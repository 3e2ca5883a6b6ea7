"""Host logic of the EM drivers (no GPU): the loops of epigrahmm_b200.em driven over the CPU
oracle backend, plus the small helpers mirrored from R/base-utils.R, R/base-checkers.R,
R/controlEM.R and R/estimateTransitionProb.R."""
import numpy as np
import pytest

from oracle_backend import OracleBackend

from epigrahmm_b200 import em, synth


def test_fdr_control_reference_test():
    """tests/testthat/test-base.R:1-4"""
    assert em.fdrControl(np.arange(0, 1.0005, 0.001), 0.05).sum() == 101
    with pytest.raises(ValueError):
        em.fdrControl(np.array([1.2]), 0.05)
    with pytest.raises(ValueError):
        em.fdrControl(np.array([0.2]), 1.0)


def test_control_em_validation():
    """tests/testthat/test-controlEM.R style: defaults and rejected values"""
    c = em.controlEM()
    assert len(c) == 18 and c["maxIterEM"] == 500 and c["probCut"] == 0.05 and c["minZero"] == em.XMIN
    assert em.controlEM(epsilonEM={"MRCPE": 1e-2})["epsilonEM"] == {"MRCPE": 1e-2, "MACPE": 1e-3, "ARCEL": 1e-3}
    for bad in (dict(maxIterEM=0), dict(gapIterEM=5), dict(probCut=1.0), dict(criterion="x"), dict(pattern=3),
                dict(maxDisp=0), dict(pruningThreshold=1.5), dict(epsilonInnerEM=0)):
        with pytest.raises(ValueError):
            em.controlEM(**bad)


def test_estimate_transition_prob():
    chain = np.array([1, 1, 1, 2, 2, 1, 1, 2, 2, 2])
    P = em.estimateTransitionProb(chain, 2)
    assert np.allclose(P, [[3 / 5, 2 / 5], [1 / 4, 3 / 4]])
    assert np.allclose(em.estimateTransitionProb(np.array([0, 0, 1, 2, 2, 1, 0]), 3).sum(1), 1.0)


def test_get_error_and_convergence_bookkeeping():
    ctl = em.controlEM(minIterEM=3, gapIterEM=3)
    hist = [em.controlList()]
    par = []
    for it in range(1, 6):
        hist[-1]["iteration"] = it
        th = dict(pi=np.array([0.9, 0.1]), gamma=np.eye(2) * 0.8 + 0.1, psi=[np.array([1.0 + 0.1 * it, 2.0]), np.array([3.0, 4.0])])
        em._update_hist(hist, par, th, ctl, 0.0, bic=10.0 * it, q=-100.0 - it, conv=0)
    # iteration 5 > minIterEM: compared with iteration 2 (gap 3)
    e = hist[4]["error"]
    assert np.isclose(e[1], 0.3) and np.isclose(e[0], 0.3 / 1.2) and np.isclose(e[2], 3.0 / 102.0)
    # iteration 2 <= minIterEM: gap 1
    assert np.isclose(hist[1]["error"][1], 0.1)
    assert em.checkConvergence(hist, ctl) == 0


def _callpeaks_data(seed=210423):
    """the shape of tests/testthat/test-callPeaks.R: 1000 background / 2000 enriched / 1000 background"""
    rng = np.random.default_rng(seed)

    def nb(n, mu, size):
        return rng.poisson(rng.gamma(size, mu / size, n))
    counts = np.vstack([nb(2000, 2, 10).reshape(-1, 2), nb(4000, 100, 5).reshape(-1, 2), nb(2000, 2, 10).reshape(-1, 2)])
    truth = np.r_[np.zeros(1000, bool), np.ones(2000, bool), np.zeros(1000, bool)]
    return counts.astype(np.int32), np.zeros(counts.shape), truth


def test_consensus_em_calls_peaks_like_reference_test(oracle):
    """tests/testthat/test-callPeaks.R:1-39: Viterbi and FDR peaks miss <= 5% of the enriched bins"""
    counts, offsets, truth = _callpeaks_data()
    ctl = em.controlEM(maxIterEM=15)
    # per-sample initial peaks (initializerHMM) through the same driver
    peaks = np.zeros(counts.shape)
    for i in range(2):
        c, o = counts[:, [i]], offsets[:, [i]]
        be = OracleBackend(c, o, 2, groups=synth.make_groups(c, o), rng=oracle.RNG.set_seed(1))
        peaks[:, i], _ = em.initializerHMM(c, o, be, em.controlEM(maxIterEM=5))
    assert set(np.unique(peaks)) <= {0.0, 1.0}
    z = 1 + (peaks.sum(1) > 0)
    from epigrahmm_b200 import core
    psi = core.estimateCoefficients(z, counts, offsets, ctl, "consensus")
    theta = dict(pi=np.array([0.999, 0.001]), gamma=em.estimateTransitionProb(z, 2), psi=psi)
    be = OracleBackend(counts, offsets, 2, groups=synth.make_groups(counts, offsets), rng=oracle.RNG.set_seed(2))
    fit = em.emConsensus(theta, be, 2, ctl)
    th = fit["theta_new"]
    hist = fit["controlHist"]
    assert 1 <= hist[-1]["iteration"] <= 15 and len(fit["parHist"]) == hist[-1]["iteration"]
    assert all(np.isfinite(h["bic"]) and np.isfinite(h["q"]) for h in hist[:-1])
    be.viterbi(th["pi"], th["gamma"])
    for method in ("viterbi", 0.05):
        call = em.callPeaks(be, method)
        assert (truth & ~call).sum() / truth.sum() <= 0.05
    # dataset names and shapes the reference's test-epigraHMM.R:29-41 checks
    for n in ("logBP", "logFP", "logLikelihood", "logProb1", "logProb2", "viterbi"):
        a = be.fetch(n)
        assert a.shape[0] == 4000 and np.isfinite(a).all()


def test_differential_em_runs(oracle):
    d = synth.make_differential(3000, 2, 2, seed=4)
    g = synth.make_groups(d["counts"], d["offsets"], design=d["design"])
    be = OracleBackend(d["counts"], d["offsets"], 3, design=d["design"], groups=g, rng=oracle.RNG.set_seed(3))
    th = synth.THETA_DIFFERENTIAL
    theta = dict(pi=th["pi"], gamma=th["gamma"], delta=np.full(d["B"], 1.0 / d["B"]),
                 psi=[th["psi"][:2].copy(), th["psi"][2:].copy()])
    fit = em.outerEMDifferential(theta, be, 4, em.controlEM(maxIterEM=3))
    t = fit["theta_new"]
    assert fit["controlHist"][-1]["iteration"] == 3
    assert np.isclose(np.sum(t["delta"]), 1.0) and np.allclose(np.asarray(t["gamma"]).sum(1), 1.0)
    assert em.unlist(t).size == 3 + 9 + d["B"] + 4
    # the simulated enrichment (mu 12 vs 2) is recovered in sign and rough size
    assert 1.0 < t["psi"][1][0] < 2.5


def test_optim_differential_reference_test(oracle):
    """tests/testthat/test-base.R:83-103: ChIP = 1, offsets = -1, unit weights; the optimum sits
    at par = (1, log maxDisp, 0, > 0)"""
    rng = np.random.default_rng(210427)
    n = 1000
    G = 25
    grp = rng.integers(1, G + 1, n)
    id_ = np.repeat(np.arange(1, 5), 250)
    dsg = np.column_stack([np.repeat([1, 0], 500), np.repeat([0, 1], 500)]).astype(np.uint8)

    class Table:  # glmMixNB over an explicit long table (the reference passes rcWeights directly)
        def glm_mix_nb(self, par, minZero, grad=True):
            return oracle.glmMixNB(par, id_, np.ones(n), np.ones(n), -np.ones(n), dsg, 4, minZero, grad)
    ctl = em.controlEM()
    opt = em.optimDifferential(np.array([1.0, 2.0, 3.0, 4.0]), Table(), ctl)
    assert abs(opt["par"][0] - 1.0) < 1e-4 and abs(opt["par"][1] - np.log(ctl["maxDisp"])) < 1e-4
    assert abs(opt["par"][2]) < 1e-4 and opt["par"][3] > 0
    del grp


def test_zinb_parameter_maps_and_em_on_oracle_backend(oracle):
    """invertDispersion(model='zinb') / optimNB's back-transformation (R/base-utils.R:174-180,
    R/base-optim.R:43-47), estimateCoefficients(dist='zinb') and a short emConsensus(dist='zinb') over the
    oracle backend: the structural zeros are picked up by the zero-inflation intercept."""
    from oracle_backend import OracleBackend
    from epigrahmm_b200 import core, em, synth
    p3, p5 = np.array([-1.0, 0.7, 3.0]), np.array([-1.0, 0.2, 0.7, -0.1, 3.0])
    assert np.allclose(em.invertDispersion(p3, "zinb"), [0.7, 1 / 3.0, -1.0])
    assert np.allclose(em.invertDispersion(p5, "zinb"), [0.7, -0.1, 1 / 3.0, -1.0, 0.2])
    assert np.allclose(em._zinb_from_optim(em.invertDispersion(p3, "zinb")), p3)
    assert np.allclose(em._zinb_from_optim(em.invertDispersion(p5, "zinb")), p5)
    M, N = 6007, 2
    d = synth.make_consensus(M, N, seed=77)
    rng = np.random.default_rng(78)
    counts = d["counts"].copy()
    counts[(rng.random(counts.shape) < 0.3) & (d["z"][:, None] == 0)] = 0
    counts = np.asfortranarray(counts)
    z = 1 + (d["z"] == 1)
    ctl = em.controlEM(maxIterEM=5)
    psi = core.estimateCoefficients(z, counts, d["offsets"], ctl, "consensus", "zinb")
    assert psi[0].size == 3 and psi[1].size == 2 and psi[0][0] <= 0.0       # log(zip), zip in [0.01, 1]
    theta = dict(pi=np.array([0.999, 0.001]), gamma=em.estimateTransitionProb(z, 2), psi=psi)
    be = OracleBackend(counts, d["offsets"], 2, groups=synth.make_groups(counts, d["offsets"]), rng=oracle.RNG.set_seed(5))
    fit = em.emConsensus(theta, be, N, ctl, dist="zinb")
    th = fit["theta_new"]
    assert len(fit["parHist"]) == 5 and th["psi"][0].size == 3
    assert all(np.all(np.isfinite(em.unlist(p))) for p in fit["parHist"])
    zi = 1 / (1 + np.exp(-th["psi"][0][0]))
    assert 0.15 < zi < 0.45, zi
    bics = [h["bic"] for h in fit["controlHist"][:-1]]
    assert bics[-1] < bics[0]


def test_prune_patterns_loop_semantics(monkeypatch):
    """prunePatterns (R/base-utils.R:398-447) with the fit stubbed out: the rarest pattern goes first
    (which.min takes the first of ties), control$pattern lists the survivors as 1-based condition sets,
    the loop ends when every posterior proportion reaches the threshold, pruningThreshold is restored"""
    from epigrahmm_b200 import core, em
    calls = []
    post = {None: [0.40, 0.01, 0.30, 0.01, 0.20, 0.08],          # full model, 3 conditions -> 6 patterns
            5: [0.41, 0.30, 0.01, 0.20, 0.08],
            4: [0.42, 0.30, 0.20, 0.08]}
    full = ["EBB", "BEB", "BBE", "EEB", "EBE", "BEE"]

    class FakeSession:
        closed = 0

        def close(self):
            FakeSession.closed += 1

    def fake_fit(counts, offsets, peaks, condition, control, dist="nb", **kw):
        pat = control.get("pattern")
        calls.append(pat)
        names = full if pat is None else ["".join("E" if c + 1 in p else "B" for c in range(3)) for p in pat]
        key = None if pat is None else len(pat)
        assert control["pruningThreshold"] is None
        return dict(session=FakeSession(), theta=dict(delta=np.array(post[key])), mixturePatterns=names,
                    history=dict(controlHist=[dict(bic=-123.0), dict(bic=-123.0)]))
    monkeypatch.setattr(core, "differentialHMM", fake_fit)
    ctl = em.controlEM(pruningThreshold=0.05)
    fit = core.epigraHMM(None, None, None, ["A", "B", "C"], ctl, type="differential")
    assert calls[0] is None
    assert calls[1] == [[1], [3], [1, 2], [1, 3], [2, 3]]       # 'BEB' (the first of the two 0.01s) removed
    assert calls[2] == [[1], [3], [1, 3], [2, 3]]               # then 'EEB'
    assert len(calls) == 3 and fit["pruned"] == [[2], [1, 2]]
    assert fit["mixturePatterns"] == ["EBB", "BBE", "EBE", "BEE"] and FakeSession.closed == 2
    assert fit["control"]["pruningThreshold"] == 0.05 and ctl["pruningThreshold"] == 0.05
    assert fit["fullBIC"] == -123.0


def test_call_patterns_types():
    """callPatterns (R/callPaterns.R:78-112): 'all', 'fdr', 'max' (which.max: the first of ties) and 'ranges'"""
    from epigrahmm_b200 import em

    class B:
        def fetch(self, name):
            assert name == "mixtureProb"
            return np.array([[0.7, 0.2, 0.1], [0.1, 0.1, 0.8], [0.4, 0.4, 0.2], [0.05, 0.9, 0.05], [0.98, 0.01, 0.01]])
    peaks = np.array([True, False, True, True, True])
    names = ["A", "B", "A-B"]
    idx, all_ = em.callPatterns(B(), peaks, names)
    assert idx.tolist() == [0, 2, 3, 4] and all_.shape == (4, 3)
    idx, mx = em.callPatterns(B(), peaks, names, type="max")
    assert mx.tolist() == ["A", "A", "B", "A"]
    idx, f = em.callPatterns(B(), peaks, names, type="fdr", fdr=0.2, pattern="A")
    assert f.tolist() == em.fdrControl(np.array([0.7, 0.4, 0.05, 0.98]), 0.2).tolist() and f.tolist() == [True, False, False, True]
    idx, r = em.callPatterns(B(), peaks, names, type="ranges", ranges=np.array([False, True, True, False, True]))
    assert idx.tolist() == [2, 4] and r.shape == (2, 3)
    with pytest.raises(ValueError):
        em.callPatterns(B(), peaks, names, type="fdr", fdr=0.2, pattern="C")


def test_sharded_host_helpers_properties():
    """rc_offsets: the ranks' slices of the uniform stream are disjoint, ordered column-major then by block, and
    cover [0, total); combine_carries: forward / backward carries equal the brute-force operator products"""
    from epigrahmm_b200.sharded import combine_carries, rc_offsets
    rng = np.random.default_rng(3)
    for _ in range(50):
        P, Ccols = int(rng.integers(1, 9)), int(rng.integers(1, 33))
        counts = rng.integers(0, 1000, (P, Ccols))
        spans = []
        for r in range(P):
            offs, total = rc_offsets(counts, r)
            assert total == counts.sum()
            spans += [(int(offs[c]), int(offs[c] + counts[r, c]), c, r) for c in range(Ccols)]
        spans.sort()
        pos = 0
        for lo, hi, c, r in spans:
            assert lo == pos
            pos = hi
        assert pos == counts.sum()
        order = [(c, r) for lo, hi, c, r in spans if hi > lo]
        assert order == sorted(order)                      # column-major, then ascending block
    for K in (2, 3):
        for _ in range(20):
            P = int(rng.integers(1, 7))
            ops = np.concatenate([rng.random((P, K * K)) + 0.01, rng.normal(-50, 30, (P, 1))], axis=1)
            start = np.append(rng.dirichlet(np.ones(K)), 0.0)
            for r in range(P):
                fwd, bwd = combine_carries(ops, r, start, K)
                v, ls = start[:K].copy(), 0.0
                for q in range(r):
                    v, ls = v @ ops[q, :K * K].reshape(K, K), ls + ops[q, K * K]
                u, lu = np.ones(K), 0.0
                for q in range(P - 1, r, -1):
                    u, lu = ops[q, :K * K].reshape(K, K) @ u, lu + ops[q, K * K]
                assert np.allclose(np.log(fwd[:K]) + fwd[K], np.log(v) + ls, rtol=0, atol=1e-10)
                assert np.allclose(np.log(bwd[:K]) + bwd[K], np.log(u) + lu, rtol=0, atol=1e-10)
