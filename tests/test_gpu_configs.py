"""The BASELINE.json configurations as parity cases: helas3 (c1) end to end, the shapes of c3/c4/c5
at reduced M under the swapped-backend protocol, and size-independent properties at the full c2 size."""
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


def test_c1_helas3_consensus_fit(ehmm, oracle):
    """BASELINE configs[0]: the bundled helas3 counts (H3K27me3.1/.2, 9 994 bins), consensus model,
    offsets all zero as in the package: the same EM driver over the GPU session and over the
    oracle must give the same trajectory, Viterbi path and peak calls."""
    a = np.load(os.path.join(GOLD, "helas3_counts.npz"))["counts"]
    counts = np.asfortranarray(a[:, 2:4])
    offsets = np.zeros(counts.shape, order="F")
    M, N = counts.shape
    # deterministic initial peaks: the per-sample initializer's quantile split (R/base-core.R:36-38)
    score = np.log1p(counts)
    peaks = (score > np.quantile(score, 0.75, axis=0)).astype(float)
    z = 1 + (peaks.sum(1) > 0)
    from epigrahmm_b200 import core
    ctl = em.controlEM(maxIterEM=12)
    theta = dict(pi=np.array([0.999, 0.001]), gamma=em.estimateTransitionProb(z, 2),
                 psi=core.estimateCoefficients(z, counts, offsets, ctl, "consensus"))
    g = synth.make_groups(counts, offsets)
    be = OracleBackend(counts, offsets, 2, groups=g, rng=oracle.RNG.set_seed(210423))
    fo = em.emConsensus(theta, be, N, ctl)
    with ehmm.Session("helas3", M, 2) as s:
        s.set_data(counts, offsets)
        assert s.build_groups() == g["G"]
        s.rng_set_state(oracle.RNG.set_seed(210423).state())
        fg = em.emConsensus(theta, s, N, ctl)
        assert len(fg["parHist"]) == len(fo["parHist"])
        # free-running, 1e-9 at every iteration (tests/test_gpu_em.py also replays this trajectory iteration by iteration)
        worst, tight = 0.0, 0
        for it, (pg, po) in enumerate(zip(fg["parHist"], fo["parHist"])):
            r = _rel(em.unlist(pg), em.unlist(po))
            worst, tight = max(worst, r), tight + (r <= 1e-9)
            assert r <= 1e-9, (it + 1, r)
            assert _rel(fg["controlHist"][it]["bic"], fo["controlHist"][it]["bic"]) <= max(1e-9, 10 * worst)
            assert _rel(fg["controlHist"][it]["q"], fo["controlHist"][it]["q"]) <= max(1e-9, 10 * worst)
        assert tight >= len(fg["parHist"]) // 2
        tg, to = fg["theta_new"], fo["theta_new"]
        vg, vo = s.viterbi(tg["pi"], tg["gamma"]), be.viterbi(to["pi"], to["gamma"])
        assert np.mean(vg != vo) <= (0 if worst <= 1e-9 else 1e-3)
        assert np.mean(em.callPeaks(s, 0.05) != em.callPeaks(be, 0.05)) <= (0 if worst <= 1e-9 else 1e-3)
        assert np.max(np.abs(np.exp(s.fetch("logProb1")) - np.exp(be.fetch("logProb1")))) <= max(1e-8, 10 * worst)
        assert 0.02 < em.callPeaks(s, "viterbi").mean() < 0.8


@pytest.mark.parametrize("n_cond,n_rep,M", [(3, 2, 6007), (5, 2, 3001), (4, 3, 3001)])
def test_differential_shapes_of_c3_c4_c5(ehmm, oracle, n_cond, n_rep, M):
    """configs[2..4] at reduced M: B = 6 / 30 / 14 mixture components, N = 6 / 10 / 12 samples"""
    d = synth.make_differential(M, n_cond, n_rep, seed=20261000 + n_cond)
    g = synth.make_groups(d["counts"], d["offsets"], design=d["design"])
    N, B = n_cond * n_rep, d["B"]
    assert B == 2 ** n_cond - 2
    th = synth.THETA_DIFFERENTIAL
    theta = dict(pi=th["pi"], gamma=th["gamma"], delta=np.full(B, 1.0 / B), psi=[th["psi"][:2].copy(), th["psi"][2:].copy()])
    ctl = em.controlEM(maxIterEM=2)
    rng = oracle.RNG.set_seed(5)
    be = OracleBackend(d["counts"], d["offsets"], 3, design=d["design"], groups=g, rng=rng)
    fo = em.outerEMDifferential(theta, be, N, ctl)
    with ehmm.Session("shape", M, 3) as s:
        s.set_data(d["counts"], d["offsets"])
        s.set_design(d["design"])
        assert s.build_groups() == g["G"]
        s.rng_set_state(oracle.RNG.set_seed(5).state())
        fg = em.outerEMDifferential(theta, s, N, ctl)
        worst = max(_rel(em.unlist(pg), em.unlist(po)) for pg, po in zip(fg["parHist"], fo["parHist"]))
        assert worst <= 1e-9
        if worst <= 1e-9:
            assert np.array_equal(s.rng_get_state(), rng.state())
        assert np.max(np.abs(s.fetch("mixtureProb") - be.fetch("mixtureProb"))) <= max(1e-9, 10 * worst)
        tg, to = fg["theta_new"], fo["theta_new"]
        assert np.mean(s.viterbi(tg["pi"], tg["gamma"]) != be.viterbi(to["pi"], to["gamma"])) <= (0 if worst <= 1e-9 else 1e-3)


def test_c2_full_size_properties(ehmm):
    """BASELINE configs[1] at its full size (15 000 000 bins x 4 samples): properties that do not
    need the oracle -- normalisation, consistency of the fused statistics with the datasets, the
    two emission paths, agreement of a two-block run with the one-block run, Viterbi sanity."""
    from epigrahmm_b200.sharded import combine_carries
    M, N = 15_000_000, 4
    d = synth.make_consensus(M, N, seed=20261002)
    th = synth.THETA_CONSENSUS
    with ehmm.Session("c2", M, 2) as s:
        s.set_data(d["counts"], d["offsets"])
        s.loglik_consensus(th["psi"], MINZERO)                 # per-cell path
        direct = s.fetch("logLikelihood")
        G = s.build_groups()
        assert 1000 < G < 200000
        s.loglik_consensus(th["psi"], MINZERO)                 # dictionary-encoded path
        logf = s.fetch("logLikelihood")
        assert np.array_equal(direct, logf)
        del direct
        s.expstep(th["pi"], th["gamma"])
        L1 = s.fetch("logProb1")
        assert np.all(np.isfinite(L1)) and np.all(L1 <= 0)
        assert np.max(np.abs(np.exp(L1).sum(1) - 1.0)) <= 1e-12
        st = s.estep_stats()
        assert abs(st[:4].sum() - (M - 1)) <= 1e-6 * M        # joint posteriors of bins 1..M-1 sum to 1 each
        assert abs(st[4] - np.sum(np.exp(L1) * logf)) <= 1e-9 * abs(st[4])
        FP, BP = s.fetch("logFP"), s.fetch("logBP")
        ll = s.loglik()
        # logProb1 = logFP + logBP - ll up to the cancellation noise of 1e8-magnitude sums
        resid = (FP + BP - ll) - L1
        assert np.max(np.abs(resid[L1 > -20])) <= 1e-6
        assert abs(np.logaddexp(FP[-1, 0], FP[-1, 1]) - ll) <= 1e-9 * abs(ll)
        L2 = s.fetch("logProb2")[1:]
        assert np.max(np.abs(np.exp(L2).sum(1) - 1.0)) <= 1e-10   # tile-local log scales are O(1e4): ulp ~ 2e-12
        # marginalising the joint over the previous state gives the marginal posterior
        p_from_joint = np.exp(L2[:, 0]) + np.exp(L2[:, 2])
        assert np.max(np.abs(p_from_joint - np.exp(L1[1:, 0]))) <= 1e-9
        del L2, BP, FP
        m = s.maxstepprob()
        assert np.allclose(np.asarray(m["gamma"]).sum(1), 1.0) and abs(m["pi"].sum() - 1) < 1e-12
        # the same chain as two bin blocks with exchanged carries
        cut = 7_000_001
        blocks = []
        for r, (lo, hi) in enumerate(((0, cut), (cut, M))):
            b = ehmm.Session("c2b%d" % r, hi - lo, 2)
            b.set_block(lo, M)
            blocks.append((b, lo, hi))
        try:
            ops = np.stack([b.expstep_reduce(th["pi"], th["gamma"], logf[lo:hi]) for b, lo, hi in blocks])
            for r, (b, lo, hi) in enumerate(blocks):
                fwd, bwd = combine_carries(ops, r, np.append(th["pi"], 0.0), 2)
                b.expstep_apply(fwd, bwd)
                assert np.max(np.abs(np.exp(b.fetch("logProb1")) - np.exp(L1[lo:hi]))) <= 1e-10
        finally:
            for b, _, _ in blocks:
                b.close()
        # rejection control: uniform consumption equals the flagged count, buckets are sane
        from epigrahmm_b200.rng import RState
        s.rng_set_state(RState(7).state.copy())
        n = s.rc_consensus(0.05)
        x = np.exp(L1)
        assert n == int(np.count_nonzero((x < 0.05) & (x > 0)))
        for k in range(2):
            b = s.rc_buckets(k)
            assert b[0] == 0 and np.all(b >= 0)
        # thinning preserves the expected weight: bucket totals stay within sampling error
        tot = s.rc_buckets(1).sum()
        assert abs(tot - x[:, 1].sum()) <= 6 * np.sqrt(0.05 * x[:, 1].sum()) + 1
        # Viterbi: valid states, close to the posterior mode and to the simulated truth
        v = s.viterbi(th["pi"], th["gamma"])
        assert set(np.unique(v)) <= {0.0, 1.0}
        assert np.mean((v == 1) == (x[:, 1] > 0.5)) > 0.98
        assert np.mean((v == 1) == (d["z"] == 1)) > 0.95


@pytest.mark.parametrize("M,n_cond,n_rep", [(15_000_000, 3, 2), (6_000_000, 5, 2)], ids=["c3", "c4"])
def test_c3_c4_full_size_properties(ehmm, M, n_cond, n_rep):
    """BASELINE configs[2] and [3] at their full sizes (15 000 000 bins x 6 samples, B = 6; 6 000 000 bins x
    10 samples, B = 30; K = 3): size-independent properties of the differential path -- the two emission
    sources agree, posteriors and mixture posteriors are normalised, the fused inner M-step equals the
    sums over the datasets, the rejection control consumes exactly one uniform per flagged (bin, column),
    Viterbi follows the truth."""
    d = synth.make_differential(M, n_cond, n_rep, seed=20261000 + n_cond)
    B = d["B"]
    th = synth.THETA_DIFFERENTIAL
    delta = np.full(B, 1.0 / B)
    psi = np.concatenate([th["psi"][:2], th["psi"][2:]])
    with ehmm.Session("c3", M, 3) as s:
        s.set_data(d["counts"], d["offsets"])
        s.set_design(d["design"])
        s.loglik_differential_hmm(delta, psi, MINZERO)        # per-cell source
        direct = s.fetch("logLikelihood")
        G = s.build_groups()
        assert 1000 < G < 2_000_000
        s.loglik_differential_hmm(delta, psi, MINZERO)        # dictionary-encoded source
        logf = s.fetch("logLikelihood")
        assert np.max(np.abs(direct - logf)) <= 1e-9 * np.max(np.abs(logf))
        del direct
        s.expstep(th["pi"], th["gamma"])
        L1 = s.fetch("logProb1")
        assert np.all(np.isfinite(L1)) and np.all(L1 <= 0)
        P1 = np.exp(L1)
        assert np.max(np.abs(P1.sum(1) - 1.0)) <= 1e-12
        st = s.estep_stats()
        assert abs(st[:9].sum() - (M - 1)) <= 1e-6 * M
        assert abs(st[9] - np.sum(P1 * logf)) <= 1e-9 * abs(st[9])
        del logf, L1
        m = s.maxstepprob()
        assert np.allclose(np.asarray(m["gamma"]).sum(1), 1.0) and abs(m["pi"].sum() - 1) < 1e-12
        s.inner_expstep(delta, psi, MINZERO)
        eta = s.fetch("mixtureProb")
        assert eta.shape == (M, B) and np.max(np.abs(eta.sum(1) - 1.0)) <= 1e-12
        dl = s.inner_maxstepprob()
        want = (P1[:, 1:2] * eta).sum(0) / P1[:, 1].sum()
        assert abs(dl.sum() - 1) <= 1e-12 and np.max(np.abs(dl - want)) <= 1e-9
        # the differential windows were simulated with uniform patterns: the mixing proportions are near 1/B
        assert np.max(np.abs(dl - 1.0 / B)) < 0.02
        from epigrahmm_b200.rng import RState
        s.rng_set_state(RState(11).state.copy())
        n = s.rc_differential(0.05)
        cols = [P1[:, 0]] + [P1[:, 1] * eta[:, b] for b in range(B)] + [P1[:, 2]]
        flagged = sum(int(np.count_nonzero((x < 0.05) & (x / 0.05 != 0))) for x in cols)
        assert n == flagged
        b0 = s.rc_buckets(0)
        assert b0[0] == 0 and np.all(b0 >= 0)
        N = n_cond * n_rep
        tot = s.rc_buckets(B + 1).sum()                      # thinning preserves the expected weight
        assert abs(tot - N * P1[:, 2].sum()) <= N * (6 * np.sqrt(0.05 * P1[:, 2].sum()) + 1)
        del eta, cols
        v = s.viterbi(th["pi"], th["gamma"])
        assert set(np.unique(v)) <= {0.0, 1.0, 2.0}
        assert np.mean(v == d["z"]) > 0.9
