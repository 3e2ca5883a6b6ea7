"""Pins the CPU oracle: known answers the reference's own tests state, independent library
implementations of the third-party arithmetic, a numpy twin of the native recursions, and the
committed golden vectors."""
import os
import sys

import numpy as np
import pytest
from scipy.special import digamma as sp_digamma, logsumexp
from scipy.stats import nbinom

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MINZERO = 2.2250738585072014e-308


def test_dnbinom_matches_scipy(oracle):
    rng = np.random.default_rng(1)
    x = rng.integers(0, 300, 4000).astype(float)
    size = np.exp(rng.normal(1, 2, 4000))
    mu = np.exp(rng.normal(1, 2, 4000))
    a = oracle.dnbinom(x, size, mu, log=True)
    b = nbinom.logpmf(x, size, size / (size + mu))
    assert np.max(np.abs(a - b) / np.maximum(1, np.abs(b))) < 1e-11  # scipy's gammaln differences lose digits
    # 50-digit truth on a subsample: the saddle-point algorithm stays at a few ulp
    import mpmath as mp
    mp.mp.dps = 50
    for i in range(0, 4000, 20):
        xx, ss, mm = mp.mpf(x[i]), mp.mpf(size[i]), mp.mpf(mu[i])
        t = mp.loggamma(xx + ss) - mp.loggamma(ss) - mp.loggamma(xx + 1) + ss * mp.log(ss / (ss + mm)) + xx * mp.log(mm / (ss + mm))
        assert abs(a[i] - float(t)) <= 2e-14 * max(1.0, abs(float(t))), (x[i], size[i], mu[i])
    assert np.allclose(oracle.dnbinom(x, size, mu), np.exp(a), rtol=1e-12)
    # branches of dnbinom_mu: x == 0, x < 1e-10*size, size >> mu, mu >> size
    for xx, ss, mm in [(0, 3.0, 2.0), (0, 2.0, 30.0), (1, 1e12, 2.0), (5, 1000.0, 0.01), (7, 0.01, 500.0)]:
        X, S, Mu = mp.mpf(xx), mp.mpf(ss), mp.mpf(mm)
        ref = float(mp.loggamma(X + S) - mp.loggamma(S) - mp.loggamma(X + 1) + S * mp.log(S / (S + Mu)) + X * mp.log(Mu / (S + Mu)))
        assert abs(oracle.dnbinom(xx, ss, mm, log=True) - ref) < 1e-11 * max(1, abs(ref))  # the x < 1e-10*size branch is O(mu^2/size) approximate in R itself


def test_reference_kat_loglik_differential_hmm(oracle):
    """tests/testthat/test-base.R:6-42: three unique rows equal to closed-form dnbinom sums"""
    M = 1000
    counts = np.ones((M, 2), dtype=np.int32)
    offsets = -np.ones((M, 2))
    design = np.array([[1, 0], [0, 1]], dtype=np.uint8)
    ll = oracle.loglikDifferentialHMM(counts, offsets, design, [0.1, 0.9], [0.0, 1.0, 0.0, 1.0], 1e-32)
    rows = np.unique(ll, axis=0)
    assert rows.shape == (1, 3)
    e = np.e
    d1 = nbinom.logpmf(1, e, e / (e + np.exp(-1.0)))
    d2 = nbinom.logpmf(1, e * e, e * e / (e * e + np.exp(-1.0)))
    assert np.allclose(rows[0], [2 * d1, d1 + d2, 2 * d2], rtol=0, atol=1.5e-8)  # expect_equal tolerance


def test_digamma(oracle):
    x = np.exp(np.random.default_rng(2).uniform(-4, 8, 3000))
    assert np.max(np.abs(oracle.digamma(x) - sp_digamma(x)) / np.maximum(1, np.abs(sp_digamma(x)))) < 1e-14


def test_r_rng_known_answers(oracle):
    """values printed by R: set.seed(1); runif(3)  and  set.seed(42); runif(2)"""
    assert np.allclose(oracle.RNG.set_seed(1).runif(3), [0.2655087, 0.3721239, 0.5728534], atol=5e-8)
    assert np.allclose(oracle.RNG.set_seed(42).runif(2), [0.9148060, 0.9370754], atol=5e-8)
    g = oracle.RNG.set_seed(123)  # set.seed(123); runif(1) = 0.2875775
    assert abs(g.unif_rand() - 0.2875775) < 5e-8


def test_product_rng_matches_oracle(oracle):
    from epigrahmm_b200.rng import RState
    a, b = RState(210423), oracle.RNG.set_seed(210423)
    assert np.array_equal(a.runif(5000), b.runif(5000))
    assert np.array_equal(a.state, b.state())


def _twin_expstep(pi, gamma, logf):
    """textbook log-space forward-backward written independently of the oracle"""
    M, K = logf.shape
    lg = np.log(gamma)
    F = np.empty((M, K))
    B = np.zeros((M, K))
    F[0] = np.log(pi) + logf[0]
    for j in range(1, M):
        F[j] = logsumexp(F[j - 1][:, None] + lg, axis=0) + logf[j]
    for j in range(M - 2, -1, -1):
        B[j] = logsumexp(lg + (logf[j + 1] + B[j + 1])[None, :], axis=1)
    ll = logsumexp(F[-1])
    L1 = F + B - ll
    L1 -= logsumexp(L1, axis=1, keepdims=True)
    L2 = np.zeros((M, K * K))
    for j in range(1, M):
        L2[j] = (F[j - 1][:, None] + lg + (logf[j] + B[j])[None, :] - ll).ravel()
    L2 -= logsumexp(L2, axis=1, keepdims=True)
    return F, B, L1, L2, ll


@pytest.mark.parametrize("K", [2, 3])
def test_expstep_against_numpy_twin(oracle, K):
    from epigrahmm_b200 import synth
    M = 3000
    rng = np.random.default_rng(K)
    logf = rng.normal(-5, 2, (M, K))
    th = synth.THETA_CONSENSUS if K == 2 else synth.THETA_DIFFERENTIAL
    h = oracle.expStep(th["pi"], th["gamma"], logf)
    F, B, L1, L2, ll = _twin_expstep(th["pi"], th["gamma"], logf)
    assert np.allclose(h["logFP"], F, rtol=1e-12) and np.allclose(h["logBP"], B, rtol=1e-12, atol=1e-12)
    assert np.allclose(h["logProb1"], L1, atol=1e-9) and np.allclose(h["logProb2"], L2, atol=1e-9)
    # maxStepProb / Q / BIC from their definitions
    P2 = np.exp(L2[1:])
    gam = (P2.sum(0).reshape(K, K).T / P2.sum(0).reshape(K, K).sum(1)).T
    gam = np.maximum(gam, 1e-6)
    gam /= gam.sum(1, keepdims=True)
    m = oracle.maxStepProb(h)
    assert np.allclose(m["gamma"], gam, rtol=1e-9)
    p0 = np.maximum(np.exp(L1[0]), 1e-6)
    assert np.allclose(m["pi"], p0 / p0.sum(), rtol=1e-9)
    q = np.exp(L1[0]) @ np.log(th["pi"]) + np.sum(np.exp(L1) * logf) + np.sum(P2 * np.log(th["gamma"]).ravel())
    assert abs(oracle.computeQFunction(h, th["pi"], th["gamma"]) - q) < 1e-9 * abs(q)
    assert abs(oracle.computeBIC(h, 6, 4) - (-2 * ll + 6 * np.log(4 * M))) < 1e-9 * abs(ll)
    # Viterbi: the reference's recursion scores bins with logFP (src/computeViterbiSequence.cpp:20)
    V = np.log(th["pi"]) + F[0]
    lg = np.log(th["gamma"])
    psi = np.zeros((M, K), dtype=int)
    for j in range(1, M):
        c = V[:, None] + lg
        psi[j] = c.argmax(0)
        V = c.max(0) + F[j]
    S = np.empty(M)
    S[-1] = V.argmax()
    for j in range(M - 2, -1, -1):
        S[j] = psi[j + 1][int(S[j + 1])]
    assert np.array_equal(oracle.computeViterbiSequence(h, th["pi"], th["gamma"]), S)


def test_reweight_aggregate_semantics(oracle):
    x = np.array([0.0, 0.5, 0.01, 0.2, 0.9, 1e-320])
    u = np.array([0.5, 0.1, 0.3])  # consumed by 0.01, 0.2 and the subnormal, in that order; 0.0 draws nothing
    w, n = oracle.reweight(x, 0.3, uniforms=u)
    assert n == 3
    # rbinom(1, q): q <= .5 -> [u >= 1-q];  q > .5 -> 1 - [u >= 1-(1-q)]
    assert np.array_equal(w, [0.0, 0.5, 0.0, 0.3, 0.9, 0.0])
    w3, _ = oracle.reweight(x, 0.3, uniforms=np.array([0.97, 0.999, 0.3]))
    assert np.array_equal(w3, [0.0, 0.5, 0.3, 0.0, 0.9, 0.0])
    w2, n2 = oracle.reweight(x, 0.0, uniforms=u)
    assert n2 == 0 and np.array_equal(w2, x)
    # aggregate: only non-zero sums, ascending id, only the first len(x) ids are read
    a = oracle.aggregate(np.array([1.0, 0.0, 2.0, 0.5]), np.array([3, 1, 3, 7, 9, 9]))
    assert np.array_equal(a, [[3.0, 3.0], [0.5, 7.0]])


def test_rbinom_probability(oracle):
    rng = oracle.RNG.set_seed(99)
    x = np.full(200000, 0.06)
    w, n = oracle.reweight(x, 0.3, rng=rng)
    assert n == x.size and abs((w > 0).mean() - 0.2) < 0.005 and set(np.unique(w)) == {0.0, 0.3}


def test_oracle_reproduces_golden(oracle):
    from epigrahmm_b200 import synth
    g = np.load(os.path.join(GOLD, "golden_small.npz"))
    th = synth.THETA_CONSENSUS
    logf = oracle.loglikConsensusHMM(g["c_counts"], g["c_offsets"], th["psi"], MINZERO)
    assert np.array_equal(logf, g["c_logf"])
    h = oracle.expStep(th["pi"], th["gamma"], logf)
    for n in ("logFP", "logBP", "logProb1", "logProb2"):
        assert np.array_equal(h[n], g["c_" + n])
    assert np.array_equal(oracle.computeViterbiSequence(h, th["pi"], th["gamma"]), g["c_viterbi"])
    rc, n = oracle.consensusRejectionControlled(h, g["c_group"], 0.05, uniforms=g["c_unif"], dense=True)
    assert n == g["c_rc_n"] and np.array_equal(rc, g["c_rc"])
    th = synth.THETA_DIFFERENTIAL
    delta = np.full(6, 1 / 6)
    logf = oracle.loglikDifferentialHMM(g["d_counts"], g["d_offsets"], g["d_design"], delta, th["psi"], MINZERO)
    assert np.array_equal(logf, g["d_logf"])
    assert np.array_equal(oracle.innerExpStep(g["d_counts"], g["d_offsets"], g["d_design"], delta, th["psi"], MINZERO),
                          g["d_eta"])


def test_helas3_fixture():
    a = np.load(os.path.join(GOLD, "helas3_counts.npz"))["counts"]
    assert a.shape == (9994, 6) and a.dtype == np.int32
    assert a.sum(0).tolist() == [104862, 75257, 86302, 59193, 91121, 85604] and a.max() == 191


def _r_ifelse(test, yes, no):
    """R's ifelse(): the result has the SHAPE OF `test`; `yes` / `no` are recycled (or truncated) to it.
    With a length-1 test only yes[1] or no[1] is returned -- base R semantics, ?ifelse: "ifelse returns
    a value with the same shape as test"."""
    test = np.atleast_1d(np.asarray(test, dtype=bool))
    yes, no = np.atleast_1d(yes), np.atleast_1d(no)
    return np.where(test, np.resize(yes, test.shape), np.resize(no, test.shape))


def test_consensus_controls_follow_r_ifelse_semantics(oracle):
    """loglikConsensusHMM with a controls assay, R/base-loglikelihood.R:110-122 transcribed line by line
    with R's ifelse() semantics and scipy's nbinom: `ifelse(useControl, theta[2]*controls, 0)` has a
    length-1 test, so the control of the table's first cell is used for every cell.  Checks the oracle
    against that transcription (not against itself) and shows the per-cell reading differs."""
    from scipy.stats import nbinom
    rng = np.random.default_rng(5)
    M, N = 211, 3
    y = rng.negative_binomial(3, 0.3, (M, N)).astype(np.int32)
    off = np.round(rng.normal(0, 0.1, (M, N)), 3)
    ctrl = np.log1p(rng.poisson(4.0, (M, N)).astype(float))
    mz = 2.2250738585072014e-308
    th = [np.array([np.log(2.0), 0.3, 3.0]), np.array([np.log(12.0), 0.1, 4.0])]
    assert _r_ifelse(True, np.array([7.0, 8.0, 9.0]), 0).tolist() == [7.0]        # R: ifelse(TRUE, c(7,8,9), 0) is 7
    ChIP, offsets, controls = (a.ravel(order="F") for a in (y.astype(float), off, ctrl))  # dt's long columns
    useControl = True
    want = np.empty((M, 2))
    for k in range(2):
        mu = np.exp(th[k][0] + _r_ifelse(useControl, th[k][1] * controls, 0) + offsets)
        size = _r_ifelse(useControl, th[k][2], th[k][1])
        cell = np.log(nbinom.pmf(ChIP, size, size / (size + mu)) + mz)
        want[:, k] = cell.reshape(M, N, order="F").sum(1)   # rowSums(matrix(..., nrow = M, ncol = N, byrow = FALSE))
    got = oracle.loglikConsensusHMM(np.asfortranarray(y), np.asfortranarray(off), th, mz, np.asfortranarray(ctrl))
    assert np.max(np.abs(got - want)) <= 1e-11
    # hand-computed cell: bin 2 of sample 2 gets the control of bin 1 of sample 1
    mu = np.exp(th[0][0] + th[0][1] * ctrl[0, 0] + off[1, 1])
    p = th[0][2] / (th[0][2] + mu)
    from math import lgamma, log
    lp = lgamma(y[1, 1] + 3.0) - lgamma(3.0) - lgamma(y[1, 1] + 1.0) + 3.0 * log(p) + y[1, 1] * log(1 - p)
    one = oracle.loglikConsensusHMM(np.asfortranarray(y[1:2, 1:2]), np.asfortranarray(off[1:2, 1:2]), th, mz,
                                    np.asfortranarray(ctrl[0:1, 0:1]))
    assert abs(one[0, 0] - lp) <= 1e-12 * abs(lp)
    # the "natural" per-cell reading is a different function
    percell = np.log(nbinom.pmf(y, 3.0, 3.0 / (3.0 + np.exp(th[0][0] + th[0][1] * ctrl + off))) + mz).sum(1)
    assert np.max(np.abs(percell - want[:, 0])) > 1e-3


# ---- dist = 'zinb'  ---------------------------------------------------------

def test_zinb_emission_against_scipy_twin(oracle):
    """R/base-loglikelihood.R:123-141 written out with scipy's nbinom (independent arithmetic)"""
    from scipy.stats import nbinom
    rng = np.random.default_rng(11)
    M, N = 300, 3
    y = rng.negative_binomial(3, 0.4, (M, N)).astype(np.int32)
    y[rng.random((M, N)) < 0.3] = 0
    off = np.round(rng.normal(0, 0.1, (M, N)), 3)
    ctrl = np.log1p(rng.poisson(3.0, (M, N)).astype(float))
    mz = 2.2250738585072014e-308

    def dnb(x, mu, size):
        return nbinom.pmf(x, size, size / (size + mu))
    for useC in (False, True):
        psi = [np.array([-1.0, 0.2, 0.7, -0.1, 3.0]), np.array([2.0, 0.05, 4.0])] if useC else \
            [np.array([-1.0, 0.7, 3.0]), np.array([2.0, 4.0])]
        LL = oracle.loglikConsensusHMMzinb(np.asfortranarray(y), np.asfortranarray(off), psi, mz,
                                           np.asfortranarray(ctrl) if useC else None)
        if useC:   # l.126-129,137 word for word, with R's ifelse (see _r_ifelse)
            cv = ctrl.ravel(order="F")
            xb = _r_ifelse(useC, psi[0][0] + psi[0][1] * cv, psi[0][0]) + off
            mu1, s1 = np.exp(_r_ifelse(useC, psi[0][2] + psi[0][3] * cv, psi[0][1]) + off), psi[0][4]
            mu2, s2 = np.exp(_r_ifelse(useC, psi[1][0] + psi[1][1] * cv, psi[1][0]) + off), psi[1][2]
        else:
            xb, mu1, s1 = psi[0][0] + off, np.exp(psi[0][1] + off), psi[0][2]
            mu2, s2 = np.exp(psi[1][0] + off), psi[1][1]
        zp = 1 / (1 + np.exp(-xb))
        cell = np.where(y == 0, np.log(zp + (1 - zp) * (dnb(0, mu1, s1) + mz)), np.log(1 - zp) + np.log(dnb(y, mu1, s1) + mz))
        assert np.max(np.abs(LL[:, 0] - cell.sum(1))) <= 1e-11
        assert np.max(np.abs(LL[:, 1] - np.log(dnb(y, mu2, s2) + mz).sum(1))) <= 1e-11


def test_glm_zinb_gradient_and_nb_limit(oracle):
    """derivZINB is the gradient of glmZINB (central differences); with lambda -> -inf the value is glmNB's
    (up to the minZero floor, which both apply inside the log)"""
    rng = np.random.default_rng(12)
    G = 400
    y = rng.negative_binomial(2, 0.3, G).astype(float)
    y[rng.random(G) < 0.3] = 0
    o, w = np.round(rng.normal(0, 0.1, G), 3), rng.random(G)
    ctrl = np.log1p(rng.poisson(3.0, G).astype(float))
    mz = 2.2250738585072014e-308
    for x, par in ((np.ones((G, 1)), np.array([0.8, 0.4, -0.5])),
                   (np.column_stack([np.ones(G), ctrl]), np.array([0.8, -0.1, 0.4, -0.5, 0.2]))):
        v, g = oracle.glmZINB(par, y, x, o, w, mz)
        num = np.zeros(par.size)
        for i in range(par.size):
            h = 1e-6
            a, b = par.copy(), par.copy()
            a[i] += h
            b[i] -= h
            num[i] = (oracle.glmZINB(a, y, x, o, w, mz, grad=False)[0] - oracle.glmZINB(b, y, x, o, w, mz, grad=False)[0]) / (2 * h)
        assert np.max(np.abs(g - num)) <= 1e-5 * max(1.0, np.max(np.abs(num)))
        P = x.shape[1]
        nozi = par.copy()
        nozi[P + 1] = -40.0
        nozi[P + 2:] = 0.0
        vz = oracle.glmZINB(nozi, y, x, o, w, mz, grad=False)[0]
        vn = oracle.glmNB(par[:P + 1], y, x, o, w, grad=False)[0]
        assert abs(vz - vn) <= 1e-9 * abs(vn)


def test_differential_zinb_oracle_against_twin_and_differences(oracle):
    """loglikDifferential(HMM) with dist='zinb' against a scipy twin; derivMixZINB against central differences"""
    from scipy.stats import nbinom
    rng = np.random.default_rng(21)
    M, N, B = 200, 4, 2
    design = np.array([[1, 1, 0, 0], [0, 0, 1, 1]], dtype=np.uint8)
    y = rng.negative_binomial(3, 0.4, (M, N)).astype(np.int32)
    y[rng.random((M, N)) < 0.3] = 0
    off = np.round(rng.normal(0, 0.1, (M, N)), 3)
    psi = np.array([-1.0, 0.7, 1.1, 1.3, 0.3])
    delta = np.array([0.4, 0.6])
    mz = 2.2250738585072014e-308
    zp = np.exp(psi[0]) / (1 + np.exp(psi[0]))

    def dnb(x, mu, size):
        return nbinom.pmf(x, size, size / (size + mu))
    lz = np.log(zp * (y == 0) + (1 - zp) * (dnb(y, np.exp(psi[1] + off), np.exp(psi[2])) + mz))
    ln = nbinom.logpmf(y, np.exp(psi[2] + psi[4]), np.exp(psi[2] + psi[4]) / (np.exp(psi[2] + psi[4]) + np.exp(psi[1] + psi[3] + off)))
    ll = np.stack([np.log(delta[b]) + np.where(design[b][None, :] == 1, ln, lz).sum(1) for b in range(B)], axis=1)
    got = oracle.loglikDifferentialZINB(np.asfortranarray(y), np.asfortranarray(off), design, delta, psi, mz)
    assert np.max(np.abs(got - ll)) <= 1e-10
    LL = oracle.loglikDifferentialHMMzinb(np.asfortranarray(y), np.asfortranarray(off), design, delta, psi, mz)
    assert np.max(np.abs(LL[:, 0] - lz.sum(1))) <= 1e-10 and np.max(np.abs(LL[:, 2] - ln.sum(1))) <= 1e-10
    assert np.max(np.abs(LL[:, 1] - np.log(np.exp(ll).sum(1)))) <= 1e-10
    eta = oracle.innerExpStepZINB(np.asfortranarray(y), np.asfortranarray(off), design, delta, psi, mz)
    assert np.max(np.abs(eta - np.exp(ll) / np.exp(ll).sum(1, keepdims=True))) <= 1e-12
    # objective / gradient over a long table
    n = 300
    idv = rng.integers(1, B + 3, n).astype(np.int32)
    yy = rng.negative_binomial(2, 0.3, n).astype(float)
    yy[rng.random(n) < 0.3] = 0
    o, w = np.round(rng.normal(0, 0.1, n), 3), rng.random(n)
    dsg = rng.integers(0, 2, (n, B)).astype(np.uint8)
    par = np.array([-0.7, 0.6, 1.0, 2.2, 1.3])
    v, g = oracle.glmMixZINB(par, idv, w, yy, o, dsg, B + 2, mz)
    num = np.zeros(5)
    for i in range(5):
        a, b = par.copy(), par.copy()
        a[i] += 1e-6
        b[i] -= 1e-6
        num[i] = (oracle.glmMixZINB(a, idv, w, yy, o, dsg, B + 2, mz, grad=False)[0] -
                  oracle.glmMixZINB(b, idv, w, yy, o, dsg, B + 2, mz, grad=False)[0]) / 2e-6
    assert np.max(np.abs(g - num)) <= 1e-5 * max(1.0, np.max(np.abs(num)))


def test_oracle_reproduces_zinb_golden(oracle):
    """tests/golden/golden_zinb.npz (make_golden.py::zinb) pins the oracle's dist = 'zinb' functions"""
    sys.path.insert(0, GOLD)
    import make_golden as mg
    g = np.load(os.path.join(GOLD, "golden_zinb.npz"))
    assert np.array_equal(oracle.loglikConsensusHMMzinb(g["c_counts"], g["c_offsets"], mg.PSI_ZC, MINZERO), g["c_logf"])
    v, gr = oracle.glmZINB(mg.PAR_GLM, g["c_u_chip"], np.ones((g["c_u_chip"].size, 1)), g["c_u_offsets"], g["c_w"], MINZERO)
    assert v == float(g["c_glm_v"]) and np.array_equal(gr, g["c_glm_g"])
    LL = oracle.loglikDifferentialHMMzinb(g["d_counts"], g["d_offsets"], g["d_design"], g["d_delta"], mg.PSI_ZD, MINZERO)
    assert np.array_equal(LL, g["d_logf"])
    eta = oracle.innerExpStepZINB(g["d_counts"], g["d_offsets"], g["d_design"], g["d_delta"], mg.PSI_ZD, MINZERO)
    assert np.array_equal(eta, g["d_eta"])
    B = g["d_design"].shape[0]
    mv, mgr = oracle.glmMixZINB(mg.PAR_MIX, g["m_id"], g["m_w"], g["m_y"], g["m_off"], g["m_dsg"], B + 2, MINZERO)
    assert mv == float(g["m_v"]) and np.array_equal(mgr, g["m_g"])
