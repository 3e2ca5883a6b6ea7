"""GPU parity: emission + expStep + maxStepProb + Q + BIC through the C ABI against the oracle.

Tolerances (BASELINE.json north_star): posteriors <= 1e-8 absolute, log-likelihood (and the
logFP/logBP it is built from) <= 1e-9 relative, parameters <= 1e-9 relative.
"""
import numpy as np
import pytest

from epigrahmm_b200 import synth

pytestmark = pytest.mark.gpu

MINZERO = 2.2250738585072014e-308


def _theta(K):
    t = synth.THETA_CONSENSUS if K == 2 else synth.THETA_DIFFERENTIAL
    return t["pi"], t["gamma"]


def _check_estep(h, g, M, K):
    for n in ("logFP", "logBP"):
        rel = np.max(np.abs(h[n] - g[n]) / np.maximum(1.0, np.abs(h[n])))
        assert rel <= 1e-9, (n, rel)
    for n in ("logProb1", "logProb2"):
        d = np.max(np.abs(np.exp(h[n]) - np.exp(g[n])))
        assert d <= 1e-8, (n, d)
        # the raw log datasets agree too wherever the probability is not negligible
        m = h[n] > -30
        assert np.max(np.abs(h[n][m] - g[n][m])) <= 1e-6, n
        assert np.all(np.isfinite(g[n])), n
    assert np.allclose(g["logProb2"][0], -np.log(K * K))


@pytest.mark.parametrize("K", [2, 3])
@pytest.mark.parametrize("M", [1, 2, 7, 400, 959, 960, 961, 1280, 1281, 7000, 100003])
def test_expstep_matches_oracle(ehmm, oracle, K, M):
    rng = np.random.default_rng(1000 * K + M)
    z = synth.markov_chain(M, synth.GAMMA2 if K == 2 else synth.GAMMA3, rng)
    logf = rng.normal(-6.0, 2.0, (M, K))
    logf[np.arange(M), z] += 5.0
    pi, gamma = _theta(K)
    h = oracle.expStep(pi, gamma, logf)
    with ehmm.Session("t_estep", M, K) as s:
        s.expstep(pi, gamma, logf)
        g = s.datasets()
        _check_estep(h, g, M, K)
        assert np.array_equal(g["logLikelihood"], np.asfortranarray(logf))
        ll = oracle.computeBIC(h, 0, 1) / -2.0
        assert abs(s.loglik() - ll) <= 1e-9 * abs(ll)
        if M >= 2:
            mo, mg = oracle.maxStepProb(h), s.maxstepprob()
            assert np.allclose(mo["pi"], mg["pi"], rtol=1e-9, atol=0)
            assert np.allclose(mo["gamma"], mg["gamma"], rtol=1e-9, atol=0)
            qo, qg = oracle.computeQFunction(h, pi, gamma), s.qfunction(pi, gamma)
            assert abs(qo - qg) <= 1e-9 * abs(qo)
        bo, bg = oracle.computeBIC(h, 7, 3), s.bic(7, 3)
        assert abs(bo - bg) <= 1e-9 * abs(bo)
        assert np.allclose(s.marginal_probability(), np.exp(h["logProb1"]), atol=1e-8, rtol=0)


@pytest.mark.parametrize("K", [2, 3])
def test_expstep_emission_underflow(ehmm, oracle, K):
    """states whose emission underflows exp() must keep finite log posteriors (rejection control
    distinguishes exp(logProb1) == 0 from a subnormal)"""
    M = 3000
    rng = np.random.default_rng(7)
    logf = rng.normal(-5.0, 1.0, (M, K))
    for j in (0, 4, 5, 100, 959, 960, 1500, M - 1):
        logf[j, 1] = -900.0
    logf[2000, 0] = -1500.0
    pi, gamma = _theta(K)
    h = oracle.expStep(pi, gamma, logf)
    with ehmm.Session("t_underflow", M, K) as s:
        s.expstep(pi, gamma, logf)
        g = s.datasets()
    _check_estep(h, g, M, K)
    for n in ("logProb1", "logProb2"):
        m = h[n] < -300
        assert m.any()
        assert np.max(np.abs(h[n][m] - g[n][m]) / np.abs(h[n][m])) <= 1e-9


def test_expstep_errors(ehmm):
    with pytest.raises(ehmm.EhmmError):
        ehmm.Session("bad", 10, 5)
    with ehmm.Session("t_err", 10, 2) as s:
        with pytest.raises(ehmm.EhmmError):
            s.maxstepprob()          # the reference would fail to load logProb1 from the file
        with pytest.raises(ehmm.EhmmError):
            s.expstep([0.5, 0.5], [[0.9, 0.1], [0.1, 0.9]])  # no logf anywhere


def test_functions_on_stored_datasets(ehmm, oracle):
    """maxStepProb / computeQFunction / computeBIC read their inputs from the 'file': feeding the
    oracle's datasets must reproduce the oracle's outputs (function-boundary parity)."""
    M, K = 5000, 3
    rng = np.random.default_rng(5)
    logf = rng.normal(-4.0, 2.0, (M, K))
    pi, gamma = _theta(K)
    h = oracle.expStep(pi, gamma, logf)
    with ehmm.Session("t_stored", M, K) as s:
        for n in ("logLikelihood", "logFP", "logBP", "logProb1", "logProb2"):
            s.store(n, h[n])
        mo, mg = oracle.maxStepProb(h), s.maxstepprob()
        assert np.allclose(mo["pi"], mg["pi"], rtol=1e-12, atol=0)
        assert np.allclose(mo["gamma"], mg["gamma"], rtol=1e-12, atol=0)
        assert abs(oracle.computeQFunction(h, pi, gamma) - s.qfunction(pi, gamma)) <= 1e-12 * abs(s.qfunction(pi, gamma))
        assert abs(oracle.computeBIC(h, 4, 2) - s.bic(4, 2)) <= 1e-13 * abs(s.bic(4, 2))


@pytest.mark.parametrize("controls", [False, True])
def test_loglik_consensus(ehmm, oracle, controls):
    M, N = 20011, 3
    d = synth.make_consensus(M, N, seed=11, controls=controls)
    d["counts"][5, 0] = 9000  # beyond the lgamma table
    d["counts"][6, 1] = 40000
    psi = [np.array([np.log(2.0), 0.3, 3.0]), np.array([np.log(12.0), 0.1, 4.0])] if controls else \
        [np.array([np.log(2.0), 3.0]), np.array([np.log(12.0), 4.0])]
    ref = oracle.loglikConsensusHMM(d["counts"], d["offsets"], psi, MINZERO, d["controls"])
    with ehmm.Session("t_cons", M, 2) as s:
        s.set_data(d["counts"], d["offsets"], d["controls"])
        s.loglik_consensus(psi, MINZERO)
        got = s.fetch("logLikelihood")
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) <= 1e-11


def test_controls_quirk_on_bin_blocks_and_encoded_tables(ehmm, oracle):
    """the reference's emission uses controls[1] of the WHOLE table for every cell (its ifelse() quirk,
    R/base-loglikelihood.R:113): a bin block with m0 > 0 must be given block 0's value, and the
    dictionary-encoded upload finds it through the first cell's group"""
    M, N = 6001, 3
    d = synth.make_consensus(M, N, seed=13, controls=True)
    g = synth.make_groups(d["counts"], d["offsets"], d["controls"])
    psi = [np.array([np.log(2.0), 0.3, 3.0]), np.array([np.log(12.0), 0.1, 4.0])]
    ref = oracle.loglikConsensusHMM(d["counts"], d["offsets"], psi, MINZERO, d["controls"])
    lo = 2500
    with ehmm.Session("t_ctrl_blk", M - lo, 2) as s:
        s.set_block(lo, M)
        s.set_data(d["counts"][lo:], d["offsets"][lo:], d["controls"][lo:])
        assert s.controls_first() == d["controls"][lo, 0]          # the block's own first cell: wrong for the chain
        s.set_controls_first(d["controls"][0, 0])
        s.loglik_consensus(psi, MINZERO)
        got = s.fetch("logLikelihood")
    assert np.max(np.abs(got - ref[lo:]) / np.maximum(1.0, np.abs(ref[lo:]))) <= 1e-11
    with ehmm.Session("t_ctrl_enc", M, 2) as s:
        s.set_data_encoded(g["group"], g["u_chip"], g["u_offsets"], g["u_controls"])
        assert s.controls_first() == d["controls"][0, 0]
        s.loglik_consensus(psi, MINZERO)
        got = s.fetch("logLikelihood")
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) <= 1e-11


def test_loglik_init(ehmm, oracle):
    M = 9001
    d = synth.make_consensus(M, 1, seed=12)
    psi = [np.array([0.4, 2.5]), np.array([2.2, 900.0])]
    ref = oracle.loglikInit(d["counts"], d["offsets"], psi)
    with ehmm.Session("t_init", M, 2) as s:
        s.set_data(d["counts"], d["offsets"])
        s.loglik_init(psi)
        got = s.fetch("logLikelihood")
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) <= 1e-11


@pytest.mark.parametrize("n_cond,n_rep", [(2, 1), (3, 2), (4, 3), (5, 2)])
def test_loglik_differential(ehmm, oracle, n_cond, n_rep):
    M = 6007
    d = synth.make_differential(M, n_cond, n_rep, seed=13)
    B = d["B"]
    delta = np.random.default_rng(1).dirichlet(np.ones(B))
    psi = synth.THETA_DIFFERENTIAL["psi"]
    ref = oracle.loglikDifferentialHMM(d["counts"], d["offsets"], d["design"], delta, psi, MINZERO)
    eta_ref = oracle.innerExpStep(d["counts"], d["offsets"], d["design"], delta, psi, MINZERO)
    with ehmm.Session("t_diff", M, 3) as s:
        s.set_data(d["counts"], d["offsets"])
        s.set_design(d["design"])
        s.loglik_differential_hmm(delta, psi, MINZERO)
        got = s.fetch("logLikelihood")
        s.inner_expstep(delta, psi, MINZERO)
        eta = s.fetch("mixtureProb")
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) <= 1e-11
    assert np.max(np.abs(eta - eta_ref)) <= 1e-10
    assert np.allclose(eta.sum(1), 1.0, atol=1e-12)


def test_reference_kat_loglik_differential(ehmm, oracle):
    """the reference's only emission known-answer test (tests/testthat/test-base.R:6-42):
    ChIP = 1, offsets = -1, psi = (0,1;0,1), delta = (.1,.9): three unique rows with closed forms."""
    M = 1000
    counts = np.ones((M, 2), dtype=np.int32)
    offsets = -np.ones((M, 2))
    design = np.array([[1, 0], [0, 1]], dtype=np.uint8)
    psi = np.array([0.0, 1.0, 0.0, 1.0])  # unlist(list(betas=c(0,1), lambdas=c(0,1)))
    with ehmm.Session("t_kat", M, 3) as s:
        s.set_data(counts, offsets)
        s.set_design(design)
        s.loglik_differential_hmm([0.1, 0.9], psi, 1e-32)
        ll = s.fetch("logLikelihood")
    rows = np.unique(ll, axis=0)
    assert rows.shape == (1, 3)
    # psi order follows unlist(theta$psi) = (0,1,0,1): mu = exp(0 + 0*D - 1), size = exp(1 + 1*D)
    d1 = oracle.dnbinom(1, np.exp(1.0), np.exp(-1.0), log=True)
    d2 = oracle.dnbinom(1, np.exp(2.0), np.exp(-1.0), log=True)
    assert abs(rows[0, 0] - 2 * d1) <= 1.5e-8
    assert abs(rows[0, 1] - (d1 + d2)) <= 1.5e-8
    assert abs(rows[0, 2] - 2 * d2) <= 1.5e-8


def test_inner_maxstepprob(ehmm, oracle):
    M, K = 4001, 3
    rng = np.random.default_rng(3)
    logf = rng.normal(-4.0, 2.0, (M, K))
    pi, gamma = _theta(K)
    h = oracle.expStep(pi, gamma, logf)
    eta = rng.dirichlet(np.ones(6), M)
    h["mixtureProb"] = eta
    ref = oracle.innerMaxStepProb(h)
    with ehmm.Session("t_inner", M, K) as s:
        s.store("logProb1", h["logProb1"])
        s.store("mixtureProb", eta)
        got = s.inner_maxstepprob()
    assert np.allclose(got, ref, rtol=1e-12, atol=0)


@pytest.mark.parametrize("model", ["consensus", "consensus_ctrl", "differential"])
def test_dictionary_encoded_emission_is_bit_identical(ehmm, model):
    """with dt$Group / dtUnique uploaded the emission gathers per-group densities instead of
    evaluating every cell; both paths must give the same bits"""
    M = 30011
    if model == "differential":
        d = synth.make_differential(M, 3, 2, seed=41)
        g = synth.make_groups(d["counts"], d["offsets"], design=d["design"])
        delta = np.random.default_rng(2).dirichlet(np.ones(d["B"]))
        psi = synth.THETA_DIFFERENTIAL["psi"]
        with ehmm.Session("t_dict", M, 3) as s:
            s.set_data(d["counts"], d["offsets"])
            s.set_design(d["design"])
            s.loglik_differential_hmm(delta, psi, MINZERO)
            a = s.fetch("logLikelihood")
            s.inner_expstep(delta, psi, MINZERO)
            ea = s.fetch("mixtureProb")
            s.set_groups(g["group"], g["u_chip"], g["u_offsets"], None, g["u_dsg"])
            s.profile(True)
            s.loglik_differential_hmm(delta, psi, MINZERO)
            b = s.fetch("logLikelihood")
            s.inner_expstep(delta, psi, MINZERO)
            eb = s.fetch("mixtureProb")
            assert s.profile_read()["emis_tables"][0] >= 2
        assert np.array_equal(a, b) and np.array_equal(ea, eb)
        return
    ctrl = model == "consensus_ctrl"
    d = synth.make_consensus(M, 3, seed=42, controls=ctrl)
    g = synth.make_groups(d["counts"], d["offsets"], d["controls"])
    psi = [np.array([np.log(2.0), 0.3, 3.0]), np.array([np.log(12.0), 0.1, 4.0])] if ctrl else synth.THETA_CONSENSUS["psi"]
    with ehmm.Session("t_dict", M, 2) as s:
        s.set_data(d["counts"], d["offsets"], d["controls"])
        s.loglik_consensus(psi, MINZERO)
        a = s.fetch("logLikelihood")
        s.set_groups(g["group"], g["u_chip"], g["u_offsets"], g["u_controls"])
        s.profile(True)
        s.loglik_consensus(psi, MINZERO)
        b = s.fetch("logLikelihood")
        assert s.profile_read()["emis_gather"][0] == 1
        # a dtUnique that does not describe the data must not be used for the emission
        s.set_groups(g["group"], g["u_chip"] + 1.0, g["u_offsets"], g["u_controls"])
        s.profile(True)
        s.loglik_consensus(psi, MINZERO)
        c = s.fetch("logLikelihood")
        assert s.profile_read()["emis_gather"][0] == 0
    assert np.array_equal(a, b) and np.array_equal(a, c)


@pytest.mark.parametrize("model", ["consensus", "differential"])
def test_dictionary_encoded_upload(ehmm, model):
    """ehmm_set_data_encoded (dt$Group + dtUnique only) gives the same emission as the full table"""
    M = 20011
    if model == "consensus":
        d = synth.make_consensus(M, 3, seed=51, controls=True)
        g = synth.make_groups(d["counts"], d["offsets"], d["controls"])
        psi = [np.array([np.log(2.0), 0.3, 3.0]), np.array([np.log(12.0), 0.1, 4.0])]
        with ehmm.Session("t_enc_a", M, 2) as a, ehmm.Session("t_enc_b", M, 2) as b:
            a.set_data(d["counts"], d["offsets"], d["controls"])
            a.loglik_consensus(psi, MINZERO)
            b.set_data_encoded(g["group"], g["u_chip"], g["u_offsets"], g["u_controls"])
            b.loglik_consensus(psi, MINZERO)
            assert np.array_equal(a.fetch("logLikelihood"), b.fetch("logLikelihood"))
            assert g["G"] < 65536
            b.set_data_encoded(g["group"].astype(np.uint16), g["u_chip"], g["u_offsets"], g["u_controls"])  # 16-bit ids
            b.loglik_consensus(psi, MINZERO)
            assert np.array_equal(a.fetch("logLikelihood"), b.fetch("logLikelihood"))
            with pytest.raises(ehmm.EhmmError):
                b.set_data_encoded(g["group"] + g["G"], g["u_chip"], g["u_offsets"], g["u_controls"])  # ids out of range
    else:
        d = synth.make_differential(M, 3, 2, seed=52)
        g = synth.make_groups(d["counts"], d["offsets"], design=d["design"])
        delta = np.full(d["B"], 1.0 / d["B"])
        psi = synth.THETA_DIFFERENTIAL["psi"]
        with ehmm.Session("t_enc_a", M, 3) as a, ehmm.Session("t_enc_b", M, 3) as b:
            a.set_data(d["counts"], d["offsets"])
            a.set_design(d["design"])
            a.loglik_differential_hmm(delta, psi, MINZERO)
            a.inner_expstep(delta, psi, MINZERO)
            b.set_data_encoded(g["group"], g["u_chip"], g["u_offsets"], None, g["u_dsg"], design=d["design"])
            b.loglik_differential_hmm(delta, psi, MINZERO)
            b.inner_expstep(delta, psi, MINZERO)
            assert np.array_equal(a.fetch("logLikelihood"), b.fetch("logLikelihood"))
            assert np.array_equal(a.fetch("mixtureProb"), b.fetch("mixtureProb"))


# ---- the EM-internal form of the E-step (ehmm_expstep_mode) ---------------------------------------

def _lean_logf(K, M, seed):
    rng = np.random.default_rng(seed)
    z = synth.markov_chain(M, synth.GAMMA2 if K == 2 else synth.GAMMA3, rng)
    logf = rng.normal(-6.0, 2.0, (M, K))
    logf[np.arange(M), z] += 5.0
    for j in (0, 5, 100, 959, 960, 1279, 1280, 1500, M - 1):   # emissions that underflow exp()
        if j < M:
            logf[j, 1] = -900.0
    if M > 2000:
        logf[2000, 0] = -1500.0
    return logf


@pytest.mark.parametrize("K", [2, 3])
@pytest.mark.parametrize("M", [1, 2, 959, 1281, 7000, 100003])
def test_lean_expstep_owes_the_same_datasets(ehmm, oracle, K, M):
    """mode 1 writes P1 and the statistics only; what it produces on demand holds the bits of a mode-0 E-step,
    and everything the EM loop derives from it (pi, gamma, Q, BIC, log-likelihood) is the same number"""
    logf = _lean_logf(K, M, 50 * K + M)
    pi, gamma = _theta(K)
    with ehmm.Session("t_full", M, K) as s:
        s.expstep(pi, gamma, logf)
        full = s.datasets()
        ref = (s.maxstepprob() if M >= 2 else None, s.qfunction(pi, gamma), s.bic(5, 2), s.loglik())
    with ehmm.Session("t_lean", M, K) as s:
        s.expstep_mode(1)
        s.expstep(pi, gamma, logf)
        got = (s.maxstepprob() if M >= 2 else None, s.qfunction(pi, gamma), s.bic(5, 2), s.loglik())
        lean = s.datasets()
    for n in ("logLikelihood", "logFP", "logBP", "logProb1", "logProb2"):
        assert np.array_equal(full[n], lean[n]), n
    assert got[1:] == ref[1:]
    if M >= 2:
        assert np.array_equal(got[0]["pi"], ref[0]["pi"]) and np.array_equal(got[0]["gamma"], ref[0]["gamma"], equal_nan=True)


def test_lean_expstep_survives_the_next_emission(ehmm, oracle):
    """the datasets are those of the LAST E-step even when a new emission has been computed since (the
    reference's file still holds the old E-step at that point)"""
    M, N = 30011, 3
    d = synth.make_consensus(M, N, seed=11)
    th = synth.THETA_CONSENSUS
    psi2 = [th["psi"][0] * np.array([1.1, 0.9]), th["psi"][1] * np.array([0.95, 1.2])]
    with ehmm.Session("t_keep", M, 2) as s:
        s.set_data(d["counts"], d["offsets"])
        s.loglik_consensus(th["psi"], MINZERO)
        s.expstep(th["pi"], th["gamma"])
        want = s.datasets()
        s.expstep_mode(1)
        s.loglik_consensus(th["psi"], MINZERO)
        s.expstep(th["pi"], th["gamma"])
        s.loglik_consensus(psi2, MINZERO)          # the next iteration's emission
        newer = s.fetch("logLikelihood")
        got = s.datasets()
        assert np.array_equal(got["logLikelihood"], newer) and not np.array_equal(newer, want["logLikelihood"])
        for n in ("logFP", "logBP", "logProb1", "logProb2"):
            assert np.array_equal(got[n], want[n]), n
        # and the E-step after that one runs on the new emission
        s.expstep(th["pi"], th["gamma"])
        h = oracle.expStep(th["pi"], th["gamma"], newer)
        assert np.max(np.abs(np.exp(s.fetch("logProb1")) - np.exp(h["logProb1"]))) <= 1e-8


@pytest.mark.parametrize("p", [0.9, 0.05])
def test_lean_expstep_feeds_the_rejection_control(ehmm, oracle, p):
    """K = 2: the lean kernel counts the weights below the announced cut itself and the rejection control thins
    the linear posteriors; support, draw count and sums as from the oracle's posteriors"""
    M, N = 40009, 3
    d = synth.make_consensus(M, N, seed=3)
    g = synth.make_groups(d["counts"], d["offsets"])
    th = synth.THETA_CONSENSUS
    u = oracle.RNG.set_seed(5).runif(2 * M)
    logf = oracle.loglikConsensusHMM(d["counts"], d["offsets"], th["psi"], MINZERO)
    h = oracle.expStep(th["pi"], th["gamma"], logf)
    # bins whose posterior underflows: exp(logProb1) == 0 draws nothing, a subnormal does
    ref, n_ref = oracle.consensusRejectionControlled(h, g["group"], p, uniforms=u, dense=True)
    with ehmm.Session("t_lean_rc", M, 2) as s:
        s.set_data(d["counts"], d["offsets"])
        s.set_groups(g["group"], g["u_chip"], g["u_offsets"])
        s.expstep_mode(1)
        s.rc_hint(p)
        s.loglik_consensus(th["psi"], MINZERO)
        s.profile(True)
        s.expstep(th["pi"], th["gamma"])
        n = s.rc_consensus(p, uniforms=u)
        prof = s.profile_read()
        s.profile(False)
        assert prof["fb_apply_em"][0] == 1 and prof["fb_apply"][0] == 0 and prof["rc_count"][0] == 0
        assert n == n_ref
        for c in range(2):
            got = s.rc_buckets(c)
            assert np.array_equal(got != 0, ref[c] != 0)
            # (the oracle's own posteriors carry ulp(|logLik|) of noise: 1e-9 as in the trajectory gates)
            assert np.max(np.abs(got - ref[c]) / np.maximum(np.abs(ref[c]), 1e-300)) <= 1e-9
        # a cut that was not announced: the counting kernel runs after all
        s.profile(True)
        s.rc_consensus(0.5 * p, uniforms=u)
        assert s.profile_read()["rc_count"][0] == 1
        s.profile(False)
