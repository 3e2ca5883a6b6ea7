"""callPeaks on the device (R/callPeaks.R:61-68, fdrControl R/base-utils.R:11-25): called windows bit-identical
to the oracle's (stable sort, R's long double cumsum), on both summation paths; merged peaks; the final flush."""
import numpy as np
import pytest

from conftest import has_gpu
from epigrahmm_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_gpu(), reason="needs a CUDA device")]


def _session_with_posteriors(ehmm, prob, K=2):
    """a session whose logProb1[, 2] is log(prob) (function boundary: the dataset is stored as the file would hold it)"""
    M = prob.size
    lp = np.empty((M, K), order="F")
    with np.errstate(divide="ignore"):
        lp[:, 1] = np.log(prob)
        lp[:, 0] = np.log1p(-prob) if K == 2 else np.log1p(-prob) - np.log(2.0)
    if K == 3:
        lp[:, 2] = lp[:, 0]
    lp = np.maximum(lp, -1e300)
    s = ehmm.Session("t_peaks", M, K)
    s.store("logProb1", lp)
    return s, np.exp(lp[:, 1])


@pytest.mark.parametrize("M", [1, 2, 31, 4095, 4096, 4097, 100_003, 1_500_000])
@pytest.mark.parametrize("fdr", [0.05, 0.3])
def test_fdr_calls_match_the_oracle(ehmm, oracle, M, fdr):
    rng = np.random.default_rng(M + int(100 * fdr))
    prob = np.where(rng.random(M) < 0.15, 1.0 - rng.random(M) ** 3 * 0.2, rng.random(M) ** 6)
    prob[rng.integers(0, M, M // 7 + 1)] = 1.0      # exp(0): ties at notpp = 0
    prob[rng.integers(0, M, M // 9 + 1)] = 0.5      # ties in the middle
    s, pr = _session_with_posteriors(ehmm, prob)
    with s:
        want = oracle.fdrControl(pr, fdr)
        got = s.call_peaks(fdr)
        assert s.call_peaks_path().startswith("fp64")
        assert np.array_equal(got, want), int(np.sum(got != want))
        slow = s.call_peaks(fdr, exact=True)
        assert s.call_peaks_path().startswith("emulated") and np.array_equal(slow, want)
        if want.any():
            assert s.peaks_cutoff == np.max(1.0 - pr[want])


def test_reference_known_answer(ehmm, oracle):
    """tests/testthat/test-base.R:1-4: sum(fdrControl(seq(0, 1, 0.001), 0.05)) == 101, a vector whose 101st running mean
    IS the level up to rounding.  The oracle reproduces the 101 on the exact sequence (tests/test_ldsum.py); the device
    receives the probabilities as logProb1, i.e. exp(log(.)) of the sequence, and must agree with the oracle on those,
    window by window -- through the emulated long double, the margin being too small to certify"""
    s, pr = _session_with_posteriors(ehmm, np.arange(1001) * 0.001)
    with s:
        got = s.call_peaks(0.05)
        assert s.call_peaks_path() == "emulated long double (margin too small)"
    assert got.sum() in (100, 101) and np.array_equal(got, oracle.fdrControl(pr, 0.05))


def test_margin_too_small_falls_to_r_arithmetic(ehmm, oracle):
    """running means that sit ON the level (every notpp equal to it, so each mean is the level up to rounding): the
    fp64 prefix sum cannot certify the comparison and the emulated long double decides, as R would"""
    M = 50_000
    prob = np.full(M, 1.0 - 0.05)
    s, pr = _session_with_posteriors(ehmm, prob)
    with s:
        got = s.call_peaks(0.05)
        assert s.call_peaks_path() == "emulated long double (margin too small)"
        assert np.array_equal(got, oracle.fdrControl(pr, 0.05))


def test_probabilities_outside_unit_interval(ehmm):
    lp = np.zeros((100, 2), order="F")
    lp[7, 1] = 0.5     # exp(0.5) > 1
    with ehmm.Session("t_badp", 100, 2) as s:
        s.store("logProb1", lp)
        with pytest.raises(ehmm.EhmmError, match="between 0 and 1"):
            s.call_peaks(0.05)
        with pytest.raises(ehmm.EhmmError, match="fdr"):
            s.call_peaks(1.5)


def _runs_host(calls, breaks):
    starts, ends = [], []
    M = calls.size
    bset = set(int(b) for b in breaks)
    j = 0
    while j < M:
        if calls[j]:
            a = j
            while j + 1 < M and calls[j + 1] and (j + 1) not in bset:
                j += 1
            starts.append(a)
            ends.append(j)
        j += 1
    return np.array(starts, dtype=np.int64), np.array(ends, dtype=np.int64)


@pytest.mark.parametrize("M", [1, 5000, 300_007])
def test_peak_runs(ehmm, M):
    rng = np.random.default_rng(M)
    z = synth.markov_chain(M, synth.GAMMA2, rng) if M > 1 else np.array([1])
    vit = z.astype(np.float64)
    breaks = np.unique(rng.integers(1, max(M, 2), 6)) if M > 1 else np.array([], dtype=np.int64)
    with ehmm.Session("t_runs", M, 2) as s:
        s.store("viterbi", vit)
        calls = s.call_peaks("viterbi")
        assert np.array_equal(calls, vit == 1)
        for b in (np.array([], dtype=np.int64), breaks):
            st, en = s.peak_runs(b)
            ws, we = _runs_host(calls, b)
            assert np.array_equal(st, ws) and np.array_equal(en, we)


def test_flush_and_callpeaks_after_a_fit(ehmm, oracle):
    """end of a fit run with the EM-internal E-step: one flush hands over what the reference's file holds, and the
    calls on it equal the host path on the fetched posteriors"""
    from epigrahmm_b200 import em
    M, N = 30_011, 3
    d = synth.make_consensus(M, N, seed=21)
    th = synth.THETA_CONSENSUS
    with ehmm.Session("t_flush", M, 2) as s:
        s.set_data(d["counts"], d["offsets"])
        s.expstep_mode(1)
        s.loglik_consensus(th["psi"], 2.2250738585072014e-308)
        s.expstep(th["pi"], th["gamma"])
        s.viterbi(th["pi"], th["gamma"], fetch=False)
        out = s.flush()
        assert sorted(out) == ["logBP", "logFP", "logLikelihood", "logProb1", "logProb2", "viterbi"]
        for n, a in out.items():
            assert np.array_equal(a, s.fetch(n)), n
        prob = np.exp(out["logProb1"][:, 1])
        for method in (0.05, 0.2):
            assert np.array_equal(em.callPeaks(s, method), oracle.fdrControl(prob, method))
        assert np.array_equal(em.callPeaks(s, "viterbi"), out["viterbi"][:, 0] == 1)
