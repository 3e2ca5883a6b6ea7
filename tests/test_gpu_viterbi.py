"""computeViterbiSequence (src/computeViterbiSequence.cpp:12-54): the parallel (max,+) scan against the
oracle's sequential loop and against the library's own sequential kernel, bit for bit, on the same logFP.

The recursion is scored with logFP, so LOGV reaches 1e15 and its decisions are taken at ulp granularity;
the scan replaces fp64 additions by integer additions inside one binade of LOGV.  The cases below aim at
what that model has to get right: binade crossings, half-way roundings (ties-to-even depends on the
running value), large differences between the states, tile edges, and inputs it cannot represent."""
import numpy as np
import pytest

from conftest import has_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_gpu(), reason="needs a CUDA device")]

PI = {2: np.array([0.7, 0.3]), 3: np.array([0.5, 0.3, 0.2])}
GAMMA = {2: np.array([[0.998, 0.002], [0.01, 0.99]]),
         3: np.array([[0.998, 0.001, 0.001], [0.005, 0.99, 0.005], [0.001, 0.001, 0.998]])}


def forward_like(rng, M, K, step=8.0, spread=3.0):
    """something shaped like logFP: a decreasing common part plus a per-state offset"""
    base = -np.cumsum(rng.gamma(2.0, step / 2.0, M))
    return np.asfortranarray(base[:, None] - rng.gamma(1.0, spread, (M, K)))


def run_all(ehmm, oracle, logFP, K, pi=None, gamma=None):
    pi = PI[K] if pi is None else pi
    gamma = GAMMA[K] if gamma is None else gamma
    M = logFP.shape[0]
    ref = oracle.computeViterbiSequence({"logFP": logFP}, pi, gamma)
    with ehmm.Session("vit", M, K, 0) as s:
        s.store("logFP", logFP)
        got = s.viterbi(pi, gamma)
        info = s.viterbi_info()
        s.viterbi_mode(1)
        seq = s.viterbi(pi, gamma)
        assert s.viterbi_info()["path"] == "sequential"
    assert np.array_equal(seq, ref), "sequential kernel differs from the oracle in %d bins" % int(np.sum(seq != ref))
    assert np.array_equal(got, ref), "scan differs from the oracle in %d bins (first %d)" % (
        int(np.sum(got != ref)), int(np.flatnonzero(got != ref)[0]))
    return info, ref


@pytest.mark.parametrize("K", [2, 3])
@pytest.mark.parametrize("M", [1, 2, 3, 511, 512, 513, 1024, 2047, 2049, 40_009, 300_007])
def test_scan_matches_oracle(ehmm, oracle, K, M):
    rng = np.random.default_rng(100 * K + M % 97)
    info, ref = run_all(ehmm, oracle, forward_like(rng, M, K), K)
    assert info["path"] == "scan"
    if M > 100_000:
        assert info["break_tiles"] < 0.25 * info["tiles"]   # the bulk of the chain runs in integers
        assert 0 < ref.mean() < K - 1 + 1e-9


@pytest.mark.parametrize("K", [2, 3])
def test_two_million_bins(ehmm, oracle, K):
    rng = np.random.default_rng(7 + K)
    info, _ = run_all(ehmm, oracle, forward_like(rng, 2_000_003, K), K)
    assert info["path"] == "scan" and info["break_tiles"] < 200


@pytest.mark.parametrize("K", [2, 3])
def test_halfway_roundings(ehmm, oracle, K):
    """logFP on a grid of 1/16 with steps of 1e11..1e12: LOGV enters the binades whose ulp is 1/8 .. 1/2
    within a few thousand bins, where every second addend is a half-way case"""
    rng = np.random.default_rng(31 + K)
    M = 60_000
    base = -np.cumsum(np.round(rng.uniform(1e11, 1e12, M) * 16) / 16)
    lf = np.asfortranarray(base[:, None] - np.round(rng.gamma(1.0, 3.0, (M, K)) * 16) / 16)
    info, _ = run_all(ehmm, oracle, lf, K)
    assert info["path"] == "scan" and info["halfway_tiles"] > 0


@pytest.mark.parametrize("K", [2, 3])
def test_rare_halfway_roundings(ehmm, oracle, K):
    """a handful of half-way addends in an otherwise generic chain: those tiles alone leave the integer path"""
    rng = np.random.default_rng(57 + K)
    M = 200_000
    lf = forward_like(rng, M, K, step=4e9)           # ulp of LOGV: 2^-8 .. 2^-3 over most of the chain
    at = rng.choice(np.arange(50_000, M), 40, replace=False)
    lf[at, rng.integers(0, K, 40)] = np.round(lf[at, 0] * 8) / 8 + 1.0 / 16
    info, _ = run_all(ehmm, oracle, lf, K)
    assert info["path"] == "scan" and 0 < info["halfway_tiles"] <= 40


@pytest.mark.parametrize("K", [2, 3])
def test_large_state_differences(ehmm, oracle, K):
    """states that are hundreds of log units apart (an emission that all but excludes a state)"""
    rng = np.random.default_rng(77 + K)
    M = 150_000
    lf = forward_like(rng, M, K)
    lf[rng.choice(M, 3000, replace=False), rng.integers(0, K, 3000)] -= rng.uniform(100, 5000, 3000)
    info, _ = run_all(ehmm, oracle, lf, K)
    assert info["path"] == "scan"


@pytest.mark.parametrize("K", [2, 3])
def test_flat_and_tiny_inputs(ehmm, oracle, K):
    rng = np.random.default_rng(91 + K)
    for lf in (np.zeros((5000, K)), -rng.uniform(0, 1e-12, (5000, K)), -np.full((700, K), 1e-300)):
        run_all(ehmm, oracle, np.asfortranarray(lf), K)


@pytest.mark.parametrize("K", [2, 3])
def test_inputs_outside_the_integer_model(ehmm, oracle, K):
    """-inf and positive entries (the reference would take them as they are): the sequential kernel decides"""
    rng = np.random.default_rng(13 + K)
    lf = forward_like(rng, 30_000, K)
    lf[12_345, 0] = -np.inf
    info, _ = run_all(ehmm, oracle, lf, K)
    assert info["path"] == "scan abandoned"
    lf = forward_like(rng, 30_000, K)
    lf[100:200] += 5000.0
    info, _ = run_all(ehmm, oracle, lf, K)
    assert info["path"] == "scan abandoned"
    # a zero in pi or gamma: log gives -inf, every tile is planned sequential
    pi = np.array([1.0, 0.0, 0.0][:K])
    g = GAMMA[K].copy()
    g[0, 1] += g[0, 0] - 0.0
    g[0, 0] = 0.0
    info, _ = run_all(ehmm, oracle, forward_like(rng, 20_000, K), K, pi=pi, gamma=g)
    assert info["break_tiles"] == info["tiles"]


def test_bin_blocks_hand_on_logv(ehmm, oracle):
    """ehmm_viterbi_forward / _backtrace over three uneven blocks equal the whole chain"""
    K, M = 3, 250_000
    rng = np.random.default_rng(5)
    lf = forward_like(rng, M, K)
    ref = oracle.computeViterbiSequence({"logFP": lf}, PI[K], GAMMA[K])
    cuts = [0, 70_001, 70_001 + 512 * 100, M]
    sess = []
    v = None
    for i in range(3):
        a, b = cuts[i], cuts[i + 1]
        s = ehmm.Session("vb%d" % i, b - a, K, 0)
        s.set_block(a, M)
        s.store("logFP", np.asfortranarray(lf[a:b]))
        v = s.viterbi_forward(PI[K], GAMMA[K], v)
        assert s.viterbi_info()["path"] == "scan"
        sess.append(s)
    state, out = -1, [None] * 3
    for i in (2, 1, 0):
        out[i], state = sess[i].viterbi_backtrace(state)
        sess[i].close()
    assert np.array_equal(np.concatenate(out), ref)
