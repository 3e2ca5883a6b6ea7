"""GPU parity: rejection control, aggregation, NB-GLM objective/gradient and Viterbi through the
C ABI against the oracle, on shared inputs (function-boundary protocol)."""
import numpy as np
import pytest

from epigrahmm_b200 import synth

pytestmark = pytest.mark.gpu

MINZERO = 2.2250738585072014e-308


def _consensus_case(oracle, M=20011, N=3, seed=21, controls=False):
    d = synth.make_consensus(M, N, seed=seed, controls=controls)
    g = synth.make_groups(d["counts"], d["offsets"], d["controls"])
    th = synth.THETA_CONSENSUS
    psi = th["psi"] if not controls else [np.array([np.log(2.0), 0.2, 3.0]), np.array([np.log(12.0), 0.1, 4.0])]
    logf = oracle.loglikConsensusHMM(d["counts"], d["offsets"], psi, MINZERO, d["controls"])
    h = oracle.expStep(th["pi"], th["gamma"], logf)
    return d, g, h, psi


def _differential_case(oracle, M=8009, n_cond=3, n_rep=2, seed=22):
    d = synth.make_differential(M, n_cond, n_rep, seed=seed)
    g = synth.make_groups(d["counts"], d["offsets"], design=d["design"])
    th = synth.THETA_DIFFERENTIAL
    delta = np.full(d["B"], 1.0 / d["B"])
    logf = oracle.loglikDifferentialHMM(d["counts"], d["offsets"], d["design"], delta, th["psi"], MINZERO)
    h = oracle.expStep(th["pi"], th["gamma"], logf)
    h["mixtureProb"] = oracle.innerExpStep(d["counts"], d["offsets"], d["design"], delta, th["psi"], MINZERO)
    return d, g, h, delta


def _open(ehmm, key, d, g, K, h, design=None):
    s = ehmm.Session(key, d["counts"].shape[0], K)
    s.set_data(d["counts"], d["offsets"], d.get("controls"))
    if design is not None:
        s.set_design(design)
    s.set_groups(g["group"], g["u_chip"], g["u_offsets"], g["u_controls"], g["u_dsg"])
    for n in ("logLikelihood", "logFP", "logBP", "logProb1", "logProb2"):
        s.store(n, h[n])
    if "mixtureProb" in h:
        s.store("mixtureProb", h["mixtureProb"])
    return s


@pytest.mark.parametrize("p", [0.9, 0.05, 0.0])
def test_rc_consensus_shared_uniforms(ehmm, oracle, p):
    d, g, h, _ = _consensus_case(oracle)
    M = d["counts"].shape[0]
    u = oracle.RNG.set_seed(210423).runif(2 * M)
    ref, n_ref = oracle.consensusRejectionControlled(h, g["group"], p, uniforms=u, dense=True)
    with _open(ehmm, "t_rc", d, g, 2, h) as s:
        n = s.rc_consensus(p, uniforms=u)
        assert n == n_ref
        for k in range(2):
            b = s.rc_buckets(k)
            assert np.array_equal(b != 0, ref[k] != 0)            # identical support
            assert np.allclose(b, ref[k], rtol=1e-12, atol=0)
            tab = s.rc_table(k)
            ids = np.nonzero(ref[k])[0]
            assert np.array_equal(tab[:, 1], ids.astype(float))   # ascending group id, non-zero rows only
    if p == 0.0:
        assert n_ref == 0


def test_rc_consensus_device_rng(ehmm, oracle):
    """uniforms drawn on the device from R's Mersenne-Twister state: same draws, same final state"""
    d, g, h, _ = _consensus_case(oracle, M=30011)
    rng = oracle.RNG.set_seed(42)
    rng.runif(1000)                      # start from mti in the middle of a block
    st0 = rng.state()
    ref, n_ref = oracle.consensusRejectionControlled(h, g["group"], 0.05, rng=rng, dense=True)
    assert n_ref > 2 * 624
    with _open(ehmm, "t_rc_rng", d, g, 2, h) as s:
        s.rng_set_state(st0)
        n = s.rc_consensus(0.05)
        assert n == n_ref
        assert np.array_equal(s.rng_get_state(), rng.state())
        for k in range(2):
            assert np.allclose(s.rc_buckets(k), ref[k], rtol=1e-12, atol=0)
        # a second call continues the stream exactly where R would
        ref2, n2 = oracle.consensusRejectionControlled(h, g["group"], 0.3, rng=rng, dense=True)
        assert s.rc_consensus(0.3) == n2
        assert np.array_equal(s.rng_get_state(), rng.state())
        assert np.allclose(s.rc_buckets(1), ref2[1], rtol=1e-12, atol=0)


def test_rc_uniform_buffer_too_short(ehmm, oracle):
    d, g, h, _ = _consensus_case(oracle, M=5000)
    with _open(ehmm, "t_rc_short", d, g, 2, h) as s:
        with pytest.raises(ehmm.EhmmError):
            s.rc_consensus(0.5, uniforms=np.full(3, 0.5))


@pytest.mark.parametrize("p", [0.9, 0.05])
def test_rc_differential(ehmm, oracle, p):
    d, g, h, _ = _differential_case(oracle)
    M, N = d["counts"].shape
    B = d["B"]
    u = oracle.RNG.set_seed(7).runif((B + 2) * M)
    ref, n_ref = oracle.differentialRejectionControlled(h, g["group"], p, N, uniforms=u, dense=True)
    with _open(ehmm, "t_rcd", d, g, 3, h, design=d["design"]) as s:
        n = s.rc_differential(p, uniforms=u)
        assert n == n_ref
        for c in range(B + 2):
            b = s.rc_buckets(c)
            assert np.array_equal(b != 0, ref[c] != 0)
            assert np.allclose(b, ref[c], rtol=1e-12, atol=0)


@pytest.mark.parametrize("n_cond,n_rep", [(5, 1), (4, 2), (5, 2)])
def test_rc_differential_many_columns(ehmm, oracle, n_cond, n_rep):
    """B = 30 (BASELINE configs[3]: 32 columns, handled eight at a time) and B = 14 (16 columns): tables, support and
    draw count against the oracle on the same posteriors and the same uniform buffer"""
    d, g, h, _ = _differential_case(oracle, M=6011, n_cond=n_cond, n_rep=n_rep, seed=31 + n_cond)
    M, N = d["counts"].shape
    B = d["B"]
    u = oracle.RNG.set_seed(11).runif((B + 2) * M)
    for p in (0.9, 0.05):
        ref, n_ref = oracle.differentialRejectionControlled(h, g["group"], p, N, uniforms=u, dense=True)
        with _open(ehmm, "t_rcd_many", d, g, 3, h, design=d["design"]) as s:
            assert s.rc_differential(p, uniforms=u) == n_ref
            for c in range(B + 2):
                b = s.rc_buckets(c)
                assert np.array_equal(b != 0, ref[c] != 0), c
                assert np.allclose(b, ref[c], rtol=1e-12, atol=0), c


def _exact_tables(weights, group_cols, G):
    """correctly rounded sums by group (math.fsum): what the exact accumulators must return"""
    import math
    out = np.zeros(G + 1)
    gi = np.concatenate([np.asarray(g, dtype=np.int64) for g in group_cols])
    w = np.tile(weights, len(group_cols))
    order = np.argsort(gi, kind="stable")
    gi, w = gi[order], w[order]
    cuts = np.flatnonzero(np.diff(gi)) + 1
    for seg_g, seg_w in zip(gi[np.concatenate([[0], cuts])], np.split(w, cuts)):
        out[seg_g] = math.fsum(seg_w.tolist())
    return out


@pytest.mark.parametrize("model", ["consensus", "differential"])
def test_rc_tables_are_exact_and_reproducible(ehmm, oracle, model):
    """The rejection tables are accumulated as exact fixed-point integers (csrc/rc.cu): every entry is
    the correctly rounded sum of the thinned weights of its group -- bit-equal to math.fsum -- whatever
    order the hardware adds them in, so repeated calls give identical bits.  Few groups on purpose: every
    accumulator is hit by hundreds of concurrent atomics."""
    p = 0.05
    if model == "consensus":
        d, g, h, _ = _consensus_case(oracle, M=40009)
        K, design, cols = 2, None, 2
    else:
        d, g, h, _ = _differential_case(oracle, M=12007)
        K, design, cols = 3, d["design"], d["B"] + 2
    M, N = d["counts"].shape
    grp = np.asfortranarray((g["group"] - 1) % 37 + 1).astype(np.int32)      # 37 hot groups
    G = 37
    u = oracle.RNG.set_seed(99).runif(cols * M)
    with ehmm.Session("t_rc_exact", M, K) as s:
        s.set_data(d["counts"], d["offsets"])
        if design is not None:
            s.set_design(design)
        udsg = None if design is None else np.asfortranarray(np.zeros((G, d["B"]), dtype=np.uint8))
        s.set_groups(grp, np.zeros(G), np.zeros(G), None, udsg)
        for n in ("logLikelihood", "logFP", "logBP", "logProb1", "logProb2"):
            s.store(n, h[n])
        if design is not None:
            s.store("mixtureProb", h["mixtureProb"])
        P1 = s.marginal_probability()            # exp(logProb1) with the device's exp: the kernels' own inputs
        first = None
        for rep in range(3):
            n = s.rc_differential(p, uniforms=u) if design is not None else s.rc_consensus(p, uniforms=u)
            tabs = np.stack([s.rc_buckets(c) for c in range(cols)])
            if first is None:
                first, n0 = tabs, n
            assert n == n0 and np.array_equal(tabs, first)          # same bits every time
        at = 0
        for c in range(cols):
            if design is None:
                x = P1[:, c].copy()
            else:
                x = P1[:, 0].copy() if c == 0 else (P1[:, 2].copy() if c == cols - 1 else P1[:, 1] * h["mixtureProb"][:, c - 1])
            w, used = oracle.reweight(x, p, uniforms=u[at:])
            at += used
            gcols = [grp[:, 0]] if design is None else [grp[:, i] for i in range(N)]   # consensus: replicate 1 only (Q6)
            assert np.array_equal(first[c], _exact_tables(w, gcols, G)), c
        assert at == n0


def test_reweight_and_aggregate_helpers(ehmm, oracle):
    rng = np.random.default_rng(3)
    x = rng.random(50001) ** 4
    x[::97] = 0.0
    u = oracle.RNG.set_seed(1).runif(x.size)
    ref, n_ref = oracle.reweight(x, 0.3, uniforms=u)
    got, n = ehmm.reweight(x, 0.3, u)
    assert n == n_ref and np.array_equal(got, ref)
    f = rng.integers(1, 400, x.size)
    a_ref, a = oracle.aggregate(ref, f), ehmm.aggregate(ref, f)
    assert np.array_equal(a, a_ref)    # the reference's in-order sums, bit for bit (src/aggregate.cpp:23-30)
    # empty input
    assert ehmm.aggregate(np.zeros(0), np.zeros(0, dtype=np.int32)).shape == (0, 2)


def _scale_nb(par, y, x, off, w):
    """per gradient entry of derivNB (R/base-optim.R:120-130): the sum of the absolute values of the pieces
    that are added up -- the yardstick for a relative tolerance on a quantity that cancels to ~0"""
    from scipy.special import digamma
    par = np.asarray(par, dtype=float)
    beta, phi = par[:-1], par[-1]
    mu = np.exp(x @ beta + off)
    den = 1.0 + phi * mu
    sb = (w * np.abs((y - mu) / den))[:, None] * np.abs(x)
    pieces = (np.abs(np.log(den)) + np.abs(phi * (y - mu) / den) + np.abs(digamma(y + 1 / phi)) + abs(digamma(1 / phi))) / phi ** 2
    return np.concatenate([sb.sum(0), [np.sum(w * pieces)]])


def _scale_mix(par, id_, w, y, off, dsg, ncol):
    """the same yardstick for derivMixNB (R/base-optim.R:216-255), par = (b, db, l, dl)"""
    from scipy.special import digamma
    b, db, l, dl = par
    D = np.where(id_ == 1, 0, np.where(id_ == ncol, 1, dsg[np.arange(id_.size), np.clip(id_ - 2, 0, dsg.shape[1] - 1)])).astype(float)
    mu, sz = np.exp(b + db * D + off), np.exp(l + dl * D)
    a = w * (np.abs(y) + np.abs((y + sz) / (mu + sz) * mu))
    c = w * sz * (np.abs(digamma(y + sz)) + np.abs(digamma(sz)) + 1 + np.abs(np.log(sz)) + (y + sz) / (mu + sz) + np.abs(np.log(mu + sz)))
    return np.array([a.sum(), (a * D).sum(), c.sum(), (c * D).sum()])


@pytest.mark.parametrize("controls", [False, True])
def test_glm_nb(ehmm, oracle, controls):
    d, g, h, psi = _consensus_case(oracle, controls=controls)
    M = d["counts"].shape[0]
    u = oracle.RNG.set_seed(5).runif(2 * M)
    ref, _ = oracle.consensusRejectionControlled(h, g["group"], 0.05, uniforms=u, dense=True)
    with _open(ehmm, "t_glm", d, g, 2, h) as s:
        s.rc_consensus(0.05, uniforms=u)
        for k in range(2):
            ids = np.nonzero(ref[k])[0]
            y, off, w = g["u_chip"][ids - 1], g["u_offsets"][ids - 1], ref[k][ids]
            x = np.ones((ids.size, 1)) if not controls else np.column_stack([np.ones(ids.size), g["u_controls"][ids - 1]])
            for par in ([psi[k][0]] + ([psi[k][1]] if controls else []) + [1.0 / psi[k][-1]],
                        [0.3] + ([-0.2] if controls else []) + [0.01], [2.0] + ([0.4] if controls else []) + [1.7]):
                v_ref, g_ref = oracle.glmNB(par, y, x, off, w)
                v, gr = s.glm_nb(k, par)
                assert abs(v - v_ref) <= 1e-12 * abs(v_ref), (k, par)
                # gradient <= 1e-12 relative -- to the size of what is summed (the entries
                # themselves cancel to ~0 at the optimum)
                assert np.all(np.abs(gr - g_ref) <= 1e-12 * _scale_nb(par, y, x, off, w)), (k, par, gr, g_ref)


def test_glm_mix_nb(ehmm, oracle):
    d, g, h, delta = _differential_case(oracle)
    M, N = d["counts"].shape
    B = d["B"]
    u = oracle.RNG.set_seed(9).runif((B + 2) * M)
    ref, _ = oracle.differentialRejectionControlled(h, g["group"], 0.05, N, uniforms=u, dense=True)
    # the long table rbindlist(rcWeights, idcol='id') of R/base-EM.R:204-209
    rows = [(c + 1, np.nonzero(ref[c])[0]) for c in range(B + 2)]
    id_ = np.concatenate([np.full(ix.size, c) for c, ix in rows])
    gi = np.concatenate([ix for _, ix in rows])
    w = np.concatenate([ref[c - 1][ix] for c, ix in rows])
    y, off, dsg = g["u_chip"][gi - 1], g["u_offsets"][gi - 1], g["u_dsg"][gi - 1]
    with _open(ehmm, "t_mix", d, g, 3, h, design=d["design"]) as s:
        s.rc_differential(0.05, uniforms=u)
        for par in ([np.log(2), np.log(6), np.log(3), np.log(4 / 3)], [0.2, 0.9, 0.5, -0.3], [1.5, 0.1, 3.0, 1.0]):
            v_ref, g_ref = oracle.glmMixNB(par, id_, w, y, off, dsg, B + 2, MINZERO)
            v, gr = s.glm_mix_nb(par, MINZERO)
            assert abs(v - v_ref) <= 1e-12 * abs(v_ref), par
            assert np.all(np.abs(gr - g_ref) <= 1e-12 * _scale_mix(par, id_, w, y, off, dsg, B + 2)), (par, gr, g_ref)


@pytest.mark.parametrize("K", [2, 3])
@pytest.mark.parametrize("M", [1, 2, 3, 2047, 2048, 2049, 50001, 300007])
def test_viterbi_bit_exact(ehmm, oracle, K, M):
    """same logFP bits in -> identical state sequence out (LOGV reaches ~1e12 at M = 3e5, so late
    decisions are taken at coarse ulp granularity)"""
    rng = np.random.default_rng(100 * K + M % 97)
    gam = synth.GAMMA2 if K == 2 else synth.GAMMA3
    z = synth.markov_chain(M, gam, rng)
    logf = rng.normal(-6.0, 2.0, (M, K))
    logf[np.arange(M), z] += 4.0
    th = synth.THETA_CONSENSUS if K == 2 else synth.THETA_DIFFERENTIAL
    h = oracle.expStep(th["pi"], th["gamma"], logf)
    ref = oracle.computeViterbiSequence(h, th["pi"], th["gamma"])
    with ehmm.Session("t_vit", M, K) as s:
        s.store("logFP", h["logFP"])
        got = s.viterbi(th["pi"], th["gamma"])
        assert np.array_equal(got, ref)
        assert np.array_equal(s.fetch("viterbi")[:, 0], ref)


def test_viterbi_end_to_end_small(ehmm, oracle):
    """at test sizes the GPU's own logFP (scan order differs from the reference's sequential
    recursion in the last bits) still decodes to the reference's path"""
    M, K = 7000, 3
    d = synth.make_differential(M, 3, 2, seed=31)
    th = synth.THETA_DIFFERENTIAL
    delta = np.full(d["B"], 1.0 / d["B"])
    logf = oracle.loglikDifferentialHMM(d["counts"], d["offsets"], d["design"], delta, th["psi"], MINZERO)
    h = oracle.expStep(th["pi"], th["gamma"], logf)
    ref = oracle.computeViterbiSequence(h, th["pi"], th["gamma"])
    with ehmm.Session("t_vit2", M, K) as s:
        s.expstep(th["pi"], th["gamma"], logf)
        got = s.viterbi(th["pi"], th["gamma"])
    assert np.mean(got != ref) <= 1e-3


def test_rc_device_rng_parallel_substreams(ehmm, oracle):
    """more draws than one sub-stream holds (2^18 words): the jump-ahead windows must continue R's
    stream exactly, across calls, from any position inside a 624-word block"""
    M = 700001
    rng_np = np.random.default_rng(8)
    lp1 = np.log(np.column_stack([rng_np.random(M) * 0.04 + 1e-6, np.ones(M)]))
    lp1[:, 1] = np.log1p(-np.exp(lp1[:, 0]))
    lp1[::5, 0] = -np.inf                                      # zero posterior: no draw
    f = rng_np.integers(1, 5000, (M, 2)).astype(np.int32)
    h = {"logProb1": np.asfortranarray(lp1)}
    rng = oracle.RNG.set_seed(2024)
    rng.runif(333)
    with ehmm.Session("t_rc_par", M, 2) as s:
        s.set_data(np.zeros((M, 2), dtype=np.int32), np.zeros((M, 2)))
        s.set_groups(f, np.zeros(4999), np.zeros(4999))
        s.store("logProb1", lp1)
        s.rng_set_state(rng.state())
        for p in (0.05, 0.9, 0.05):
            ref, n_ref = oracle.consensusRejectionControlled(h, f, p, rng=rng, dense=True)
            n = s.rc_consensus(p)
            assert n == n_ref and n_ref > 2 ** 18
            assert np.array_equal(s.rng_get_state(), rng.state())
            for k in range(2):
                assert np.allclose(s.rc_buckets(k), ref[k], rtol=1e-12, atol=0)


@pytest.mark.parametrize("model", ["consensus", "consensus_ctrl", "differential"])
def test_build_groups_on_device_matches_grp(ehmm, model):
    """ehmm_build_groups == data.table's .GRP numbering (host reference: synth.make_groups)"""
    M = 50021
    if model == "differential":
        d = synth.make_differential(M, 3, 2, seed=81)
        ref = synth.make_groups(d["counts"], d["offsets"], design=d["design"])
        K = 3
    else:
        d = synth.make_consensus(M, 3, seed=82, controls=(model == "consensus_ctrl"))
        d["offsets"][7, 1] = -0.0     # -0.0 and +0.0 belong to one group
        d["offsets"][9, 1] = 0.0
        ref = synth.make_groups(d["counts"], d["offsets"], d["controls"])
        K = 2
    with ehmm.Session("t_grp", M, K) as s:
        s.set_data(d["counts"], d["offsets"], d.get("controls"))
        if model == "differential":
            s.set_design(d["design"])
        G = s.build_groups()
        got = s.groups_fetch()
        assert G == ref["G"]
        assert np.array_equal(got["group"], ref["group"])
        assert np.array_equal(got["u_chip"], ref["u_chip"]) and np.array_equal(got["u_offsets"], ref["u_offsets"])
        if model == "consensus_ctrl":
            assert np.array_equal(got["u_controls"], ref["u_controls"])
        if model == "differential":
            assert np.array_equal(got["u_dsg"], ref["u_dsg"])
        # the device-built dictionary drives the gather emission and the rejection control
        if K == 2:
            psi = [np.array([np.log(2.0), 0.3, 3.0]), np.array([np.log(12.0), 0.1, 4.0])] if model == "consensus_ctrl" \
                else synth.THETA_CONSENSUS["psi"]
            s.profile(True)
            s.loglik_consensus(psi, MINZERO)
            assert s.profile_read()["emis_gather"][0] == 1


def test_prefetched_group_upload_equals_direct(ehmm):
    """ehmm_prefetch_groups: the staged copy is what set_data_encoded[16] installs, for both id widths,
    over several rounds (two staging buffers in turn), and an unrelated pointer falls back to a direct copy"""
    import torch
    M, N = 40013, 3
    d = synth.make_consensus(M, N, seed=91)
    g = synth.make_groups(d["counts"], d["offsets"])
    th = synth.THETA_CONSENSUS
    with ehmm.Session("t_pref", M, 2) as s:
        s.set_data(d["counts"], d["offsets"])
        s.set_groups(g["group"], g["u_chip"], g["u_offsets"])
        s.loglik_consensus(th["psi"], MINZERO)
        want = s.fetch("logLikelihood")
        for dtype in (np.uint16, np.int32):
            if dtype == np.uint16 and g["G"] >= 65536:
                continue
            bufs = []
            for k in range(3):                       # three different tables: the same ids, permuted bins
                perm = np.random.default_rng(k).permutation(M)
                t = torch.empty(M * N, dtype=torch.uint16 if dtype == np.uint16 else torch.int32).pin_memory()
                a = t.numpy()
                a[:] = g["group"][perm].astype(dtype).ravel(order="F")
                bufs.append((t, a, perm))
            s.prefetch_groups(bufs[0][1])
            for k, (t, a, perm) in enumerate(bufs):
                s.set_data_encoded(a, g["u_chip"], g["u_offsets"])
                if k + 1 < len(bufs):
                    s.prefetch_groups(bufs[k + 1][1])
                s.loglik_consensus(th["psi"], MINZERO)
                assert np.array_equal(s.fetch("logLikelihood"), want[perm]), (dtype, k)
            # not prefetched: direct copy
            s.set_data_encoded(g["group"].astype(dtype), g["u_chip"], g["u_offsets"])
            s.loglik_consensus(th["psi"], MINZERO)
            assert np.array_equal(s.fetch("logLikelihood"), want)
