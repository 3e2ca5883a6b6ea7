"""dist = 'zinb' (SURVEY.md section 8f rank 3): zero-inflated background state of the consensus
model -- emission (R/base-loglikelihood.R:123-141), glmZINB/derivZINB (R/base-optim.R:136-181) and
the EM trajectory with optimNB(dist = 'zinb') for state 1 (R/base-EM.R:140), GPU vs oracle."""
import numpy as np
import pytest

from oracle_backend import OracleBackend

from epigrahmm_b200 import core, em, synth

pytestmark = pytest.mark.gpu
MINZERO = 2.2250738585072014e-308


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _zinb_data(M, N, seed, inflate=0.25, controls=False):
    d = synth.make_consensus(M, N, seed=seed)
    rng = np.random.default_rng(seed + 1)
    counts = d["counts"].copy()
    drop = (rng.random(counts.shape) < inflate) & (d["z"][:, None] == 0)   # structural zeros in the background
    counts[drop] = 0
    d["counts"] = np.asfortranarray(counts)
    d["controls"] = np.asfortranarray(np.log1p(rng.poisson(3.0, counts.shape).astype(float))) if controls else None
    return d


@pytest.mark.parametrize("controls", [False, True])
def test_zinb_emission_matches_oracle_on_both_paths(ehmm, oracle, controls):
    M, N = 20011, 3
    d = _zinb_data(M, N, 31, controls=controls)
    psi = [np.array([-0.8, 0.3, 0.7, -0.1, 3.0]), np.array([2.4, 0.05, 4.0])] if controls else \
        [np.array([-0.8, 0.7, 3.0]), np.array([2.4, 4.0])]
    ref = oracle.loglikConsensusHMMzinb(d["counts"], d["offsets"], psi, MINZERO, d["controls"])
    with ehmm.Session("zinb-e", M, 2) as s:
        s.set_data(d["counts"], d["offsets"], d["controls"])
        s.loglik_consensus_zinb(psi, MINZERO)                  # per-cell kernel
        cell = s.fetch("logLikelihood")
        assert _rel(cell, ref) <= 1e-12
        s.build_groups()
        s.loglik_consensus_zinb(psi, MINZERO)                  # dictionary-encoded path
        assert np.array_equal(s.fetch("logLikelihood"), cell)
        # an extreme zero-inflation predictor and a tiny size stay finite (minZero keeps the logs alive)
        hard = [psi[0].copy(), psi[1]]
        hard[0][0] = 30.0
        s.loglik_consensus_zinb(hard, MINZERO)
        got = s.fetch("logLikelihood")
        want = oracle.loglikConsensusHMMzinb(d["counts"], d["offsets"], hard, MINZERO, d["controls"])
        assert np.all(np.isfinite(got)) and _rel(got, want) <= 1e-9


@pytest.mark.parametrize("controls", [False, True])
def test_glm_zinb_value_and_gradient(ehmm, oracle, controls):
    M, N = 30011, 2
    d = _zinb_data(M, N, 47, controls=controls)
    g = synth.make_groups(d["counts"], d["offsets"], controls=d["controls"])
    th = synth.THETA_CONSENSUS
    be = OracleBackend(d["counts"], d["offsets"], 2, controls=d["controls"], groups=g, rng=oracle.RNG.set_seed(3))
    psi_nb = [np.array([0.7, 0.0, 3.0]), np.array([2.4, 0.0, 4.0])] if controls else th["psi"]
    be.loglik_consensus(psi_nb, MINZERO)
    be.expstep(th["pi"], th["gamma"])
    be.rc_consensus(0.05)
    with ehmm.Session("zinb-g", M, 2) as s:
        s.set_data(d["counts"], d["offsets"], d["controls"])
        s.set_groups(g["group"], g["u_chip"], g["u_offsets"], g.get("u_controls"))
        s.store("logProb1", be.fetch("logProb1"))
        s.rng_set_state(oracle.RNG.set_seed(3).state())
        s.rc_consensus(0.05)
        pars = [np.array([0.6, 0.05, 0.4, -0.7, 0.1]), np.array([1.5, -0.2, 2.0, 1.0, -0.3])] if controls else \
            [np.array([0.6, 0.4, -0.7]), np.array([1.5, 2.0, 1.0]), np.array([0.1, 0.001, -6.0])]
        for par in pars:
            v, gr = s.glm_zinb(0, par, MINZERO)
            vo, go = be.glm_zinb(0, par, MINZERO)
            assert abs(v - vo) <= 1e-11 * abs(vo), (par, v, vo)
            assert np.max(np.abs(gr - go)) <= 1e-9 * max(1.0, np.max(np.abs(go))), (par, gr, go)


def test_em_trajectory_zinb(ehmm, oracle):
    """same driver, backend swapped (T2) with dist = 'zinb'; gates as in tests/test_gpu_em.py"""
    M, N = 20011, 2
    d = _zinb_data(M, N, 53)
    g = synth.make_groups(d["counts"], d["offsets"])
    z = 1 + (d["z"] == 1)
    ctl = em.controlEM(maxIterEM=6)
    theta = dict(pi=np.array([0.999, 0.001]), gamma=em.estimateTransitionProb(z, 2),
                 psi=core.estimateCoefficients(z, d["counts"], d["offsets"], ctl, "consensus", "zinb"))
    assert theta["psi"][0].size == 3 and theta["psi"][1].size == 2
    rng = oracle.RNG.set_seed(99)
    be = OracleBackend(d["counts"], d["offsets"], 2, groups=g, rng=rng)
    fo = em.emConsensus(theta, be, N, ctl, dist="zinb")
    with ehmm.Session("zinb-em", M, 2) as s:
        s.set_data(d["counts"], d["offsets"])
        assert s.build_groups() == g["G"]
        s.rng_set_state(oracle.RNG.set_seed(99).state())
        fg = em.emConsensus(theta, s, N, ctl, dist="zinb")
        assert len(fg["parHist"]) == len(fo["parHist"])
        worst, tight = 0.0, 0
        for it, (pg, po) in enumerate(zip(fg["parHist"], fo["parHist"])):
            r = _rel(em.unlist(pg), em.unlist(po))
            worst, tight = max(worst, r), tight + (r <= 1e-9)
            assert r <= 1e-9, (it + 1, r)
            assert _rel(fg["controlHist"][it]["bic"], fo["controlHist"][it]["bic"]) <= max(1e-9, 10 * worst)
        assert tight >= len(fg["parHist"]) // 2
        print("zinb trajectory: worst relative parameter difference %.3g" % worst)
        th = fg["theta_new"]
        # the fit sees the structural zeros: the zero-inflation probability at offset 0 is well above the floor
        zi = 1.0 / (1.0 + np.exp(-th["psi"][0][0]))
        assert 0.1 < zi < 0.5, zi
        if worst <= 1e-9:
            assert np.array_equal(s.rng_get_state(), rng.state())
            to = fo["theta_new"]
            assert np.array_equal(s.viterbi(th["pi"], th["gamma"]), be.viterbi(to["pi"], to["gamma"]))


# ---- differential model with dist = 'zinb' -------------------------------------------------------

def _zinb_diff_data(M, n_cond, n_rep, seed, inflate=0.25):
    d = synth.make_differential(M, n_cond, n_rep, seed=seed)
    rng = np.random.default_rng(seed + 1)
    counts = d["counts"].copy()
    counts[(rng.random(counts.shape) < inflate) & (d["z"][:, None] == 0)] = 0
    d["counts"] = np.asfortranarray(counts)
    return d


PSI_Z = np.array([-1.0, np.log(2.0), np.log(3.0), np.log(6.0), np.log(4.0 / 3.0)])   # (lambda, b, l, db, dl)


def test_differential_zinb_emission_and_mixture_posteriors(ehmm, oracle):
    M, nc, nr = 9001, 3, 2
    d = _zinb_diff_data(M, nc, nr, 61)
    B = d["B"]
    delta = np.linspace(1, 2, B)
    delta /= delta.sum()
    LLo = oracle.loglikDifferentialHMMzinb(d["counts"], d["offsets"], d["design"], delta, PSI_Z, MINZERO)
    eta_o = oracle.innerExpStepZINB(d["counts"], d["offsets"], d["design"], delta, PSI_Z, MINZERO)
    with ehmm.Session("zinb-d", M, 3) as s:
        s.set_data(d["counts"], d["offsets"])
        s.set_design(d["design"])
        for grouped in (False, True):
            if grouped:
                s.build_groups()
            s.loglik_differential_hmm_zinb(delta, PSI_Z, MINZERO)
            assert _rel(s.fetch("logLikelihood"), LLo) <= 1e-10, grouped
            s.inner_expstep_zinb(delta, PSI_Z, MINZERO)
            assert np.max(np.abs(s.fetch("mixtureProb") - eta_o)) <= 1e-9, grouped


def test_glm_mix_zinb_value_and_gradient(ehmm, oracle):
    M, nc, nr = 9001, 3, 2
    d = _zinb_diff_data(M, nc, nr, 67)
    g = synth.make_groups(d["counts"], d["offsets"], design=d["design"])
    th = synth.THETA_DIFFERENTIAL
    B = d["B"]
    delta = np.full(B, 1.0 / B)
    be = OracleBackend(d["counts"], d["offsets"], 3, design=d["design"], groups=g, rng=oracle.RNG.set_seed(4))
    be.loglik_differential_hmm_zinb(delta, PSI_Z, MINZERO)
    be.expstep(th["pi"], th["gamma"])
    be.inner_expstep_zinb(delta, PSI_Z, MINZERO)
    be.rc_differential(0.05)
    with ehmm.Session("zinb-dg", M, 3) as s:
        s.set_data(d["counts"], d["offsets"])
        s.set_design(d["design"])
        s.set_groups(g["group"], g["u_chip"], g["u_offsets"], None, g["u_dsg"])
        s.store("logProb1", be.fetch("logProb1"))
        s.store("mixtureProb", be.fetch("mixtureProb"))
        s.rng_set_state(oracle.RNG.set_seed(4).state())
        s.rc_differential(0.05)
        for par in (np.array([-1.0, 0.7, 1.1, 2.5, 1.4]), np.array([0.5, 0.2, 0.1, 1.9, 2.0]), np.array([-4.0, 1.0, 3.0, 2.0, 0.5])):
            v, gr = s.glm_mix_zinb(par, MINZERO)
            vo, go = be.glm_mix_zinb(par, MINZERO)
            assert abs(v - vo) <= 1e-11 * abs(vo), (par, v, vo)
            assert np.max(np.abs(gr - go)) <= 1e-9 * max(1.0, np.max(np.abs(go))), (par, gr, go)


def test_em_trajectory_differential_zinb(ehmm, oracle):
    M, nc, nr = 8009, 2, 2
    d = _zinb_diff_data(M, nc, nr, 71)
    g = synth.make_groups(d["counts"], d["offsets"], design=d["design"])
    N, B = nc * nr, d["B"]
    th = synth.THETA_DIFFERENTIAL
    theta = dict(pi=th["pi"], gamma=th["gamma"], delta=np.full(B, 1.0 / B), psi=[PSI_Z[:3].copy(), PSI_Z[3:].copy()])
    ctl = em.controlEM(maxIterEM=3)
    rng = oracle.RNG.set_seed(8)
    be = OracleBackend(d["counts"], d["offsets"], 3, design=d["design"], groups=g, rng=rng)
    fo = em.outerEMDifferential(theta, be, N, ctl, dist="zinb")
    with ehmm.Session("zinb-dem", M, 3) as s:
        s.set_data(d["counts"], d["offsets"])
        s.set_design(d["design"])
        assert s.build_groups() == g["G"]
        s.rng_set_state(oracle.RNG.set_seed(8).state())
        fg = em.outerEMDifferential(theta, s, N, ctl, dist="zinb")
        worst, tight = 0.0, 0
        for pg, po in zip(fg["parHist"], fo["parHist"]):
            r = _rel(em.unlist(pg), em.unlist(po))
            worst, tight = max(worst, r), tight + (r <= 1e-9)
            assert r <= 1e-9, r
        assert tight >= len(fg["parHist"]) // 2
        print("differential zinb trajectory: worst relative parameter difference %.3g" % worst)
        assert fg["theta_new"]["psi"][0].size == 3 and fg["theta_new"]["psi"][1].size == 2
        zi = 1.0 / (1.0 + np.exp(-fg["theta_new"]["psi"][0][0]))
        assert 0.1 < zi < 0.5, zi


def test_cuda_path_against_committed_zinb_golden(ehmm):
    """tests/golden/golden_zinb.npz: the CUDA path against the committed oracle outputs (no oracle call here)"""
    import os
    import sys
    gold = os.path.join(os.path.dirname(__file__), "golden")
    sys.path.insert(0, gold)
    import make_golden as mg
    g = np.load(os.path.join(gold, "golden_zinb.npz"))
    M, N = g["c_counts"].shape
    with ehmm.Session("gz-c", M, 2) as s:
        s.set_data(g["c_counts"], g["c_offsets"])
        s.loglik_consensus_zinb(mg.PSI_ZC, MINZERO)
        assert _rel(s.fetch("logLikelihood"), g["c_logf"]) <= 1e-12
        # glmZINB over the golden table: one group per row, weights as the "rejection table" of column 0
        G = g["c_u_chip"].size
        grp = np.ones((M, N), dtype=np.int32, order="F")
        s.set_groups(grp, g["c_u_chip"], g["c_u_offsets"])
        s.store("logProb1", np.log(np.full((M, 2), 0.5)))
        s.rc_consensus(0.0)                                  # allocates the tables (p = 0: no thinning)
        ptr, n, integer = s.rc_buckets_device()
        assert not integer                                  # below the exact-accumulation cut: plain fp64 sums
        import torch
        class _B:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}
        t = torch.as_tensor(_B(), device="cuda")
        t.zero_()
        t[1:G + 1] = torch.as_tensor(g["c_w"], device="cuda")
        torch.cuda.synchronize()
        v, gr = s.glm_zinb(0, mg.PAR_GLM, MINZERO)
        assert abs(v - float(g["c_glm_v"])) <= 1e-11 * abs(float(g["c_glm_v"]))
        assert np.max(np.abs(gr - g["c_glm_g"])) <= 1e-9 * max(1.0, np.max(np.abs(g["c_glm_g"])))
    M, N = g["d_counts"].shape
    with ehmm.Session("gz-d", M, 3) as s:
        s.set_data(g["d_counts"], g["d_offsets"])
        s.set_design(g["d_design"])
        s.build_groups()
        s.loglik_differential_hmm_zinb(g["d_delta"], mg.PSI_ZD, MINZERO)
        assert _rel(s.fetch("logLikelihood"), g["d_logf"]) <= 1e-10
        s.inner_expstep_zinb(g["d_delta"], mg.PSI_ZD, MINZERO)
        assert np.max(np.abs(s.fetch("mixtureProb") - g["d_eta"])) <= 1e-9
