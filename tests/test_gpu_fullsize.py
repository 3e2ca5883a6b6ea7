"""Function-boundary parity against the oracle at BASELINE's FULL sizes (SURVEY.md 8c protocol T1): configs[1]
(15 M bins x 4, consensus, K = 2) and configs[2] (15 M bins x 6, differential, K = 3, B = 6).  Every
function of the path gets the same inputs on both sides; what is compared and to what tolerance is
north_star's: posteriors <= 1e-8 absolute, log-likelihood / Q / BIC / pi / gamma <= 1e-9 relative,
rejection tables with a shared uniform buffer <= 1e-12 relative with identical support and draw count,
Viterbi on shared logFP bit-exact.

The reference's own E-step is an UNSCALED log-space recursion whose values reach -1e9 at this size, so its
posteriors carry rounding noise of order ulp(|logLik|).  The tests therefore also compute a
scaled long-double recursion (oracle.expStepTruth -- a yardstick, not the reference) and record the three
pairwise distances; the table is written to gpurun_out/accuracy_r02.json (committed as
profiles/accuracy_r02.json).
"""
import json
import os
import time

import numpy as np
import pytest

from epigrahmm_b200 import synth

pytestmark = pytest.mark.gpu
MINZERO = 2.2250738585072014e-308
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M_FULL = int(os.environ.get("EHMM_FULLSIZE_BINS", 15_000_000))
_TABLE = {}


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _dump(name, row):
    _TABLE[name] = row
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "accuracy_r02.json"), "w") as f:
        json.dump(_TABLE, f, indent=1, sort_keys=True)


def _posterior_distances(P_gpu, h, P_true, row):
    P_or = np.exp(h["logProb1"])
    row["posterior_max_abs_gpu_vs_oracle"] = float(np.max(np.abs(P_gpu - P_or)))
    row["posterior_max_abs_gpu_vs_scaled_truth"] = float(np.max(np.abs(P_gpu - P_true)))
    row["posterior_max_abs_oracle_vs_scaled_truth"] = float(np.max(np.abs(P_or - P_true)))
    row["posterior_frac_bins_gpu_vs_oracle_above_1e-9"] = float(np.mean(np.max(np.abs(P_gpu - P_or), axis=1) > 1e-9))


def _rc_compare(s, ref, n, n_ref, cols, row, key):
    assert n == n_ref                                                       # same number of uniforms consumed
    worst = 0.0
    for c in range(cols):
        b = s.rc_buckets(c)
        assert np.array_equal(b != 0, ref[c] != 0), c                        # identical support
        nz = ref[c] != 0
        worst = max(worst, _rel(b[nz], ref[c][nz]))
    assert worst <= 1e-12
    row[key + "_tables_max_rel"] = worst
    row[key + "_uniforms_consumed"] = int(n)


def test_c2_full_size_against_oracle(ehmm, oracle):
    M, N, K = M_FULL, 4, 2
    t0 = time.time()
    d = synth.make_consensus(M, N, seed=20261002)
    th = synth.THETA_CONSENSUS
    pi, gamma, psi = th["pi"], th["gamma"], th["psi"]
    row = {"config": "BASELINE configs[1]", "bins": M, "samples": N, "states": K}
    with ehmm.Session("full-c2", M, K) as s:
        s.set_data(d["counts"], d["offsets"])
        G = s.build_groups()
        # emission
        s.loglik_consensus(psi, MINZERO)
        logf = s.fetch("logLikelihood")
        logf_or = oracle.loglikConsensusHMM(d["counts"], d["offsets"], psi, MINZERO)
        row["emission_max_rel_vs_oracle"] = _rel(logf, logf_or)
        assert row["emission_max_rel_vs_oracle"] <= 1e-11
        del logf_or
        # expStep on the SAME logf
        s.expstep(pi, gamma)
        h = oracle.expStep(pi, gamma, logf)
        P_true, ll_true = oracle.expStepTruth(pi, gamma, logf)
        P_gpu = s.marginal_probability()
        _posterior_distances(P_gpu, h, P_true, row)
        ll_or = oracle.computeBIC(h, 0, 1) / -2.0
        row.update(loglik_gpu=s.loglik(), loglik_oracle=ll_or, loglik_scaled_truth=ll_true,
                   loglik_rel_gpu_vs_oracle=_rel(s.loglik(), ll_or), loglik_rel_gpu_vs_truth=_rel(s.loglik(), ll_true),
                   loglik_rel_oracle_vs_truth=_rel(ll_or, ll_true))
        assert row["posterior_max_abs_gpu_vs_scaled_truth"] <= 1e-10          # the scan itself is accurate ...
        assert row["posterior_max_abs_gpu_vs_oracle"] <= max(1e-8, 2 * row["posterior_max_abs_oracle_vs_scaled_truth"])
        assert row["loglik_rel_gpu_vs_oracle"] <= 1e-9 and row["loglik_rel_gpu_vs_truth"] <= 1e-12
        for n_ in ("logFP", "logBP"):
            row[n_ + "_max_rel_vs_oracle"] = _rel(s.fetch(n_), h[n_])
            assert row[n_ + "_max_rel_vs_oracle"] <= 1e-9
        P2 = np.exp(s.fetch("logProb2")[1:])
        row["joint_posterior_max_abs_vs_oracle"] = float(np.max(np.abs(P2 - np.exp(h["logProb2"][1:]))))
        assert row["joint_posterior_max_abs_vs_oracle"] <= max(1e-8, 2 * row["posterior_max_abs_oracle_vs_scaled_truth"])
        del P2
        # maxStepProb / Q / BIC
        mg, mo = s.maxstepprob(), oracle.maxStepProb(h)
        row["pi_gamma_max_rel_vs_oracle"] = max(_rel(mg["pi"], mo["pi"]), _rel(mg["gamma"], mo["gamma"]))
        row["Q_rel_vs_oracle"] = _rel(s.qfunction(pi, gamma), oracle.computeQFunction(h, pi, gamma))
        row["BIC_rel_vs_oracle"] = _rel(s.bic(4, N), oracle.computeBIC(h, 4, N))
        assert row["pi_gamma_max_rel_vs_oracle"] <= 1e-9 and row["Q_rel_vs_oracle"] <= 1e-9 and row["BIC_rel_vs_oracle"] <= 1e-9
        # rejection control on the ORACLE's posteriors with one shared uniform buffer
        grp = s.groups_fetch()["group"]
        s.store("logProb1", h["logProb1"])
        u = oracle.RNG.set_seed(20261002).runif(K * M)
        for p in (0.05, 0.9):
            ref, n_ref = oracle.consensusRejectionControlled(h, grp, p, uniforms=u, dense=True)
            n = s.rc_consensus(p, uniforms=u)
            _rc_compare(s, ref, n, n_ref, K, row, "rc_p%g" % p)
        del u, ref
        # Viterbi on the oracle's logFP: bit-exact
        s.store("logFP", h["logFP"])
        t1 = time.time()
        vg = s.viterbi(pi, gamma)
        row["viterbi_gpu_s"] = time.time() - t1
        t1 = time.time()
        vo = oracle.computeViterbiSequence(h, pi, gamma)
        row["viterbi_oracle_s"] = time.time() - t1
        row["viterbi_mismatches_shared_logFP"] = int(np.sum(vg != vo))
        assert row["viterbi_mismatches_shared_logFP"] == 0
        # end to end (the GPU's own logFP against the oracle's): how many decisions flip at ulp granularity
        s.expstep(pi, gamma, logf)
        ve = s.viterbi(pi, gamma)
        row["viterbi_mismatch_rate_end_to_end"] = float(np.mean(ve != vo))
        row["viterbi_vs_simulated_truth_gpu"] = float(np.mean(ve == d["z"]))
        row["viterbi_vs_simulated_truth_oracle"] = float(np.mean(vo == d["z"]))
    row["G"] = int(G)
    row["seconds"] = time.time() - t0
    _dump("c2", row)


def test_c3_full_size_against_oracle(ehmm, oracle):
    M, K = M_FULL, 3
    t0 = time.time()
    d = synth.make_differential(M, 3, 2, seed=20261003)
    N, B = d["counts"].shape[1], d["B"]
    th = synth.THETA_DIFFERENTIAL
    pi, gamma, psi = th["pi"], th["gamma"], th["psi"]
    delta = np.full(B, 1.0 / B)
    row = {"config": "BASELINE configs[2]", "bins": M, "samples": N, "states": K, "mixture_components": B}
    with ehmm.Session("full-c3", M, K) as s:
        s.set_data(d["counts"], d["offsets"])
        s.set_design(d["design"])
        G = s.build_groups()
        s.loglik_differential_hmm(delta, psi, MINZERO)
        logf = s.fetch("logLikelihood")
        logf_or = oracle.loglikDifferentialHMM(d["counts"], d["offsets"], d["design"], delta, psi, MINZERO)
        row["emission_max_rel_vs_oracle"] = _rel(logf, logf_or)
        assert row["emission_max_rel_vs_oracle"] <= 1e-10
        del logf_or
        s.expstep(pi, gamma)
        h = oracle.expStep(pi, gamma, logf)
        P_true, ll_true = oracle.expStepTruth(pi, gamma, logf)
        P_gpu = s.marginal_probability()
        _posterior_distances(P_gpu, h, P_true, row)
        ll_or = oracle.computeBIC(h, 0, 1) / -2.0
        row.update(loglik_gpu=s.loglik(), loglik_oracle=ll_or, loglik_scaled_truth=ll_true,
                   loglik_rel_gpu_vs_oracle=_rel(s.loglik(), ll_or), loglik_rel_gpu_vs_truth=_rel(s.loglik(), ll_true),
                   loglik_rel_oracle_vs_truth=_rel(ll_or, ll_true))
        assert row["posterior_max_abs_gpu_vs_scaled_truth"] <= 1e-10
        assert row["posterior_max_abs_gpu_vs_oracle"] <= max(1e-8, 2 * row["posterior_max_abs_oracle_vs_scaled_truth"])
        assert row["loglik_rel_gpu_vs_oracle"] <= 1e-9 and row["loglik_rel_gpu_vs_truth"] <= 1e-12
        mg, mo = s.maxstepprob(), oracle.maxStepProb(h)
        row["pi_gamma_max_rel_vs_oracle"] = max(_rel(mg["pi"], mo["pi"]), _rel(mg["gamma"], mo["gamma"]))
        row["Q_rel_vs_oracle"] = _rel(s.qfunction(pi, gamma), oracle.computeQFunction(h, pi, gamma))
        row["BIC_rel_vs_oracle"] = _rel(s.bic(10, N), oracle.computeBIC(h, 10, N))
        assert row["pi_gamma_max_rel_vs_oracle"] <= 1e-9 and row["Q_rel_vs_oracle"] <= 1e-9 and row["BIC_rel_vs_oracle"] <= 1e-9
        del P_gpu, P_true
        # mixture posteriors and delta on the oracle's posteriors
        s.store("logProb1", h["logProb1"])
        s.inner_expstep(delta, psi, MINZERO)
        eta = oracle.innerExpStep(d["counts"], d["offsets"], d["design"], delta, psi, MINZERO)
        row["mixtureProb_max_abs_vs_oracle"] = float(np.max(np.abs(s.fetch("mixtureProb") - eta)))
        assert row["mixtureProb_max_abs_vs_oracle"] <= 1e-9
        h["mixtureProb"] = eta
        row["delta_max_rel_vs_oracle"] = _rel(s.inner_maxstepprob(), oracle.innerMaxStepProb(h))
        assert row["delta_max_rel_vs_oracle"] <= 1e-9
        # rejection control: both sides from the oracle's posteriors and mixture posteriors, one uniform buffer
        grp = s.groups_fetch()["group"]
        s.store("mixtureProb", eta)
        u = oracle.RNG.set_seed(20261003).runif((B + 2) * M)
        ref, n_ref = oracle.differentialRejectionControlled(h, grp, 0.05, N, uniforms=u, dense=True)
        n = s.rc_differential(0.05, uniforms=u)
        _rc_compare(s, ref, n, n_ref, B + 2, row, "rc_p0.05")
        del u
        # glmMixNB over those tables
        gg = s.groups_fetch(with_group=False)
        rows_ = [(c + 1, np.nonzero(ref[c])[0]) for c in range(B + 2)]
        id_ = np.concatenate([np.full(ix.size, c) for c, ix in rows_])
        gi = np.concatenate([ix for _, ix in rows_])
        w = np.concatenate([ref[c - 1][ix] for c, ix in rows_])
        par = [psi[0], psi[2], psi[1], psi[3]]
        v_ref, g_ref = oracle.glmMixNB(par, id_, w, gg["u_chip"][gi - 1], gg["u_offsets"][gi - 1], gg["u_dsg"][gi - 1], B + 2, MINZERO)
        v, gr = s.glm_mix_nb(par, MINZERO)
        row["glmMixNB_value_rel_vs_oracle"] = _rel(v, v_ref)
        row["glmMixNB_grad_max_abs_over_value"] = float(np.max(np.abs(gr - g_ref)) / abs(v_ref))
        assert row["glmMixNB_value_rel_vs_oracle"] <= 1e-12 and row["glmMixNB_grad_max_abs_over_value"] <= 1e-12
        del ref, eta
        # Viterbi on the oracle's logFP
        s.store("logFP", h["logFP"])
        t1 = time.time()
        vg = s.viterbi(pi, gamma)
        row["viterbi_gpu_s"] = time.time() - t1
        t1 = time.time()
        vo = oracle.computeViterbiSequence(h, pi, gamma)
        row["viterbi_oracle_s"] = time.time() - t1
        row["viterbi_mismatches_shared_logFP"] = int(np.sum(vg != vo))
        assert row["viterbi_mismatches_shared_logFP"] == 0
        s.expstep(pi, gamma, logf)
        ve = s.viterbi(pi, gamma)
        row["viterbi_mismatch_rate_end_to_end"] = float(np.mean(ve != vo))
    row["G"] = int(G)
    row["seconds"] = time.time() - t0
    _dump("c3", row)
