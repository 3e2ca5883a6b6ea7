"""Regenerates the committed golden fixtures.

  helas3_counts.npz   the 9 994 x 6 count matrix of the reference's bundled data/helas3.rda
                      (xz-compressed RDX3; read here without R by locating the INTSXP payload).
                      Needs /root/reference; column sums are checked against SURVEY.md section 8c.
  golden_zinb.npz     the same for dist = 'zinb' (consensus and differential emission, mixture posteriors,
                      glmZINB / glmMixZINB value and gradient).
  golden_small.npz    seeded small inputs (epigrahmm_b200.synth) and the CPU oracle's outputs for
                      every function on the hot path.  The reference holds no numeric golden vectors
                      and cannot be executed in this image (no R), so these pin the ORACLE: the CPU
                      tests assert the oracle still reproduces them and the GPU tests compare the
                      CUDA path against them.

Run from the repository root:  python tests/golden/make_golden.py
"""
import lzma
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
MINZERO = 2.2250738585072014e-308


def helas3():
    path = "/root/reference/data/helas3.rda"
    if not os.path.exists(path):
        print("skip helas3 (no reference tree)")
        return
    raw = lzma.open(path).read()
    n = 9994 * 6
    pat = struct.pack(">i", n)
    pos = [i for i in range(4, len(raw) - 4) if raw[i:i + 4] == pat and (struct.unpack(">i", raw[i - 4:i])[0] & 0xFF) == 13]
    assert len(pos) == 1
    a = np.frombuffer(raw[pos[0] + 4:pos[0] + 4 + 4 * n], dtype=">i4").astype(np.int32).reshape(9994, 6, order="F")
    assert a.sum(0).tolist() == [104862, 75257, 86302, 59193, 91121, 85604] and a.max() == 191
    np.savez_compressed(os.path.join(HERE, "helas3_counts.npz"), counts=a,
                        samples=np.array(["EZH2.1", "EZH2.2", "H3K27me3.1", "H3K27me3.2", "H3K36me3.1", "H3K36me3.2"]))


def small():
    from oracle import oracle as O
    from epigrahmm_b200 import synth
    out = {}
    # consensus, M = 1500, N = 2
    d = synth.make_consensus(1500, 2, seed=20261001)
    g = synth.make_groups(d["counts"], d["offsets"])
    th = synth.THETA_CONSENSUS
    logf = O.loglikConsensusHMM(d["counts"], d["offsets"], th["psi"], MINZERO)
    h = O.expStep(th["pi"], th["gamma"], logf)
    m = O.maxStepProb(h)
    u = O.RNG.set_seed(210423).runif(3000)
    rc, nu = O.consensusRejectionControlled(h, g["group"], 0.05, uniforms=u, dense=True)
    ids = np.nonzero(rc[1])[0]
    v, gr = O.glmNB([th["psi"][1][0], 1 / th["psi"][1][1]], g["u_chip"][ids - 1], np.ones((ids.size, 1)),
                    g["u_offsets"][ids - 1], rc[1][ids])
    out.update(c_counts=d["counts"], c_offsets=d["offsets"], c_group=g["group"], c_u_chip=g["u_chip"],
               c_u_offsets=g["u_offsets"], c_logf=logf, c_logFP=h["logFP"], c_logBP=h["logBP"],
               c_logProb1=h["logProb1"], c_logProb2=h["logProb2"], c_pi=m["pi"], c_gamma=m["gamma"],
               c_q=O.computeQFunction(h, th["pi"], th["gamma"]), c_bic=O.computeBIC(h, 5, 2),
               c_viterbi=O.computeViterbiSequence(h, th["pi"], th["gamma"]), c_unif=u, c_rc=rc, c_rc_n=nu,
               c_glm_v=v, c_glm_g=gr)
    # differential, M = 700, 3 conditions x 2 replicates
    d = synth.make_differential(700, 3, 2, seed=20261003)
    g = synth.make_groups(d["counts"], d["offsets"], design=d["design"])
    th = synth.THETA_DIFFERENTIAL
    delta = np.full(d["B"], 1.0 / d["B"])
    logf = O.loglikDifferentialHMM(d["counts"], d["offsets"], d["design"], delta, th["psi"], MINZERO)
    h = O.expStep(th["pi"], th["gamma"], logf)
    h["mixtureProb"] = O.innerExpStep(d["counts"], d["offsets"], d["design"], delta, th["psi"], MINZERO)
    m = O.maxStepProb(h)
    u = O.RNG.set_seed(210423).runif(8 * 700)
    rc, nu = O.differentialRejectionControlled(h, g["group"], 0.05, 6, uniforms=u, dense=True)
    out.update(d_counts=d["counts"], d_offsets=d["offsets"], d_design=d["design"], d_group=g["group"],
               d_u_chip=g["u_chip"], d_u_offsets=g["u_offsets"], d_u_dsg=g["u_dsg"], d_logf=logf,
               d_logFP=h["logFP"], d_logBP=h["logBP"], d_logProb1=h["logProb1"], d_logProb2=h["logProb2"],
               d_eta=h["mixtureProb"], d_delta=O.innerMaxStepProb(h), d_pi=m["pi"], d_gamma=m["gamma"],
               d_q=O.computeQFunction(h, th["pi"], th["gamma"]), d_bic=O.computeBIC(h, 9, 6),
               d_viterbi=O.computeViterbiSequence(h, th["pi"], th["gamma"]), d_unif=u, d_rc=rc, d_rc_n=nu)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **out)


PSI_ZC = [np.array([-0.8, 0.7, 3.0]), np.array([2.4, 4.0])]                       # consensus: (lambda, beta, size), (beta, size)
PSI_ZD = np.array([-1.0, np.log(2.0), np.log(3.0), np.log(6.0), np.log(4.0 / 3.0)])    # differential: (lambda, b, l, db, dl)
PAR_GLM = np.array([0.6, 0.4, -0.7])                                                # glmZINB: (beta, phi, lambda)
PAR_MIX = np.array([-1.0, 0.7, 1.1, 2.5, 1.4])                                      # glmMixZINB: (lambda, b_bg, l_bg, b_enr, l_enr)


def zinb():
    from oracle import oracle as O
    from epigrahmm_b200 import synth
    out = {}
    d = synth.make_consensus(1200, 3, seed=20261011)
    rng = np.random.default_rng(20261012)
    counts = d["counts"].copy()
    counts[(rng.random(counts.shape) < 0.25) & (d["z"][:, None] == 0)] = 0
    counts = np.asfortranarray(counts)
    g = synth.make_groups(counts, d["offsets"])
    logf = O.loglikConsensusHMMzinb(counts, d["offsets"], PSI_ZC, MINZERO)
    w = rng.random(g["G"])
    v, gr = O.glmZINB(PAR_GLM, g["u_chip"], np.ones((g["G"], 1)), g["u_offsets"], w, MINZERO)
    out.update(c_counts=counts, c_offsets=d["offsets"], c_logf=logf, c_u_chip=g["u_chip"], c_u_offsets=g["u_offsets"],
               c_w=w, c_glm_v=v, c_glm_g=gr)
    d = synth.make_differential(600, 3, 2, seed=20261013)
    counts = d["counts"].copy()
    counts[(rng.random(counts.shape) < 0.25) & (d["z"][:, None] == 0)] = 0
    counts = np.asfortranarray(counts)
    delta = np.linspace(1.0, 2.0, d["B"])
    delta /= delta.sum()
    LL = O.loglikDifferentialHMMzinb(counts, d["offsets"], d["design"], delta, PSI_ZD, MINZERO)
    eta = O.innerExpStepZINB(counts, d["offsets"], d["design"], delta, PSI_ZD, MINZERO)
    n = 500
    idv = rng.integers(1, d["B"] + 3, n).astype(np.int32)
    yy = rng.negative_binomial(2, 0.3, n).astype(float)
    yy[rng.random(n) < 0.3] = 0
    off, ww = np.round(rng.normal(0, 0.1, n), 3), rng.random(n)
    dsg = rng.integers(0, 2, (n, d["B"])).astype(np.uint8)
    mv, mg = O.glmMixZINB(PAR_MIX, idv, ww, yy, off, dsg, d["B"] + 2, MINZERO)
    out.update(d_counts=counts, d_offsets=d["offsets"], d_design=d["design"], d_delta=delta, d_logf=LL, d_eta=eta,
               m_id=idv, m_y=yy, m_off=off, m_w=ww, m_dsg=dsg, m_v=mv, m_g=mg)
    np.savez_compressed(os.path.join(HERE, "golden_zinb.npz"), **out)


if __name__ == "__main__":
    helas3()
    small()
    zinb()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
