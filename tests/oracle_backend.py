"""A backend for epigrahmm_b200.em that runs the CPU oracle behind the Session interface.

Test infrastructure: lets the parity tests run the SAME EM driver with the backend swapped
(protocol T2) and lets the host logic be tested without a GPU.
"""
import numpy as np

from oracle import oracle as O


class OracleBackend:
    def __init__(self, counts, offsets, K, controls=None, design=None, groups=None, rng=None):
        self.counts = np.asfortranarray(counts, dtype=np.int32)
        self.offsets = np.asfortranarray(offsets, dtype=np.float64)
        self.controls = controls
        self.design = design
        self.M, self.N = self.counts.shape
        self.K = K
        self.B = 0 if design is None else design.shape[0]
        self.g = groups
        self.rng = rng if rng is not None else O.RNG.set_seed(0)
        self.h = {}
        self.rc = None

    # emission
    def loglik_consensus(self, psi, minZero):
        self.h["logLikelihood"] = O.loglikConsensusHMM(self.counts, self.offsets, psi, minZero, self.controls)

    def loglik_consensus_zinb(self, psi, minZero):
        self.h["logLikelihood"] = O.loglikConsensusHMMzinb(self.counts, self.offsets, psi, minZero, self.controls)

    def loglik_init(self, psi):
        self.h["logLikelihood"] = O.loglikInit(self.counts[:, 0], self.offsets[:, 0], psi)

    def loglik_differential_hmm(self, delta, psi, minZero):
        self.h["logLikelihood"] = O.loglikDifferentialHMM(self.counts, self.offsets, self.design, delta, psi, minZero)

    def loglik_differential_hmm_zinb(self, delta, psi, minZero):
        self.h["logLikelihood"] = O.loglikDifferentialHMMzinb(self.counts, self.offsets, self.design, delta, psi, minZero)

    def inner_expstep_zinb(self, delta, psi, minZero):
        self.h["mixtureProb"] = O.innerExpStepZINB(self.counts, self.offsets, self.design, delta, psi, minZero)

    def inner_expstep(self, delta, psi, minZero):
        self.h["mixtureProb"] = O.innerExpStep(self.counts, self.offsets, self.design, delta, psi, minZero)

    # E/M
    def expstep(self, pi, gamma, logf=None):
        if logf is not None:
            self.h["logLikelihood"] = np.asfortranarray(logf)
        self.h.update(O.expStep(pi, gamma, self.h["logLikelihood"]))

    def maxstepprob(self):
        return O.maxStepProb(self.h)

    def qfunction(self, pi, gamma):
        return O.computeQFunction(self.h, pi, gamma)

    def bic(self, numPar, numSamples):
        return O.computeBIC(self.h, numPar, numSamples)

    def inner_maxstepprob(self):
        return O.innerMaxStepProb(self.h)

    def viterbi(self, pi, gamma, fetch=True):
        self.h["viterbi"] = O.computeViterbiSequence(self.h, pi, gamma).reshape(-1, 1)
        return self.h["viterbi"][:, 0]

    def loglik(self):
        return O.computeBIC(self.h, 0, 1) / -2.0

    def fetch(self, name):
        return self.h[name]

    # rejection control
    def rc_consensus(self, p, uniforms=None):
        self.rc, n = O.consensusRejectionControlled(self.h, self.g["group"], p, rng=None if uniforms is not None else self.rng,
                                                    uniforms=uniforms, dense=True)
        return n

    def rc_differential(self, p, uniforms=None):
        self.rc, n = O.differentialRejectionControlled(self.h, self.g["group"], p, self.N,
                                                       rng=None if uniforms is not None else self.rng,
                                                       uniforms=uniforms, dense=True)
        return n

    def rc_buckets(self, c):
        return self.rc[c]

    # GLM over the (sum, id) tables joined with dtUnique
    def glm_nb(self, column, par, grad=True):
        ids = np.nonzero(self.rc[column])[0]
        x = np.ones((ids.size, 1))
        if self.g.get("u_controls") is not None:
            x = np.column_stack([x[:, 0], self.g["u_controls"][ids - 1]])
        return O.glmNB(par, self.g["u_chip"][ids - 1], x, self.g["u_offsets"][ids - 1], self.rc[column][ids], grad)

    def glm_zinb(self, column, par, minZero, grad=True):
        ids = np.nonzero(self.rc[column])[0]
        x = np.ones((ids.size, 1))
        if self.g.get("u_controls") is not None:
            x = np.column_stack([x[:, 0], self.g["u_controls"][ids - 1]])
        return O.glmZINB(par, self.g["u_chip"][ids - 1], x, self.g["u_offsets"][ids - 1], self.rc[column][ids], minZero, grad)

    def _mix_rows(self):
        B = self.B
        rows = [(c + 1, np.nonzero(self.rc[c])[0]) for c in range(B + 2)]
        id_ = np.concatenate([np.full(ix.size, c) for c, ix in rows])
        gi = np.concatenate([ix for _, ix in rows])
        w = np.concatenate([self.rc[c - 1][ix] for c, ix in rows])
        return id_, gi, w

    def glm_mix_nb(self, par, minZero, grad=True):
        id_, gi, w = self._mix_rows()
        return O.glmMixNB(par, id_, w, self.g["u_chip"][gi - 1], self.g["u_offsets"][gi - 1], self.g["u_dsg"][gi - 1],
                          self.B + 2, minZero, grad)

    def glm_mix_zinb(self, par, minZero, grad=True):
        id_, gi, w = self._mix_rows()
        return O.glmMixZINB(par, id_, w, self.g["u_chip"][gi - 1], self.g["u_offsets"][gi - 1], self.g["u_dsg"][gi - 1],
                            self.B + 2, minZero, grad)
