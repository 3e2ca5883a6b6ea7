import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import epigrahmm_b200 as E
from epigrahmm_b200 import em, synth, core
from oracle import oracle as O
from oracle_backend import OracleBackend
a = np.load("tests/golden/helas3_counts.npz")["counts"]
counts = np.asfortranarray(a[:, 2:4]); offsets = np.zeros(counts.shape, order="F")
M, N = counts.shape
score = np.log1p(counts); peaks = (score > np.quantile(score, 0.75, axis=0)).astype(float); z = 1 + (peaks.sum(1) > 0)
ctl = em.controlEM(maxIterEM=12)
theta = dict(pi=np.array([0.999, 0.001]), gamma=em.estimateTransitionProb(z, 2), psi=core.estimateCoefficients(z, counts, offsets, ctl, "consensus"))
g = synth.make_groups(counts, offsets)
nev = []
class Cnt(OracleBackend):
    def glm_nb(self, c, p, grad=True):
        nev.append(c); return super().glm_nb(c, p, grad)
be = Cnt(counts, offsets, 2, groups=g, rng=O.RNG.set_seed(210423))
fo = em.emConsensus(theta, be, N, ctl)
print("oracle evals", len(nev))
for rep in range(4):
    with E.Session("h", M, 2) as s:
        s.set_data(counts, offsets); s.build_groups(); s.rng_set_state(O.RNG.set_seed(210423).state())
        cnt = [0]
        orig = s.glm_nb_batch
        def gb(cols, pars, orig=orig):
            cnt[0] += len(cols); return orig(cols, pars)
        s.glm_nb_batch = gb
        fg = em.emConsensus(theta, s, N, ctl)
        d = [float(np.max(np.abs(em.unlist(a) - em.unlist(b)) / np.abs(em.unlist(b)))) for a, b in zip(fg["parHist"], fo["parHist"])]
        print(rep, "evals", cnt[0], ["%.1e" % x for x in d])
