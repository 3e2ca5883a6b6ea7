"""Wall-clock per host call of one EM iteration (where the non-kernel time goes)."""
import sys
import time


sys.path.insert(0, ".")
import bench
import epigrahmm_b200 as E
from epigrahmm_b200 import em, synth
from epigrahmm_b200.rng import RState

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 15_000_000
d, g = bench.make_workload(M, bench.N_SAMPLES, bench.SEED)
th = synth.THETA_CONSENSUS
theta = dict(pi=th["pi"], gamma=th["gamma"], psi=th["psi"])
ctl = em.controlEM()
s = E.Session("tl", M, 2)
s.set_data(d["counts"], d["offsets"])
s.set_groups(g["group"], g["u_chip"], g["u_offsets"])
s.rng_set_state(RState(1).state)
acc = {}
nev = [0]
orig = s.glm_nb


def glm(c, p, grad=True):
    nev[0] += 1
    return orig(c, p, grad)


s.glm_nb = glm
for it in range(25):
    t = [time.perf_counter()]
    s.loglik_consensus(theta["psi"], ctl["minZero"]); t.append(time.perf_counter())
    s.expstep(theta["pi"], theta["gamma"]); t.append(time.perf_counter())
    s.maxstepprob(); t.append(time.perf_counter())
    s.rc_consensus(0.05); t.append(time.perf_counter())
    o0 = em.optimNB(theta["psi"][0], s, 0, ctl); t.append(time.perf_counter())
    o1 = em.optimNB(theta["psi"][1], s, 1, ctl); t.append(time.perf_counter())
    s.bic(5, 4); s.qfunction(theta["pi"], theta["gamma"]); t.append(time.perf_counter())
    if it >= 5:
        for n, a, b in zip(["loglik(launch)", "expstep(launch)", "maxstepprob(sync)", "rc", "optim0", "optim1", "bic+q"], t[:-1], t[1:]):
            acc[n] = acc.get(n, 0.0) + (b - a) / 20
print({k: round(v * 1e3, 3) for k, v in acc.items()}, "total ms", round(sum(acc.values()) * 1e3, 3), "glm evals/iter", nev[0] / 25)
s.close()
