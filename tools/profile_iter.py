"""Short driver for ncu: a few EM iterations of the consensus workload at a given size."""
import sys


sys.path.insert(0, ".")
import bench
import epigrahmm_b200 as E
from epigrahmm_b200 import em, synth
from epigrahmm_b200.rng import RState

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d, g = bench.make_workload(M, bench.N_SAMPLES, bench.SEED)
th = synth.THETA_CONSENSUS
theta = dict(pi=th["pi"], gamma=th["gamma"], psi=th["psi"])
s = E.Session("prof", M, 2)
s.set_data(d["counts"], d["offsets"])
s.set_groups(g["group"], g["u_chip"], g["u_offsets"])
s.rng_set_state(RState(1).state)
for _ in range(iters):
    out = bench.em_iteration(s, theta, em.controlEM(), bench.N_SAMPLES)
print("ok", out["bic"])
s.close()
