"""One outer EM iteration of the differential model (3 conditions x 2 replicates, B = 6) at a given
size, with per-kernel CUDA-event times: BASELINE configs[2] when M = 15e6."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import epigrahmm_b200 as E
from epigrahmm_b200 import em, synth
from epigrahmm_b200.rng import RState

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 15_000_000
n_cond, n_rep = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3, 2)
t0 = time.time()
d = synth.make_differential(M, n_cond, n_rep, seed=20261003)
g = synth.make_groups(d["counts"], d["offsets"], design=d["design"])
print("data %.1f s, G = %d, B = %d" % (time.time() - t0, g["G"], d["B"]), file=sys.stderr)
th = synth.THETA_DIFFERENTIAL
theta = dict(pi=th["pi"], gamma=th["gamma"], delta=np.full(d["B"], 1.0 / d["B"]), psi=[th["psi"][:2].copy(), th["psi"][2:].copy()])
ctl = em.controlEM(maxIterEM=1)
s = E.Session("diff", M, 3)
s.set_data(d["counts"], d["offsets"])
s.set_design(d["design"])
s.set_groups(g["group"], g["u_chip"], g["u_offsets"], None, g["u_dsg"])
s.rng_set_state(RState(3).state)
N = n_cond * n_rep
for it in range(3):
    if it == 2:
        s.profile(True)
    s.sync()
    t0 = time.perf_counter()
    fit = em.outerEMDifferential(theta, s, N, ctl)
    s.sync()
    dt = time.perf_counter() - t0
prof = s.profile_read()
out = {"M": M, "N": N, "B": int(d["B"]), "G": g["G"], "outer_iteration_ms": dt * 1e3,
       "bin_states_per_s": M * 3 / dt,
       "kernels_ms": {k: [v[0], round(v[1], 3)] for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]) if v[0]},
       "theta_new_delta": [float(x) for x in fit["theta_new"]["delta"]]}
print(json.dumps(out))
s.close()
