import sys, numpy as np
sys.path.insert(0, ".")
import epigrahmm_b200 as E
from epigrahmm_b200 import synth
M = 50021
d = synth.make_differential(M, 3, 2, seed=81)
ref = synth.make_groups(d["counts"], d["offsets"], design=d["design"])
bad = 0
for it in range(60):
    with E.Session("t_grp%d" % it, M, 3) as s:
        s.set_data(d["counts"], d["offsets"])
        s.set_design(d["design"])
        G = s.build_groups()
        got = s.groups_fetch()
        ok = G == ref["G"] and np.array_equal(got["group"], ref["group"]) and np.array_equal(got["u_chip"], ref["u_chip"]) \
            and np.array_equal(got["u_offsets"], ref["u_offsets"]) and np.array_equal(got["u_dsg"], ref["u_dsg"])
        if not ok:
            bad += 1
            print(it, "G", G, ref["G"], "group mism", int((got["group"] != ref["group"]).sum()), "dsg mism",
                  int((got["u_dsg"] != ref["u_dsg"]).sum()) if got["u_dsg"].shape == ref["u_dsg"].shape else "shape")
print("bad", bad, "of 60")
