"""tools/profile_step.py CONFIG [BINS] [STEPS] -- a few bench steps (bench.py's em_step on bench.py's workload)
without the bench's other arms: the command ncu wraps for the launch list and the --set full captures."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import epigrahmm_b200 as E  # noqa: E402
from epigrahmm_b200.rng import RState  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = bench.CONFIGS[name]
M = int(float(sys.argv[2])) if len(sys.argv) > 2 else cfg["M"]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
d = bench.make_data(cfg, M, cfg["seed"])
theta, ctl = bench.theta_of(cfg, d), bench.control_of(cfg)
s = E.Session("profile", M, cfg["K"])
s.set_data(d["counts"], d["offsets"])
if cfg["model"] == "differential":
    s.set_design(d["design"])
s.build_groups()
s.rng_set_state(RState(cfg["seed"]).state)
for it in range(steps + 1):
    if it == steps:
        s.profile(True)
    s.sync()
    t0 = time.perf_counter()
    out = bench.em_step(s, cfg, theta, ctl)
    s.sync()
    dt = time.perf_counter() - t0
prof = s.profile_read()
print(json.dumps({"config": name, "M": M, "step_ms": dt * 1e3, "bic": out["bic"],
                  "kernels_ms": {k: [v[0], round(v[1], 4)] for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]) if v[0]}}))
s.close()
