"""Time computeViterbiSequence on the GPU (parallel scan and sequential kernel) and the oracle on one chain.
usage: python tools/viterbi_timing.py [M] [K]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import epigrahmm_b200 as E  # noqa: E402
from oracle import oracle as O  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 15_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(1)
base = -np.cumsum(rng.gamma(2.0, 4.0, M))
lf = np.asfortranarray(base[:, None] - rng.gamma(1.0, 3.0, (M, K)))
pi = np.full(K, 1.0 / K)
g = np.full((K, K), 0.004 / (K - 1)) + np.eye(K) * (0.996 - 0.004 / (K - 1))
out = {"bins": M, "states": K}
with E.Session("vt", M, K, 0) as s:
    s.store("logFP", lf)
    for mode, name in ((0, "scan"), (1, "sequential")):
        s.viterbi_mode(mode)
        s.viterbi(pi, g, fetch=False)
        s.profile(True)
        t0 = time.perf_counter()
        s.viterbi(pi, g, fetch=False)
        out[name + "_wall_ms"] = 1e3 * (time.perf_counter() - t0)
        prof = s.profile_read()
        s.profile(False)
        out[name + "_kernels_ms"] = {k: round(v[1], 4) for k, v in prof.items() if v[0]}
        out[name + "_info"] = s.viterbi_info()
        out[name] = s.fetch("viterbi").ravel()
t0 = time.perf_counter()
ref = O.computeViterbiSequence({"logFP": lf}, pi, g)
out["oracle_ms"] = 1e3 * (time.perf_counter() - t0)
out["scan_equals_oracle"] = bool(np.array_equal(out.pop("scan"), ref))
out["sequential_equals_oracle"] = bool(np.array_equal(out.pop("sequential"), ref))
print(json.dumps(out))
