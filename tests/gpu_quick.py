"""scratch driver: quick timing of the E-step kernels on one GPU (not a benchmark)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import epigrahmm_b200 as E
from epigrahmm_b200 import synth

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
for K in (2, 3):
    rng = np.random.default_rng(K)
    logf = rng.normal(-6, 2, (M, K))
    th = synth.THETA_CONSENSUS if K == 2 else synth.THETA_DIFFERENTIAL
    s = E.Session("quick", M, K)
    s.expstep(th["pi"], th["gamma"], logf)
    s.sync()
    st = torch.cuda.ExternalStream(s.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(10):
            s.expstep(th["pi"], th["gamma"])
        e1.record(st)
    s.sync()
    ms = e0.elapsed_time(e1) / 10
    s.profile(True)
    for _ in range(5):
        s.expstep(th["pi"], th["gamma"])
    print({k: round(v[1] / 5, 4) for k, v in s.profile_read().items() if v[0]})
    byt = M * (48 * K + 8 * K * K)
    print("K=%d M=%d expStep %.3f ms  %.1f GB/s (algorithmic %d B/bin)  ll=%.6f" % (K, M, ms, byt / ms / 1e6, 48 * K + 8 * K * K, s.loglik()))
    s.close()
