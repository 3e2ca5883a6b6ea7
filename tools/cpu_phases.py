"""Per-phase wall-clock of the CPU path (oracle restatement of the reference, in memory, 1 thread) on a
bounded sample of the bench workload, as SURVEY.md section 8d specifies: median of >= 5 after one
warm-up; plus the emission split over all cores as a generous bound.  Writes one JSON line."""
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from oracle import oracle as O
from oracle_backend import OracleBackend
from epigrahmm_b200 import em, synth

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
REP = 5
CFG = bench.CONFIGS["c2"]          # BASELINE configs[1]: the consensus workload of bench.py --config c2
N_SAMPLES, SEED = CFG["N"], CFG["seed"]
d = bench.make_data(CFG, M, SEED)
g = synth.make_groups(d["counts"], d["offsets"])
th = synth.THETA_CONSENSUS
ctl = em.controlEM()
be = OracleBackend(d["counts"], d["offsets"], 2, groups=g, rng=O.RNG.set_seed(SEED))


def med(f):
    f()
    ts = []
    for _ in range(REP):
        t = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t)
    return statistics.median(ts)


ncores = os.cpu_count()
ph = {}
ph["emission"] = med(lambda: be.loglik_consensus(th["psi"], ctl["minZero"]))
# the emission alone over all cores: bin slices in threads (the C call releases the GIL); NOT how the
# reference runs, only the generous bound SURVEY.md 8d asks for
from concurrent.futures import ThreadPoolExecutor
cuts = np.linspace(0, M, ncores + 1).astype(int)
slices = [(np.asfortranarray(d["counts"][a:b]), np.asfortranarray(d["offsets"][a:b])) for a, b in zip(cuts[:-1], cuts[1:])]
pool = ThreadPoolExecutor(ncores)
ph["emission_all_cores"] = med(lambda: list(pool.map(lambda cs: O.loglikConsensusHMM(cs[0], cs[1], th["psi"], ctl["minZero"]), slices)))
ph["expStep"] = med(lambda: be.expstep(th["pi"], th["gamma"]))
ph["maxStepProb"] = med(lambda: be.maxstepprob())
ph["computeQFunction+computeBIC"] = med(lambda: (be.qfunction(th["pi"], th["gamma"]), be.bic(5, N_SAMPLES)))
ph["consensusRejectionControlled"] = med(lambda: be.rc_consensus(bench.P_CUT))
ph["optimNB (both states)"] = med(lambda: em.optimNB_all(th["psi"], be, ctl))
ph["computeViterbiSequence"] = med(lambda: be.viterbi(th["pi"], th["gamma"]))
it = sum(v for k, v in ph.items() if k not in ("emission_all_cores", "computeViterbiSequence"))
print(json.dumps({"bins": M, "samples": N_SAMPLES, "states": 2, "threads": 1, "host_cores": ncores,
                  "seconds": {k: round(v, 4) for k, v in ph.items()}, "em_iteration_seconds": round(it, 4),
                  "bin_states_per_s": M * 2 / it,
                  "note": "oracle restatement of the reference, in memory (no HDF5 round trips, which flatters the reference); "
                          "emission_all_cores is the emission alone split over all cores"}))
