"""Wall-clock per host call of one sharded EM iteration (run under torchrun)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import bench
from epigrahmm_b200 import em, sharded, synth
from epigrahmm_b200.rng import RState

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 15_000_000
s = sharded.ShardedSession("tl", M, 2, local, dist)
d = synth.make_consensus(M, bench.N_SAMPLES, bench.SEED + rank)
g = sharded.global_groups(d["counts"], d["offsets"], s.comm, s.m0, s.Mtot)
s.set_data(d["counts"], d["offsets"])
s.set_groups(g["group"], g["u_chip"], g["u_offsets"])
s.rng_set_state(RState(1).state)
th = synth.THETA_CONSENSUS
theta = dict(pi=th["pi"], gamma=th["gamma"], psi=th["psi"])
ctl = em.controlEM()
acc = {}
comm = s.comm
for name in ("all_gather", "all_reduce_sum", "broadcast", "all_reduce_device"):
    f = getattr(comm, name)

    def wrap(f=f, name=name):
        def w(*a, **k):
            t = time.perf_counter()
            r = f(*a, **k)
            acc["comm:" + name] = acc.get("comm:" + name, 0.0) + (time.perf_counter() - t)
            return r
        return w
    setattr(comm, name, wrap())
tot = {}
for it in range(25):
    if it == 5:
        acc.clear()
        tot.clear()
    t = [time.perf_counter()]
    s.loglik_consensus(theta["psi"], ctl["minZero"]); t.append(time.perf_counter())
    s.expstep(theta["pi"], theta["gamma"]); t.append(time.perf_counter())
    s.maxstepprob(); t.append(time.perf_counter())
    s.rc_consensus(0.05); t.append(time.perf_counter())
    em.optimNB_all(theta["psi"], s, ctl); t.append(time.perf_counter())
    s.bic(5, 4); s.qfunction(theta["pi"], theta["gamma"]); t.append(time.perf_counter())
    for n, a, b in zip(["loglik", "expstep", "maxstepprob", "rc", "optim", "bic+q"], t[:-1], t[1:]):
        tot[n] = tot.get(n, 0.0) + (b - a)
if rank == 0:
    print({k: round(v / 20 * 1e3, 3) for k, v in tot.items()}, "total", round(sum(tot.values()) / 20 * 1e3, 3))
    print({k: round(v / 20 * 1e3, 3) for k, v in acc.items()})
s.close()
dist.destroy_process_group()
