"""Wall-clock per host call of one EM step of a bench configuration (where the non-kernel time goes).
usage: python tools/host_timeline.py [c2|c3|c5] [bins] [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import epigrahmm_b200 as E  # noqa: E402
from epigrahmm_b200.rng import RState  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = bench.CONFIGS[name]
M = int(float(sys.argv[2])) if len(sys.argv) > 2 else cfg["M"]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
d = bench.make_data(cfg, M, cfg["seed"])
theta, ctl = bench.theta_of(cfg, d), bench.control_of(cfg)
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:   # under torchrun: one bin block per rank (pass the bins per GPU as the second argument)
    import torch
    import torch.distributed as dist
    from epigrahmm_b200 import sharded
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = bench.make_data(cfg, M, cfg["seed"] + rank)
    s = sharded.ShardedSession("tl", M, cfg["K"], local, dist)
else:
    s = E.Session("tl", M, cfg["K"])
s.set_data(d["counts"], d["offsets"])
if cfg["model"] == "differential":
    s.set_design(d["design"])
if world > 1:
    sharded.global_groups_device(s.local, s.comm, s.m0, s.Mtot, design=d.get("design"))
else:
    s.build_groups()
s.rng_set_state(RState(1).state)

acc, cnt = {}, {}
on = [False]


def wrap(nm):
    f = getattr(s, nm)

    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        if on[0]:
            acc[nm] = acc.get(nm, 0.0) + time.perf_counter() - t0
            cnt[nm] = cnt.get(nm, 0) + 1
        return r
    setattr(s, nm, g)


def wrap_local(nm):   # the calls a ShardedSession makes on its block's session, and on the exchange
    for obj in (s.local, s.comm):
        if hasattr(obj, nm):
            f = getattr(obj, nm)

            def g(*a, _f=f, **k):
                t0 = time.perf_counter()
                r = _f(*a, **k)
                if on[0]:
                    acc["  ." + nm] = acc.get("  ." + nm, 0.0) + time.perf_counter() - t0
                    cnt["  ." + nm] = cnt.get("  ." + nm, 0) + 1
                return r
            setattr(obj, nm, g)


for nm in ("loglik_consensus", "loglik_differential_hmm", "expstep", "maxstepprob", "rc_consensus", "rc_differential",
           "inner_expstep", "inner_maxstepprob", "glm_nb_batch", "glm_mix_nb", "bic", "qfunction", "rc_hint", "expstep_mode"):
    wrap(nm)
if world > 1:
    for nm in ("expstep_reduce", "expstep_apply_ops", "estep_stats", "loglik", "rc_count", "rc_apply_at", "rc_finalize", "inner_mstep_sums",
               "all_gather", "all_reduce_sum", "all_reduce_device", "fetch_logprob1_row0", "estep_set_stats", "rc_buckets_device"):
        wrap_local(nm)
for _ in range(3):
    bench.em_step(s, cfg, theta, ctl)
s.sync()
on[0] = True
t0 = time.perf_counter()
for _ in range(steps):
    bench.em_step(s, cfg, theta, ctl)
s.sync()
tot = (time.perf_counter() - t0) / steps
out = {"config": name, "bins": M, "wall_ms_per_step": round(1e3 * tot, 3),
       "calls": {k: {"n_per_step": cnt[k] / steps, "ms_per_step": round(1e3 * acc[k] / steps, 3)} for k in sorted(acc, key=lambda k: -acc[k])}}
out["inside_calls_ms"] = round(sum(v["ms_per_step"] for k, v in out["calls"].items() if not k.startswith("  .")), 3)
out["python_between_calls_ms"] = round(out["wall_ms_per_step"] - out["inside_calls_ms"], 3)
if rank == 0:
    print(json.dumps(out))
s.close()
if world > 1:
    dist.destroy_process_group()
