#!/usr/bin/env python
"""bench.py -- EM iterations/s and bin-states/s of the epigraHMM EM hot path on B200.

A step is one full EM iteration of the consensus model (R/base-EM.R:106-175) at fixed theta:
NB emission -> expStep (forward-backward + posteriors) -> maxStepProb -> rejection control
(p = 0.05, uniforms from R's Mersenne-Twister generated on the device) -> L-BFGS-B over the NB-GLM
value/gradient reductions for both states -> BIC and Q.

Workload at N = 1: BASELINE.json configs[1] -- synthetic human genome, 200 bp bins
(15 000 000 bins) x 4 replicates, consensus peak calling (K = 2), generated as SURVEY.md 8(d).
At N > 1 the genome is sharded as contiguous bin blocks of 15 000 000 bins per GPU (weak scaling);
the single global chain is kept exact by exchanging the blocks' scan operators and all-reducing
the sufficient statistics (torch.distributed / NCCL).

  value : bin-states/s with the count matrix already resident in HBM
  e2e   : the same iteration through the host-pointer C ABI: every step uploads counts, offsets
          and group ids from pinned host memory and reads the updated parameters back
  --impl reference : the reference's CPU path (the oracle restatement; R is not installable
          here) on a bounded sample of the same workload, single-threaded like the reference
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_FULL = 15_000_000
N_SAMPLES = 4
K = 2
MINZERO = 2.2250738585072014e-308
P_CUT = 0.05
SEED = 20261002
CPU_SAMPLE_BINS = 1_000_000
METRIC = "bin-states/s"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region: NVML polled from a
    thread every ~2 ms (the timed region is milliseconds long, far below nvidia-smi's own loop
    period); falls back to `nvidia-smi -lms` (B200_PROFILING.md's clocks line) without NVML."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = threading.Event()
        self.th = None
        self.proc = None
        self.how = None

    def _nvml_loop(self, nv, h):
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def _smi_loop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.proc.stdout:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                self.sm.append(float(f[0]))
                self.mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(n)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self.how = "nvml"
            self.th = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.th.start()
            return
        except Exception:
            pass
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi"
            self.th = threading.Thread(target=self._smi_loop, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def stop(self):
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.th is not None:
            self.th.join(timeout=2)
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.how}


def pin_workload(d, g):
    import torch
    if g["G"] < 65536:
        g["group16"] = g["group"].astype(np.uint16)
    tdt = {np.dtype(np.int32): torch.int32, np.dtype(np.float64): torch.float64, np.dtype(np.uint16): torch.uint16}
    for k, src in (("counts", d), ("offsets", d), ("group", g), ("group16", g)):
        if k not in src:
            continue
        a = src[k]
        t = torch.empty(a.size, dtype=tdt[a.dtype]).pin_memory()
        v = t.numpy().reshape(a.shape, order="F")
        v[...] = a
        src[k] = v
        src["_pin_" + k] = t  # keep the pinned tensor alive
    return d, g


def make_workload(M, N, seed, pinned=False):
    from epigrahmm_b200 import synth
    d = synth.make_consensus(M, N, seed)
    g = synth.make_groups(d["counts"], d["offsets"])
    return pin_workload(d, g) if pinned else (d, g)


def em_iteration(be, theta, control, N, upload=None, p=None):
    """one EM iteration at theta (the loop body of emConsensus); returns the updated parameters"""
    from epigrahmm_b200 import em
    if upload is not None:
        d, g, encoded = upload[:3]
        if encoded:   # dt$Group + dtUnique only: 2 B per cell when G < 65536, else 4 B
            be.set_data_encoded(g.get("group16", g["group"]), g["u_chip"], g["u_offsets"])
            if len(upload) > 3 and upload[3]:   # the next step's table starts uploading while this one computes
                be.prefetch_groups(g.get("group16", g["group"]))
        else:         # the full long table: counts, offsets and Group, 16 B per cell
            be.set_data(d["counts"], d["offsets"])
            be.set_groups(g["group"], g["u_chip"], g["u_offsets"])
    be.loglik_consensus(theta["psi"], control["minZero"])
    be.expstep(theta["pi"], theta["gamma"])
    new = be.maxstepprob()
    be.rc_consensus(P_CUT if p is None else p)
    opt = em.optimNB_all(theta["psi"], be, control)
    new["psi"] = [o["par"] for o in opt]
    numPar = em.unlist(theta).size - 3
    new["bic"] = be.bic(numPar, N)
    new["q"] = be.qfunction(theta["pi"], theta["gamma"])
    return new


ALGO_BYTES = {  # algorithmic bytes per bin of each streaming kernel
    "fb_apply": 32 * K + 8 * K * K,          # logf in; logFP, logBP, logProb1, logProb2 out
    "fb_reduce": 8 * K,                      # logf in
    "emis_consensus": 12 * N_SAMPLES + 8 * K,  # int32 counts + fp64 offsets in; logf out
    "emis_gather": 4 * N_SAMPLES + 8 * K,    # dictionary-encoded emission: int32 group ids in; logf out
    "rc_count": 8 * K,                       # logProb1 in
    "rc_apply": 8 * K + 4,                   # logProb1 + group id in (uniforms / buckets are O(flagged))
}


def run_reference(args):
    """CPU arm: the oracle restatement of the reference path, 1 thread, bounded sample."""
    from oracle import oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_backend import OracleBackend
    from epigrahmm_b200 import em, synth
    M = CPU_SAMPLE_BINS
    d, g = make_workload(M, N_SAMPLES, SEED)
    th = synth.THETA_CONSENSUS
    theta = dict(pi=th["pi"], gamma=th["gamma"], psi=th["psi"])
    ctl = em.controlEM()
    be = OracleBackend(d["counts"], d["offsets"], K, groups=g, rng=O.RNG.set_seed(SEED))
    for _ in range(args.warmup):
        em_iteration(be, theta, ctl, N_SAMPLES)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        em_iteration(be, theta, ctl, N_SAMPLES)
    dt = time.perf_counter() - t0
    val = M * K * args.steps / dt
    sample = "first %d bins of the workload (%.1f%% of one GPU's block), one EM iteration per step" % (M, 100.0 * M / M_FULL)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "bin-states/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.gpus, reference=True),
            "cpu_baseline": {"value": val, "unit": "bin-states/s", "cores": 1, "kind": "port", "sample": sample,
                             "host_cores_available": os.cpu_count()},
            "e2e": {"value": val, "unit": "bin-states/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "em_iterations_per_s": args.steps / dt}
    print(json.dumps(line))


def config_dict(n_gpus, reference=False):
    c = {"workload": "BASELINE configs[1]: synthetic human genome 200 bp bins, consensus broad-peak calling, "
                     "one EM iteration per step at fixed theta (SURVEY.md 8d), rejection cut p=0.05",
         "bins_per_gpu": M_FULL, "samples": N_SAMPLES, "states": K, "seed": SEED,
         "l2_policy": "inputs larger than L2 (0.72 GB counts+offsets, 1.0 GB of E-step outputs per iteration vs 126 MB L2)",
         "sharding": "contiguous bin blocks, one per GPU" if n_gpus > 1 else "single GPU"}
    if reference:
        c["cpu_sample_bins"] = CPU_SAMPLE_BINS
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bins", type=int, default=M_FULL, help="bins per GPU (default: the configs[1] size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return

    # libraries (NCCL's version banner, ...) may write to fd 1: keep it for the one JSON line
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import epigrahmm_b200 as E
    from epigrahmm_b200 import em, synth
    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    M = args.bins
    th = synth.THETA_CONSENSUS
    theta = dict(pi=th["pi"], gamma=th["gamma"], psi=th["psi"])
    ctl = em.controlEM()

    if world > 1:
        from epigrahmm_b200 import sharded
        s = sharded.ShardedSession("bench", M, K, local, dist)
        d = synth.make_consensus(M, N_SAMPLES, SEED + rank)
        # one dt$Group dictionary for the whole sharded table
        g = sharded.global_groups(d["counts"], d["offsets"], s.comm, s.m0, s.Mtot)
        d, g = pin_workload(d, g)
    else:
        d, g = make_workload(M, N_SAMPLES, SEED + rank, pinned=True)
        s = E.Session("bench", M, K, local)
    s.set_data(d["counts"], d["offsets"])
    s.set_groups(g["group"], g["u_chip"], g["u_offsets"])
    s.rng_set_state(__import__("epigrahmm_b200.rng", fromlist=["RState"]).RState(SEED).state)
    stream = torch.cuda.ExternalStream(s.stream, device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, upload, p=None, pipelined=False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        t0 = time.perf_counter()
        out = None
        if pipelined:   # every upload is issued inside the timed region; the first one is not overlapped
            s.prefetch_groups(upload[1].get("group16", upload[1]["group"]))
        for k in range(nsteps):
            up = upload if not pipelined else tuple(upload) + (k + 1 < nsteps,)
            out = em_iteration(s, theta, ctl, N_SAMPLES * 1, upload=up, p=p)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall, out

    # ---- device-resident arm ----
    for _ in range(args.warmup):
        em_iteration(s, theta, ctl, N_SAMPLES)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    s.profile(True)
    ms, wall, out = timed(args.steps, None)
    prof = s.profile_read()
    s.profile(False)
    clk = clocks.stop() if rank == 0 else None
    # the first EM iteration's regime (rejectionCut = 0.9: most bins are thinned, ~M uniforms per state)
    n09 = max(3, min(args.steps, 10))
    for _ in range(2):
        em_iteration(s, theta, ctl, N_SAMPLES, p=0.9)
    ms_p09, _, _ = timed(n09, None, p=0.9)
    # ---- end-to-end arm: host buffers through the C ABI every step ----
    # (a) the table as the reference holds it (counts, offsets, Group); (b) dictionary-encoded
    # (Group + dtUnique), which is all the device needs
    for _ in range(2):
        em_iteration(s, theta, ctl, N_SAMPLES, upload=(d, g, False))
    ms_raw, _, _ = timed(args.steps, (d, g, False))
    for _ in range(2):
        em_iteration(s, theta, ctl, N_SAMPLES, upload=(d, g, True))
    ms_seq, _, _ = timed(args.steps, (d, g, True))
    # the same with the next step's upload overlapping the current step's iteration (ehmm_prefetch_groups);
    # untimed warm-up first (it allocates the two staging buffers)
    timed(3, (d, g, True), pipelined=True)
    ms_e2e, wall_e2e, _ = timed(args.steps, (d, g, True), pipelined=True)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    units = M * K * world * args.steps
    value = units / (ms * 1e-3)
    e2e_value = units / (ms_e2e * 1e-3)
    launches = sum(c for c, _ in prof.values())
    dom = max(prof.items(), key=lambda kv: kv[1][1])
    pk = peaks()
    peak_gbs = pk["hbm_gbs"] if pk else 6650.0
    roof = None
    if dom[0] in ALGO_BYTES and dom[1][0] > 0:
        per_launch = ALGO_BYTES[dom[0]] * M
        dur_ms = dom[1][1] / dom[1][0]
        ach = per_launch / (dur_ms * 1e-3) / 1e9
        traffic = None
        try:  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if tj.get("bins") == M:
                traffic = tj["traffic_bytes_per_launch"].get(dom[0])
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
                "frac": ach / peak_gbs, "traffic": traffic, "peak_source": "measured (MEASURED_PEAKS.json)" if pk else "fallback",
                "algorithmic_bytes_per_launch": per_launch, "avg_launch_ms": dur_ms,
                "share_of_step": dom[1][1] / ms}
    kernels = {}
    for name, (cnt, tms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        if cnt:
            k = {"launches_per_step": cnt / args.steps, "ms_per_step": tms / args.steps}
            if name in ALGO_BYTES and tms > 0:
                k["achieved_gbs"] = ALGO_BYTES[name] * M * cnt / (tms * 1e-3) / 1e9
                k["frac_of_hbm_peak"] = k["achieved_gbs"] / peak_gbs
            kernels[name] = k
    h2d_raw = d["counts"].nbytes + d["offsets"].nbytes + g["group"].nbytes + g["u_chip"].nbytes + g["u_offsets"].nbytes
    h2d = g.get("group16", g["group"]).nbytes + g["u_chip"].nbytes + g["u_offsets"].nbytes
    line = {"metric": METRIC, "value": value, "unit": "bin-states/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(world),
            "em_iterations_per_s": args.steps / (ms * 1e-3), "wall_ms_per_step": 1e3 * wall / args.steps,
            "clocks": clk, "gpu_launches": launches,
            "first_iteration_regime": {"rejection_cut": 0.9, "ms_per_step": ms_p09 / n09, "steps": n09,
                                       "value": M * K * world * n09 / (ms_p09 * 1e-3), "unit": "bin-states/s"},
            "e2e": {"value": e2e_value, "unit": "bin-states/s", "h2d_bytes_per_step": int(h2d) * 1,
                    "d2h_bytes_per_step": 8 * (2 + 4 + 4 + 2) + 8 * 12, "ms_per_step": ms_e2e / args.steps,
                    "upload": "dt$Group (uint16 when G < 65536, else int32) + dtUnique from pinned host memory every step (ehmm_set_data_encoded[16]); the copy of step k+1 is started right after step k's table is in place (ehmm_prefetch_groups) and overlaps step k's EM iteration; all uploads are issued inside the timed region",
                    "unpipelined": {"value": units / (ms_seq * 1e-3), "ms_per_step": ms_seq / args.steps},
                    "full_table_upload": {"value": units / (ms_raw * 1e-3), "ms_per_step": ms_raw / args.steps,
                                          "h2d_bytes_per_step": int(h2d_raw),
                                          "upload": "counts (int32) + offsets (fp64) + Group (int32) every step"}},
            "roofline": roof, "kernels": kernels,
            "result_check": {"bic": out["bic"], "q": out["q"], "pi": list(map(float, out["pi"]))}}
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle as O
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_backend import OracleBackend
        Ms = min(CPU_SAMPLE_BINS, M)
        dc = {"counts": np.asfortranarray(d["counts"][:Ms]), "offsets": np.asfortranarray(d["offsets"][:Ms])}
        gc = synth.make_groups(dc["counts"], dc["offsets"])
        be = OracleBackend(dc["counts"], dc["offsets"], K, groups=gc, rng=O.RNG.set_seed(SEED))
        t0 = time.perf_counter()
        em_iteration(be, theta, ctl, N_SAMPLES)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": Ms * K / dt, "unit": "bin-states/s", "cores": 1, "kind": "port",
                                "sample": "one EM iteration over the first %d bins of the workload (%.1f s), oracle restatement "
                                          "of the reference, in memory (no HDF5 round-trips), 1 thread as the reference" % (Ms, dt),
                                "host_cores_available": os.cpu_count()}
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    s.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
