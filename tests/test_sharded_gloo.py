"""world_size-2 gloo test (CPU) of the bin-block sharding protocol in epigrahmm_b200.sharded:
carry composition for the E-step, all-reduced statistics, the global draw order of the rejection
control and the global group dictionary.  Each rank runs a numpy/oracle block backend; the merged
result must equal the single-process oracle on the whole chain."""
import os
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINZERO = 2.2250738585072014e-308


class BlockBackend:
    """host stand-in for a Session holding bins [m0, m0+M): same sharded methods, numpy + oracle"""

    def __init__(self, O, counts, offsets, K, rng_state):
        self.O, self.counts, self.offsets, self.K = O, counts, offsets, K
        self.M, self.N = counts.shape
        self.h = {}
        self.rng = O.RNG()
        self.rng.mti = int(rng_state[0])
        for i in range(624):
            self.rng.mt[i] = int(rng_state[1 + i])

    def set_block(self, m0, Mtot):
        self.m0, self.Mtot = m0, Mtot

    def set_groups(self, g):
        self.g = g

    def loglik_consensus(self, psi, minZero):
        self.h["logLikelihood"] = self.O.loglikConsensusHMM(self.counts, self.offsets, psi, minZero)

    def _ops(self, gamma):
        f = self.h["logLikelihood"]
        m = f.max(1)
        e = np.exp(f - m[:, None])
        return m, e

    def expstep_reduce(self, pi, gamma, logf=None):
        self.pi, self.gamma = np.asarray(pi, float), np.asarray(gamma, float)
        K = self.K
        m, e = self._ops(gamma)
        X, ls = np.eye(K), 0.0
        for j in range(self.M):
            T = np.diag(e[j]) if (self.m0 == 0 and j == 0) else self.gamma * e[j][None, :]
            X = X @ T
            ls += m[j]
            if j % 7 == 6:
                s = X.max()
                X, ls = X / s, ls + np.log(s)
        return np.append(X.ravel(), ls)

    def expstep_apply(self, fwd, bwd):
        K, M = self.K, self.M
        f = self.h["logLikelihood"]
        m, e = self._ops(self.gamma)
        lg = np.log(self.gamma)
        LF, LB = np.empty((M, K)), np.empty((M, K))
        v, s = np.array(fwd[:K]), float(fwd[K])
        prev_log = np.log(v) + s
        lprev = np.empty((M, K))
        for j in range(M):
            lprev[j] = prev_log
            a = v if (self.m0 == 0 and j == 0) else v @ self.gamma
            v = a * e[j]
            s += m[j]
            c = v.max()
            v, s = v / c, s + np.log(c)
            LF[j] = np.log(v) + s
            prev_log = LF[j]
        tail = np.append(v, s)
        u, s = np.array(bwd[:K]), float(bwd[K])
        for j in range(M - 1, -1, -1):
            LB[j] = np.log(u) + s
            u = self.gamma @ (e[j] * u)
            s += m[j]
            c = u.max()
            u, s = u / c, s + np.log(c)
        S = LF + LB
        L1 = S - np.logaddexp.reduce(S, axis=1, keepdims=True)
        L2 = np.empty((M, K * K))
        for i in range(K):
            for l in range(K):
                L2[:, i * K + l] = L1[:, l] + lprev[:, i] + lg[i, l] + f[:, l] - LF[:, l]
        if self.m0 == 0:
            L2[0] = -np.log(K * K)
        self.h.update(logFP=LF, logBP=LB, logProb1=L1, logProb2=L2)
        st = np.exp(L2[1:] if self.m0 == 0 else L2).sum(0)
        self._stats = np.append(st, np.sum(np.exp(L1) * f))
        return tail

    def estep_stats(self):
        return self._stats

    def fetch_logprob1_row0(self):
        return self.h["logProb1"][0].copy()

    def estep_set_stats(self, stats, row0, ll):
        self.gstats, self.grow0, self.ll = np.asarray(stats), np.asarray(row0), float(ll)

    def maxstepprob(self):
        K = self.K
        S = self.gstats[:K * K].reshape(K, K)
        gam = np.maximum(S / S.sum(1, keepdims=True), 1e-6)
        gam /= gam.sum(1, keepdims=True)
        p = np.maximum(np.exp(self.grow0), 1e-6)
        return {"pi": p / p.sum(), "gamma": gam}

    def rc_count(self, p, differential=False):
        x = np.exp(self.h["logProb1"])
        return np.array([np.count_nonzero((x[:, k] < p) & (x[:, k] / p != 0)) for k in range(self.K)], dtype=np.int64)

    def rc_apply_at(self, p, offs, total, differential=False):
        u = self.rng.runif(int(total))          # the one global stream; every rank advances by `total`
        x = np.exp(self.h["logProb1"])
        G = self.g["G"]
        self.buckets = np.zeros((self.K, G + 1))
        f = self.g["group"][:, 0]
        for k in range(self.K):
            w, n = self.O.reweight(x[:, k], p, uniforms=u[offs[k]:])
            np.add.at(self.buckets[k], f, w)

    def rc_buckets_host(self):
        return self.buckets

    def rc_buckets_set(self, b):
        self.buckets = b


def _worker(rank, world, initfile, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import oracle as O
    from epigrahmm_b200 import sharded, synth
    dist.init_process_group("gloo", init_method="file://" + initfile, rank=rank, world_size=world)
    M, N, K = 6001, 2, 2
    d = synth.make_consensus(M, N, seed=5)
    lo, hi = (0, 2500) if rank == 0 else (2500, M)          # uneven blocks
    counts, offsets = np.asfortranarray(d["counts"][lo:hi]), np.asfortranarray(d["offsets"][lo:hi])
    st0 = O.RNG.set_seed(77).state()
    be = BlockBackend(O, counts, offsets, K, st0)
    s = sharded.ShardedSession("t", hi - lo, K, 0, dist, local=be)
    assert (s.m0, s.Mtot) == (lo, M)
    g = sharded.global_groups(counts, offsets, s.comm, s.m0, s.Mtot)
    be.set_groups(g)
    th = synth.THETA_CONSENSUS
    be.loglik_consensus(th["psi"], MINZERO)
    s.expstep(th["pi"], th["gamma"])
    mp = be.maxstepprob()
    total = s.rc_consensus(0.05)
    np.savez(os.path.join(outdir, "r%d.npz" % rank), L1=be.h["logProb1"], L2=be.h["logProb2"], FP=be.h["logFP"],
             BP=be.h["logBP"], pi=mp["pi"], gamma=mp["gamma"], ll=be.ll, buckets=be.buckets, group=g["group"],
             G=g["G"], u_chip=g["u_chip"], u_off=g["u_offsets"], total=total, state=be.rng.state())
    dist.destroy_process_group()


def test_two_rank_protocol_matches_single_process(oracle):
    import torch.multiprocessing as mp
    from epigrahmm_b200 import synth
    with tempfile.TemporaryDirectory() as td:
        initfile = os.path.join(td, "init")
        mp.spawn(_worker, args=(2, initfile, td), nprocs=2, join=True)
        r = [np.load(os.path.join(td, "r%d.npz" % i)) for i in range(2)]
        M, N = 6001, 2
        d = synth.make_consensus(M, N, seed=5)
        th = synth.THETA_CONSENSUS
        logf = oracle.loglikConsensusHMM(d["counts"], d["offsets"], th["psi"], MINZERO)
        h = oracle.expStep(th["pi"], th["gamma"], logf)
        for name, key in (("logProb1", "L1"), ("logProb2", "L2")):
            got = np.vstack([r[0][key], r[1][key]])
            assert np.max(np.abs(np.exp(got) - np.exp(h[name]))) <= 1e-8, name
        for name, key in (("logFP", "FP"), ("logBP", "BP")):
            got = np.vstack([r[0][key], r[1][key]])
            assert np.max(np.abs(got - h[name]) / np.maximum(1, np.abs(h[name]))) <= 1e-9, name
        m = oracle.maxStepProb(h)
        for i in range(2):  # every rank ends with the same global parameters
            assert np.allclose(r[i]["pi"], m["pi"], rtol=1e-9) and np.allclose(r[i]["gamma"], m["gamma"], rtol=1e-9)
            ll = oracle.computeBIC(h, 0, 1) / -2
            assert abs(r[i]["ll"] - ll) <= 1e-9 * abs(ll)
        # the global group dictionary equals .GRP over the whole table
        g = synth.make_groups(d["counts"], d["offsets"])
        assert int(r[0]["G"]) == g["G"] == int(r[1]["G"])
        assert np.array_equal(np.vstack([r[0]["group"], r[1]["group"]]), g["group"])
        assert np.array_equal(r[0]["u_chip"], g["u_chip"]) and np.array_equal(r[1]["u_off"], g["u_offsets"])
        # rejection control: same draws as the single-process reference, same final RNG state
        rng = oracle.RNG.set_seed(77)
        ref, n = oracle.consensusRejectionControlled(h, g["group"], 0.05, rng=rng, dense=True)
        assert int(r[0]["total"]) == n == int(r[1]["total"])
        for i in range(2):
            assert np.array_equal(r[i]["buckets"] != 0, ref != 0)
            # (the numpy block backend's posteriors differ from the oracle's at 1e-11; draws are identical)
            assert np.allclose(r[i]["buckets"], ref, rtol=1e-8, atol=0)
            assert np.array_equal(r[i]["state"], rng.state())


def test_rc_offsets_and_carries():
    from epigrahmm_b200.sharded import combine_carries, rc_offsets
    offs, tot = rc_offsets(np.array([[3, 1], [2, 5], [4, 0]]), 1)
    assert tot == 15 and offs.tolist() == [3, 9 + 1]
    rng = np.random.default_rng(0)
    K, P = 3, 4
    ops = np.column_stack([rng.random((P, K * K)), rng.normal(-100, 10, P)])
    pi = np.array([0.5, 0.3, 0.2, 0.0])
    for r in range(P):
        fwd, bwd = combine_carries(ops, r, pi, K)
        v, ls = pi[:K].copy(), 0.0
        for q in range(r):
            v, ls = v @ ops[q, :9].reshape(3, 3), ls + ops[q, 9]
        assert np.allclose(np.log(fwd[:K]) + fwd[K], np.log(v) + ls, rtol=1e-12)
        u, lu = np.ones(K), 0.0
        for q in range(P - 1, r, -1):
            u, lu = ops[q, :9].reshape(3, 3) @ u, lu + ops[q, 9]
        assert np.allclose(np.log(bwd[:K]) + bwd[K], np.log(u) + lu, rtol=1e-12)


def _comm_worker(rank, world, initfile, outdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from epigrahmm_b200 import sharded
    dist.init_process_group("gloo", init_method="file://" + initfile, rank=rank, world_size=world)
    ok = True
    for kind in ("shm", "torch"):
        os.environ["EHMM_COMM"] = kind
        comm = sharded.make_comm(dist)
        assert isinstance(comm, sharded.ShmComm) == (kind == "shm")
        for it in range(300):                      # many epochs: both banks, ranks drifting apart
            n = 1 + (it * 7) % 50
            a = np.arange(n, dtype=np.float64) * (rank + 1) + it
            got = comm.all_gather(a)
            ok &= got.shape == (world, n) and all(np.array_equal(got[r], np.arange(n) * (r + 1.0) + it) for r in range(world))
            if it % 3 == rank:                     # uneven pacing
                _ = np.linalg.norm(np.ones(2000))
        cnt = comm.all_gather(np.array([rank, 10 * rank], dtype=np.int64))
        ok &= cnt.dtype == np.int64 and np.array_equal(cnt, [[r, 10 * r] for r in range(world)])
        ok &= np.array_equal(comm.all_reduce_sum(np.array([1.0, rank])), [world, sum(range(world))])
        ok &= np.array_equal(comm.broadcast(np.array([float(rank)]), world - 1), [world - 1.0])
        big = comm.all_gather(np.full(20000, float(rank)))      # larger than a slot: the torch path
        ok &= big.shape == (world, 20000) and all(np.all(big[r] == r) for r in range(world))
        if kind == "shm":
            comm.close()
    open(os.path.join(outdir, "ok%d" % rank), "w").write(str(bool(ok)))
    dist.destroy_process_group()


def test_shared_memory_exchange_three_ranks():
    """ehmm_comm_* (the host exchange of the sharded fit) against torch.distributed, world size 3"""
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_comm_worker, args=(3, os.path.join(td, "init"), td), nprocs=3, join=True)
        assert [open(os.path.join(td, "ok%d" % r)).read() for r in range(3)] == ["True"] * 3


class DiffBlockBackend(BlockBackend):
    """the differential additions of a bin block: mixture emission, innerExpStep, the inner M-step sums
    and the rejection control over the B + 2 columns (N group ids per bin)"""

    def __init__(self, O, counts, offsets, design, rng_state):
        super().__init__(O, counts, offsets, 3, rng_state)
        self.design, self.B = design, design.shape[0]

    def loglik_differential_hmm(self, delta, psi, minZero):
        self.h["logLikelihood"] = self.O.loglikDifferentialHMM(self.counts, self.offsets, self.design, delta, psi, minZero)

    def inner_expstep(self, delta, psi, minZero):
        self.h["mixtureProb"] = self.O.innerExpStep(self.counts, self.offsets, self.design, delta, psi, minZero)

    def inner_mstep_sums(self):
        p = np.exp(self.h["logProb1"][:, 1])
        return np.append((p[:, None] * self.h["mixtureProb"]).sum(0), p.sum())

    def _cols(self):
        P = np.exp(self.h["logProb1"])
        return [P[:, 0]] + [P[:, 1] * self.h["mixtureProb"][:, b] for b in range(self.B)] + [P[:, 2]]

    def rc_count(self, p, differential=False):
        assert differential
        return np.array([np.count_nonzero((x < p) & (x / p != 0)) for x in self._cols()], dtype=np.int64)

    def rc_apply_at(self, p, offs, total, differential=False):
        u = self.rng.runif(int(total))
        G = self.g["G"]
        self.buckets = np.zeros((self.B + 2, G + 1))
        for c, x in enumerate(self._cols()):
            w, n = self.O.reweight(x, p, uniforms=u[offs[c]:])
            for i in range(self.N):
                np.add.at(self.buckets[c], self.g["group"][:, i], w)


def _diff_worker(rank, world, initfile, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import oracle as O
    from epigrahmm_b200 import sharded, synth
    dist.init_process_group("gloo", init_method="file://" + initfile, rank=rank, world_size=world)
    M = 3001
    d = synth.make_differential(M, 2, 2, seed=9)
    cuts = [0, 900, 2100, M]                                   # three uneven blocks
    lo, hi = cuts[rank], cuts[rank + 1]
    counts, offsets = np.asfortranarray(d["counts"][lo:hi]), np.asfortranarray(d["offsets"][lo:hi])
    be = DiffBlockBackend(O, counts, offsets, d["design"], O.RNG.set_seed(31).state())
    s = sharded.ShardedSession("td", hi - lo, 3, 0, dist, local=be)
    g = sharded.global_groups(counts, offsets, s.comm, s.m0, s.Mtot, design=d["design"])
    be.set_groups(g)
    th = synth.THETA_DIFFERENTIAL
    delta = np.array([0.3, 0.7])
    be.loglik_differential_hmm(delta, th["psi"], MINZERO)
    s.expstep(th["pi"], th["gamma"])
    be.inner_expstep(delta, th["psi"], MINZERO)
    dl = s.inner_maxstepprob()
    total = s.rc_differential(0.05)
    np.savez(os.path.join(outdir, "d%d.npz" % rank), L1=be.h["logProb1"], eta=be.h["mixtureProb"], delta=dl,
             buckets=be.buckets, total=total, state=be.rng.state(), group=g["group"], G=g["G"])
    dist.destroy_process_group()


def test_three_rank_differential_protocol(oracle):
    """the differential side of the sharded protocol over three uneven bin blocks (gloo, CPU): delta from
    all-reduced inner M-step sums, one global draw order over the B + 2 columns, summed rejection tables"""
    import torch.multiprocessing as mp
    from epigrahmm_b200 import synth
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_diff_worker, args=(3, os.path.join(td, "init"), td), nprocs=3, join=True)
        r = [np.load(os.path.join(td, "d%d.npz" % i)) for i in range(3)]
        M = 3001
        d = synth.make_differential(M, 2, 2, seed=9)
        g = synth.make_groups(d["counts"], d["offsets"], design=d["design"])
        th = synth.THETA_DIFFERENTIAL
        delta = np.array([0.3, 0.7])
        logf = oracle.loglikDifferentialHMM(d["counts"], d["offsets"], d["design"], delta, th["psi"], MINZERO)
        h = oracle.expStep(th["pi"], th["gamma"], logf)
        h["mixtureProb"] = oracle.innerExpStep(d["counts"], d["offsets"], d["design"], delta, th["psi"], MINZERO)
        L1 = np.vstack([x["L1"] for x in r])
        assert np.max(np.abs(np.exp(L1) - np.exp(h["logProb1"]))) <= 1e-8
        assert np.array_equal(np.vstack([x["group"] for x in r]), g["group"]) and int(r[0]["G"]) == g["G"]
        want_delta = oracle.innerMaxStepProb(h)
        for x in r:
            assert np.max(np.abs(x["delta"] - want_delta)) <= 1e-8
        # rejection control on the merged posteriors of the ranks, one process, one stream
        hm = dict(h, logProb1=L1, mixtureProb=np.vstack([x["eta"] for x in r]))
        rng = oracle.RNG.set_seed(31)
        u = rng.runif(int(r[0]["total"]))
        rc, n = oracle.differentialRejectionControlled(hm, g["group"], 0.05, 4, uniforms=u, dense=True)
        assert n == int(r[0]["total"]) == int(r[2]["total"])
        for x in r:
            assert np.array_equal(x["state"], rng.state())
            assert np.allclose(x["buckets"], rc, rtol=1e-8, atol=0) and np.array_equal(x["buckets"] != 0, rc != 0)
