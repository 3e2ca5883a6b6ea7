"""GPU: the bin-block (multi-GPU) form of the E-step and of the rejection control.

test_blocks_on_one_gpu drives several block Sessions of one chain from one process on one device,
performing the "collectives" by hand; it checks the device side of the protocol (operator
reduction, carries entering a block with m0 > 0, per-column stream offsets of the parallel
Mersenne-Twister) against a single Session holding the whole chain.  test_two_gpus_nccl runs the
real ShardedSession over NCCL when two devices are present."""
import os
import sys
import tempfile

import numpy as np
import pytest

from epigrahmm_b200 import synth
from epigrahmm_b200.sharded import combine_carries, rc_offsets

pytestmark = pytest.mark.gpu
MINZERO = 2.2250738585072014e-308
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("K,cuts", [(2, [0, 300007, 700001]), (2, [0, 1, 961, 5000]), (3, [0, 2500, 2501, 9000])])
def test_blocks_on_one_gpu(ehmm, oracle, K, cuts):
    M = cuts[-1]
    P = len(cuts) - 1
    th = synth.THETA_CONSENSUS if K == 2 else synth.THETA_DIFFERENTIAL
    if K == 2:
        d = synth.make_consensus(M, 2, seed=61)
        g = synth.make_groups(d["counts"], d["offsets"])
        design = None
    else:
        d = synth.make_differential(M, 2, 2, seed=62)
        design = d["design"]
        g = synth.make_groups(d["counts"], d["offsets"], design=design)
    delta = None if K == 2 else np.full(d["B"], 1.0 / d["B"])
    st0 = oracle.RNG.set_seed(5).state()

    def load(s, lo, hi):
        s.set_data(d["counts"][lo:hi], d["offsets"][lo:hi])
        if design is not None:
            s.set_design(design)
        s.set_groups(g["group"][lo:hi], g["u_chip"], g["u_offsets"], None, g["u_dsg"])
        s.rng_set_state(st0)
        if K == 2:
            s.loglik_consensus(th["psi"], MINZERO)
        else:
            s.loglik_differential_hmm(delta, th["psi"], MINZERO)
            s.inner_expstep(delta, th["psi"], MINZERO)

    whole = ehmm.Session("whole", M, K)
    load(whole, 0, M)
    whole.expstep(th["pi"], th["gamma"])
    ref = whole.datasets()
    blocks = []
    for r in range(P):
        s = ehmm.Session("blk%d" % r, cuts[r + 1] - cuts[r], K)
        s.set_block(cuts[r], M)
        load(s, cuts[r], cuts[r + 1])
        blocks.append(s)
    try:
        ops = np.stack([s.expstep_reduce(th["pi"], th["gamma"]) for s in blocks])
        tails = []
        for r, s in enumerate(blocks):
            fwd, bwd = combine_carries(ops, r, np.append(th["pi"], 0.0), K)
            tails.append(s.expstep_apply(fwd, bwd))
        for n in ("logFP", "logBP"):
            got = np.vstack([s.fetch(n) for s in blocks])
            assert np.max(np.abs(got - ref[n]) / np.maximum(1, np.abs(ref[n]))) <= 1e-11, n
        for n in ("logProb1", "logProb2"):
            got = np.vstack([s.fetch(n) for s in blocks])
            assert np.max(np.abs(np.exp(got) - np.exp(ref[n]))) <= 1e-10, n  # tile-local log scales are O(1e4): ulp ~ 2e-12
        stats = sum(s.estep_stats() for s in blocks)
        assert np.allclose(stats, whole.estep_stats(), rtol=1e-11)
        ll = tails[-1][K] + np.log(np.sum(tails[-1][:K]))
        assert abs(ll - whole.loglik()) <= 1e-12 * abs(ll)
        for s in blocks:
            s.estep_set_stats(stats, blocks[0].fetch_logprob1_row0(), ll)
        m0, m1 = whole.maxstepprob(), blocks[-1].maxstepprob()
        assert np.allclose(m0["pi"], m1["pi"], rtol=1e-11) and np.allclose(m0["gamma"], m1["gamma"], rtol=1e-11)
        assert abs(whole.bic(5, 2) - blocks[0].bic(5, 2)) <= 1e-12 * abs(whole.bic(5, 2))
        # Viterbi handed from block to block equals the whole-chain path bit for bit
        want = whole.viterbi(th["pi"], th["gamma"])
        v = None
        for b in blocks:
            v = b.viterbi_forward(th["pi"], th["gamma"], v)
        state, parts = -1, []
        for b in reversed(blocks):
            path, state = b.viterbi_backtrace(state)
            parts.append(path)
        assert np.array_equal(np.concatenate(parts[::-1]), want)
        # rejection control with the global draw order
        diff = K == 3
        for p in (0.05, 0.9):
            n_ref = whole.rc_differential(p) if diff else whole.rc_consensus(p)
            cnts = np.stack([s.rc_count(p, diff) for s in blocks])
            assert cnts.sum() == n_ref
            for r, s in enumerate(blocks):
                offs, total = rc_offsets(cnts, r)
                s.rc_apply_at(p, offs, total, diff)
            for c in range(cnts.shape[1]):
                merged = sum(s.rc_buckets(c) for s in blocks)
                want = whole.rc_buckets(c)
                assert np.array_equal(merged != 0, want != 0)
                assert np.allclose(merged, want, rtol=1e-11, atol=0)
            for s in blocks:
                assert np.array_equal(s.rng_get_state(), whole.rng_get_state())
        if diff:
            sums = sum(s.inner_mstep_sums() for s in blocks)
            assert np.allclose(sums[:-1] / sums[-1], whole.inner_maxstepprob(), rtol=1e-11)
    finally:
        whole.close()
        for s in blocks:
            s.close()


def test_rc_tables_do_not_depend_on_the_sharding(ehmm, oracle):
    """Given the same posteriors, the rejection tables of a chain cut into bin blocks are bit-identical to
    the unsharded ones: the blocks' exact integer limbs are added (here with torch, in the sharded
    driver with an NCCL all-reduce) before the single rounding of ehmm_rc_finalize."""
    import torch
    M, cuts = 50021, [0, 1, 4099, 30000, 50021]
    d = synth.make_consensus(M, 3, seed=71)
    g = synth.make_groups(d["counts"], d["offsets"])
    th = synth.THETA_CONSENSUS
    st0 = oracle.RNG.set_seed(17).state()

    def view(s):
        ptr, n, integer = s.rc_buckets_device()
        assert integer

        class _B:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}
        return torch.as_tensor(_B(), device="cuda")

    with ehmm.Session("inv-whole", M, 2) as whole:
        whole.set_data(d["counts"], d["offsets"])
        whole.set_groups(g["group"], g["u_chip"], g["u_offsets"])
        whole.loglik_consensus(th["psi"], MINZERO)
        whole.expstep(th["pi"], th["gamma"])
        lp1 = whole.fetch("logProb1")
        whole.rng_set_state(st0)
        n_ref = whole.rc_consensus(0.05)
        want = [whole.rc_buckets(c) for c in range(2)]
        blocks = []
        try:
            for r in range(len(cuts) - 1):
                lo, hi = cuts[r], cuts[r + 1]
                b = ehmm.Session("inv-blk%d" % r, hi - lo, 2)
                b.set_block(lo, M)
                b.set_data(d["counts"][lo:hi], d["offsets"][lo:hi])
                b.set_groups(g["group"][lo:hi], g["u_chip"], g["u_offsets"])
                b.store("logProb1", lp1[lo:hi])
                b.rng_set_state(st0)
                blocks.append(b)
            cnts = np.stack([b.rc_count(0.05) for b in blocks])
            assert cnts.sum() == n_ref
            for r, b in enumerate(blocks):
                offs, total = rc_offsets(cnts, r)
                b.rc_apply_at(0.05, offs, total)
            for b in blocks:
                b.sync()
            views = [view(b) for b in blocks]        # (folds the frequent groups' replicas in, on each session's stream)
            for b in blocks:
                b.sync()                             # torch adds on its own stream; the driver's NCCL call uses the session's
            acc = views[0]
            for v in views[1:]:
                acc += v                             # the "all-reduce"
            torch.cuda.synchronize()
            blocks[0].rc_finalize()
            for c in range(2):
                assert np.array_equal(blocks[0].rc_buckets(c), want[c]), c
        finally:
            for b in blocks:
                b.close()


def _nccl_worker(rank, world, initfile, outdir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import epigrahmm_b200 as E
    from epigrahmm_b200 import em, sharded
    from epigrahmm_b200.rng import RState
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="file://" + initfile, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    M, N, K = 400003, 3, 2
    d = synth.make_consensus(M, N, seed=71)
    lo, hi = (0, 150000) if rank == 0 else (150000, M)
    c, o = np.asfortranarray(d["counts"][lo:hi]), np.asfortranarray(d["offsets"][lo:hi])
    s = sharded.ShardedSession("nccl", hi - lo, K, rank, dist)
    g = sharded.global_groups(c, o, s.comm, s.m0, s.Mtot)
    s.set_data(c, o)
    s.set_groups(g["group"], g["u_chip"], g["u_offsets"])
    s.rng_set_state(RState(9).state)
    th = synth.THETA_CONSENSUS
    theta = dict(pi=th["pi"], gamma=th["gamma"], psi=[np.array([0.5, 2.0]), np.array([2.0, 2.0])])
    fit = em.emConsensus(theta, s, N, em.controlEM(maxIterEM=3))
    th_new = fit["theta_new"]
    vit = s.viterbi(th_new["pi"], th_new["gamma"])
    np.savez(os.path.join(outdir, "r%d.npz" % rank), par=em.unlist(fit["theta_new"]),
             bic=fit["controlHist"][-2]["bic"], q=fit["controlHist"][-2]["q"], L1=s.fetch("logProb1"),
             state=s.rng_get_state(), vit=vit)
    s.close()
    dist.destroy_process_group()


def test_two_gpus_nccl(ehmm):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from epigrahmm_b200 import em
    from epigrahmm_b200.rng import RState
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_nccl_worker, args=(2, os.path.join(td, "init"), td), nprocs=2, join=True)
        r = [np.load(os.path.join(td, "r%d.npz" % i)) for i in range(2)]
    M, N, K = 400003, 3, 2
    d = synth.make_consensus(M, N, seed=71)
    g = synth.make_groups(d["counts"], d["offsets"])
    th = synth.THETA_CONSENSUS
    theta = dict(pi=th["pi"], gamma=th["gamma"], psi=[np.array([0.5, 2.0]), np.array([2.0, 2.0])])
    with ehmm.Session("single", M, K) as s:
        s.set_data(d["counts"], d["offsets"])
        s.set_groups(g["group"], g["u_chip"], g["u_offsets"])
        s.rng_set_state(RState(9).state)
        fit = em.emConsensus(theta, s, N, em.controlEM(maxIterEM=3))
        L1 = s.fetch("logProb1")
        st = s.rng_get_state()
        vit = s.viterbi(fit["theta_new"]["pi"], fit["theta_new"]["gamma"])
    par = em.unlist(fit["theta_new"])
    for i in range(2):
        assert np.max(np.abs(r[i]["par"] - par) / np.abs(par)) <= 1e-9
        assert abs(r[i]["bic"] - fit["controlHist"][-2]["bic"]) <= 1e-9 * abs(r[i]["bic"])
        assert np.array_equal(r[i]["state"], st)
    got = np.vstack([r[0]["L1"], r[1]["L1"]])
    assert np.max(np.abs(np.exp(got) - np.exp(L1))) <= 1e-8
    # the sharded Viterbi path (logFP differs from the single-GPU run in the last bits only)
    assert np.mean(np.concatenate([r[0]["vit"], r[1]["vit"]]) != vit) <= 1e-4
