"""Bin-block sharding of one EM fit across GPUs (one process per GPU, torch.distributed).

The reference treats the whole genome as ONE Markov chain (R/base-core.R:181-183, src/expStep.cpp:
69-83), so GPUs hold contiguous bin blocks and exchange what keeps the result identical to the
single-process run:

* E-step: every block reduces to one scaled K x K operator (``ehmm_expstep_reduce``); the P
  operators are all-gathered (P*(K*K+1) doubles), each rank forms its forward carry
  ``pi * prod(ops[:r])`` and backward carry ``prod(ops[r+1:]) * 1`` on the host and applies them
  (``ehmm_expstep_apply``).  Sufficient statistics are all-reduced; the log-likelihood comes from
  the last block, pi's posterior row from the first.
* Rejection control: R draws uniforms column-major, then in ascending global bin order, so the
  per-column flagged counts are all-gathered and each rank consumes its own slice of the single
  Mersenne-Twister stream (``ehmm_rc_count`` / ``ehmm_rc_apply_at``); bucket sums are all-reduced
  in place in HBM, after which every rank holds the complete rejection tables and evaluates the
  NB-GLM objective redundantly (no further communication in the optimiser loop).

Everything exchanged is a few hundred bytes except the bucket vectors (columns x (G+1) doubles).
"""
from __future__ import annotations

import os

import numpy as np

from .session import Session


# ---- scaled operator algebra on the host (tiny: P operators of K x K) --------------------------

def _renorm(m, ls):
    mx = float(np.max(m))
    if not (mx > 0.0) or not np.isfinite(mx):
        return m, ls
    e = np.floor(np.log2(mx))
    return m * 2.0 ** (-e), ls + e * np.log(2.0)


def combine_carries(ops, rank, start, K):
    """ops: (P, K*K+1) block operators (row-major mantissa + log scale); start: (K+1,) forward vector
    entering block 0 (pi, 0).  Returns (fwd, bwd) for block `rank`, each (K+1,)."""
    ops = np.asarray(ops, dtype=np.float64)
    v, ls = np.array(start[:K], dtype=np.float64), float(start[K])
    for r in range(rank):
        v, ls = _renorm(v @ ops[r, :K * K].reshape(K, K), ls + ops[r, K * K])
    u, lu = np.ones(K), 0.0
    for r in range(ops.shape[0] - 1, rank, -1):
        u, lu = _renorm(ops[r, :K * K].reshape(K, K) @ u, lu + ops[r, K * K])
    return np.append(v, ls), np.append(u, lu)


def rc_offsets(all_counts, rank):
    """all_counts: (P, C) flagged counts per block and column.  R's order is column-major, then
    ascending bin, so block r's column c starts after all columns < c and blocks < r of column c."""
    all_counts = np.asarray(all_counts, dtype=np.int64)
    col_tot = all_counts.sum(axis=0)
    col_base = np.concatenate([[0], np.cumsum(col_tot)[:-1]])
    before = all_counts[:rank].sum(axis=0)
    return col_base + before, int(col_tot.sum())


class TorchComm:
    """the handful of collectives the sharded driver needs, over torch.distributed (nccl or gloo).
    Small host vectors go through preallocated pinned staging buffers and one
    all_gather_into_tensor, so an exchange costs one NCCL launch and one stream synchronisation."""

    def __init__(self, dist, device=None):
        import torch
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.nccl = dist.get_backend() == "nccl"
        self.device = device if self.nccl else "cpu"
        self._bufs = {}

    def _staging(self, n, dtype):
        key = (n, dtype)
        if key not in self._bufs:
            t = self.torch
            tdt = {np.dtype(np.float64): t.float64, np.dtype(np.int64): t.int64}[np.dtype(dtype)]
            hin = t.empty(n, dtype=tdt, pin_memory=self.nccl)
            hout = t.empty(self.world * n, dtype=tdt, pin_memory=self.nccl)
            din = t.empty(n, dtype=tdt, device=self.device) if self.nccl else hin
            dout = t.empty(self.world * n, dtype=tdt, device=self.device) if self.nccl else hout
            self._bufs[key] = (hin, hout, din, dout)
        return self._bufs[key]

    def all_gather(self, a):
        a = np.ascontiguousarray(a)
        if a.dtype not in (np.float64, np.int64) or a.size > 4096:
            return self._all_gather_general(a)
        hin, hout, din, dout = self._staging(a.size, a.dtype)
        hin.numpy()[:] = a.ravel()
        if self.nccl:
            din.copy_(hin, non_blocking=True)
        self.dist.all_gather_into_tensor(dout, din)
        if self.nccl:
            hout.copy_(dout, non_blocking=True)
            self.torch.cuda.current_stream().synchronize()
        return hout.numpy().reshape((self.world,) + a.shape).copy()

    def _all_gather_general(self, a):
        t = self.torch.as_tensor(a).to(self.device)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return np.stack([o.cpu().numpy() for o in out])

    def all_reduce_sum(self, a):
        return self.all_gather(a).sum(axis=0)

    def broadcast(self, a, src):
        return self.all_gather(a)[src]

    def all_reduce_device(self, ptr, n, stream=None, integer=False):
        """in-place sum of n doubles (or, integer=True, n 64-bit integers) at device pointer ptr (NCCL
        over NVLink).  With `stream` (the session's cudaStream_t) the collective is ordered on that
        stream -- after the kernels that produced the data and before the ones that read it -- and the
        host does not wait."""
        key = ("dev", ptr, n, integer)
        if key not in self._bufs:
            class _Buf:
                __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8" if integer else "<f8", "data": (ptr, False),
                                            "version": 3}
            self._bufs[key] = self.torch.as_tensor(_Buf(), device=self.device)
        if stream is None:
            self.dist.all_reduce(self._bufs[key])
            self.torch.cuda.current_stream().synchronize()
            return
        if ("stream", stream) not in self._bufs:
            self._bufs[("stream", stream)] = self.torch.cuda.ExternalStream(stream)
        with self.torch.cuda.stream(self._bufs[("stream", stream)]):
            self.dist.all_reduce(self._bufs[key])


class ShmComm(TorchComm):
    """TorchComm whose small host all-gathers go through the library's shared-memory exchange
    (ehmm_comm_*, one node) instead of a collective: the vectors are a few hundred bytes and both
    ends are host code, so a device collective would add two PCIe copies and a stream
    synchronisation each.  Large or oddly typed arrays and the in-HBM all-reduce stay on torch."""
    SLOT = 1 << 16

    def __init__(self, dist, device=None, timeout_s=120.0):
        super().__init__(dist, device)
        import ctypes as C
        import secrets
        from . import _lib
        self._lib, self._C = _lib, C
        # a name unique to this job, chosen by rank 0
        tok = [("/ehmm-%d-%s" % (os.getpid(), secrets.token_hex(6))) if self.rank == 0 else None]
        dist.broadcast_object_list(tok, src=0)
        h = C.c_void_p()
        _lib.check(_lib.lib.ehmm_comm_open(tok[0].encode(), self.rank, self.world, self.SLOT, float(timeout_s), C.byref(h)))
        self._h = h

    def all_gather(self, a):
        a = np.ascontiguousarray(a)
        if a.nbytes > self.SLOT or a.dtype.hasobject:
            return self._all_gather_general(a)
        out = np.empty((self.world,) + a.shape, dtype=a.dtype)
        self._lib.check(self._lib.lib.ehmm_comm_allgather(self._h, a.ctypes.data, a.nbytes, out.ctypes.data))
        return out

    def close(self):
        if self._h is not None:
            self._lib.lib.ehmm_comm_close(self._h)
            self._h = None


def make_comm(dist, device=None):
    """the exchange for ranks of one node (the bench / driver contract); set EHMM_COMM=torch to keep
    every exchange on torch.distributed (ranks on several nodes)"""
    import os
    if os.environ.get("EHMM_COMM", "shm") == "torch":
        return TorchComm(dist, device)
    return ShmComm(dist, device)


def _merge_dictionaries(comm, M, m0, Mtot, u_chip, u_offsets, u_controls, first_local, design):
    """One dt$Group dictionary from the ranks' local ones.  first_local: column-major cell index
    (sample * M + bin) of every local group's first appearance.  Returns (mine, G, rows, rep, samp):
    mine[l] = global id of local group l + 1, numbered by first appearance in the column-major GLOBAL
    table exactly as data.table's .GRP (R/base-core.R:97,187); rows[rep] = one record per global group."""
    first_local = np.asarray(first_local, dtype=np.int64)
    first_global = (first_local // M) * Mtot + m0 + (first_local % M)
    cols = [u_chip, u_offsets]
    if u_controls is not None:
        cols.append(u_controls)
    if design is not None:
        cols.append(first_local // M)  # the sample decides the Dsg.Mix columns
    rec = np.column_stack(cols + [first_global.astype(np.float64)])
    sizes = comm.all_gather(np.array([rec.shape[0]], dtype=np.int64)).ravel()
    pad = np.full((int(sizes.max()), rec.shape[1]), np.nan)
    pad[:rec.shape[0]] = rec
    allrec = comm.all_gather(pad)
    rows = np.concatenate([allrec[r, :sizes[r]] for r in range(comm.world)])
    key = rows[:, :-1].copy()
    if design is not None:  # samples with equal design columns share groups
        _, dcode = np.unique(np.asarray(design).T, axis=0, return_inverse=True)
        key[:, -1] = np.asarray(dcode).ravel()[key[:, -1].astype(np.int64)]
    uk, inv = np.unique(key, axis=0, return_inverse=True)
    inv = np.asarray(inv).ravel()
    first = np.full(uk.shape[0], np.inf)
    np.minimum.at(first, inv, rows[:, -1])
    order = np.argsort(first, kind="stable")
    rank_of = np.empty(order.size, dtype=np.int64)
    rank_of[order] = np.arange(order.size)
    gid = rank_of[inv] + 1                      # global id of every gathered record
    start = int(sizes[:comm.rank].sum())
    mine = gid[start:start + rec.shape[0]]      # local group -> global id
    rep = np.empty(order.size, dtype=np.int64)  # one gathered record per global group
    rep[gid - 1] = np.arange(gid.size)
    return mine, int(order.size), rows, rep


def _dictionary_columns(rows, rep, have_controls, design):
    out = dict(u_chip=rows[rep, 0].copy(), u_offsets=rows[rep, 1].copy())
    out["u_controls"] = rows[rep, 2].copy() if have_controls else None
    if design is not None:
        samp = rows[rep, -2].astype(np.int64)
        out["u_dsg"] = np.asfortranarray(np.asarray(design, dtype=np.uint8).T[samp])
    else:
        out["u_dsg"] = None
    return out


def global_groups(counts, offsets, comm, m0, Mtot, controls=None, design=None):
    """dt$Group over the whole sharded table, grouped on the host: every rank gets the same dictionary."""
    from .synth import make_groups
    loc = make_groups(counts, offsets, controls, design)
    M, N = np.asarray(counts).shape
    # first appearance of each local group as a column-major cell index
    g = loc["group"].ravel(order="F")
    first_local = np.full(loc["G"], M * N, dtype=np.int64)
    np.minimum.at(first_local, g - 1, np.arange(M * N, dtype=np.int64))
    mine, G, rows, rep = _merge_dictionaries(comm, M, m0, Mtot, loc["u_chip"], loc["u_offsets"],
                                             loc["u_controls"] if controls is not None else None, first_local, design)
    out = dict(group=np.asfortranarray(mine[loc["group"] - 1].astype(np.int32)), G=G)
    out.update(_dictionary_columns(rows, rep, controls is not None, design))
    return out


def global_groups_device(session, comm, m0, Mtot, design=None):
    """The same global dictionary with the per-cell work on the device: every rank groups its block
    (ehmm_build_groups), the ranks merge their G-row dictionaries on the host (a few 1e5 rows), and the
    local ids are rewritten in place (ehmm_groups_remap).  `session` already holds the block's data
    (and design).  Returns G."""
    session.build_groups()
    loc = session.groups_fetch(with_group=False)
    mine, G, rows, rep = _merge_dictionaries(comm, session.M, m0, Mtot, loc["u_chip"], loc["u_offsets"], loc["u_controls"],
                                             session.groups_first_cell(), design)
    cols = _dictionary_columns(rows, rep, loc["u_controls"] is not None, design)
    session.groups_remap(mine, cols["u_chip"], cols["u_offsets"], cols["u_controls"], cols["u_dsg"])
    return G


class ShardedSession:
    """Session-compatible backend for epigrahmm_b200.em over one bin block per rank."""

    def __init__(self, key, M, K, device, dist, local=None, comm=None):
        self.comm = comm or make_comm(dist, device="cuda:%d" % device)
        self.rank, self.world = self.comm.rank, self.comm.world
        sizes = self.comm.all_gather(np.array([M], dtype=np.int64)).ravel()
        self.Mtot, self.m0 = int(sizes.sum()), int(sizes[:self.rank].sum())
        self.local = local if local is not None else Session("%s.rank%d" % (key, self.rank), M, K, device)
        self.local.set_block(self.m0, self.Mtot)
        self.M, self.K = M, K
        self._hint = None        # rejection cut announced for the coming call (rc_hint)
        self._counts = None      # (p, differential, counts of all blocks): gathered along with another exchange

    # plain delegation -----------------------------------------------------------------------
    def __getattr__(self, name):
        return getattr(self.local, name)

    def close(self):
        self.local.close()

    def rc_hint(self, p):
        """the cut of the coming rejection-control call (R/base-EM.R:119-121 knows it before the E-step): the kernels
        that produce the posteriors count for it on the way, and the blocks' counts travel with the exchange that
        follows those kernels anyway instead of in one of their own"""
        self._hint = float(p)
        if hasattr(self.local, "rc_hint"):
            self.local.rc_hint(p)

    def _early_counts(self, differential):
        """this block's flagged counts for the announced cut, or None when no cut is announced or no groups are set.
        Both conditions hold on all ranks alike (the driver announces the cut and sets the groups everywhere), so every
        rank packs the same fields into the exchange; an error inside the count is raised, not swallowed"""
        if self._hint is None or not (self._hint > 0.0):
            return None
        groups = getattr(self.local, "G", None)
        if groups is None:                       # host-resident backend (CPU tests): a dict of group columns or None
            groups = 0 if getattr(self.local, "g", None) is None else 1
        if not groups:
            return None
        return np.asarray(self.local.rc_count(self._hint, differential), dtype=np.float64)

    # data ------------------------------------------------------------------------------------
    def _share_controls_first(self, have):
        """the reference's emission uses controls[1] of the WHOLE table for every cell (its ifelse()
        quirk, R/base-loglikelihood.R:113): every block takes block 0's value"""
        if have:
            v = self.local.controls_first() if self.rank == 0 else 0.0
            self.local.set_controls_first(float(self.comm.broadcast(np.array([v]), 0)[0]))

    def set_data(self, counts, offsets, controls=None):
        self.local.set_data(counts, offsets, controls)
        self._share_controls_first(controls is not None)

    def set_data_encoded(self, group, u_chip, u_offsets, u_controls=None, u_dsg=None, design=None):
        self.local.set_data_encoded(group, u_chip, u_offsets, u_controls, u_dsg, design)
        self._share_controls_first(u_controls is not None)

    # E-step ----------------------------------------------------------------------------------
    def expstep(self, pi, gamma, logf=None):
        K = self.K
        self._counts = None
        op = self.local.expstep_reduce(pi, gamma, logf)
        ops = self.comm.all_gather(op)
        if hasattr(self.local, "expstep_apply_ops"):       # carries composed inside the library, no wait
            self.local.expstep_apply_ops(ops, self.rank)
            stats = self.local.estep_stats()               # the E-step's one synchronisation
            ll_local = self.local.loglik()                 # forward mass after this block (the last block's is the chain's)
        else:                                              # host-resident backend (CPU tests)
            fwd, bwd = combine_carries(ops, self.rank, np.append(np.asarray(pi, dtype=np.float64), 0.0), K)
            tail = self.local.expstep_apply(fwd, bwd)      # forward vector after this block
            ll_local = float(tail[K] + np.log(np.sum(tail[:K])))
            stats = self.local.estep_stats()
        # one exchange: statistics (summed), logProb1[0,] (first block's), log-likelihood (last block's) and, for the
        # consensus model, the flagged counts of the rejection control that follows (every block or none has them:
        # the cut is announced on all ranks alike)
        cnt = self._early_counts(False) if K == 2 else None
        parts = [stats, self.local.fetch_logprob1_row0(), [ll_local]] + ([cnt] if cnt is not None else [])
        packed = self.comm.all_gather(np.concatenate(parts))
        ns = K * K + 1
        self.local.estep_set_stats(packed[:, :ns].sum(axis=0), packed[0, ns:ns + K], float(packed[-1, ns + K]))
        if cnt is not None:
            self._counts = (self._hint, False, np.rint(packed[:, ns + K + 1:]).astype(np.int64))

    # rejection control -----------------------------------------------------------------------
    def _rc(self, p, differential):
        if self._counts is not None and self._counts[0] == p and self._counts[1] == differential:
            allc = self._counts[2]          # gathered with the exchange behind the kernel that counted
        else:
            allc = self.comm.all_gather(self.local.rc_count(p, differential))
        self._counts = None
        offs, total = rc_offsets(allc, self.rank)
        self.local.rc_apply_at(p, offs, total, differential)
        if hasattr(self.local, "rc_buckets_host"):   # host-resident backend (CPU tests)
            self.local.rc_buckets_set(self.comm.all_reduce_sum(self.local.rc_buckets_host()))
        else:
            # exact integer limbs: the sum over the blocks does not depend on how the chain was cut
            ptr, n, integer = self.local.rc_buckets_device()
            self.comm.all_reduce_device(ptr, n, stream=self.local.stream, integer=integer)
            self.local.rc_finalize()
        return total

    def rc_consensus(self, p, uniforms=None):
        if uniforms is not None:
            raise ValueError("the sharded driver draws from the device generator")
        return self._rc(p, False)

    def rc_differential(self, p, uniforms=None):
        if uniforms is not None:
            raise ValueError("the sharded driver draws from the device generator")
        return self._rc(p, True)

    def inner_expstep(self, *a, **k):
        self._counts = None
        return self.local.inner_expstep(*a, **k)

    def inner_expstep_zinb(self, *a, **k):
        self._counts = None
        return self.local.inner_expstep_zinb(*a, **k)

    def inner_maxstepprob(self):
        sums = np.asarray(self.local.inner_mstep_sums(), dtype=np.float64)
        cnt = self._early_counts(True)     # innerExpStep counted for the announced cut: one exchange for both
        if cnt is None:
            s = self.comm.all_reduce_sum(sums)
        else:
            packed = self.comm.all_gather(np.concatenate([sums, cnt]))
            s = packed[:, :sums.size].sum(axis=0)
            self._counts = (self._hint, True, np.rint(packed[:, sums.size:]).astype(np.int64))
        return s[:-1] / s[-1]

    def viterbi(self, pi, gamma, fetch=True):
        """computeViterbiSequence over the blocks: the recursion is serial, so the ranks take turns
        (K doubles forwards, one state backwards).  Returns this rank's block of the path."""
        v = None
        for r in range(self.world):
            mine = self.local.viterbi_forward(pi, gamma, v) if r == self.rank else np.zeros(self.K)
            v = self.comm.broadcast(mine, r)
        state, out = -1, None
        for r in range(self.world - 1, -1, -1):
            if r == self.rank:
                out, prev = self.local.viterbi_backtrace(state, fetch)
            else:
                prev = 0
            state = int(self.comm.broadcast(np.array([float(prev)]), r)[0])
        return out
