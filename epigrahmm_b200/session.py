"""Device-resident EM session: a thin object wrapper over the C ABI.

A session plays the role of the reference's HDF5 file (``metadata(object)$output``):
it is created per fit, keyed by the same path string, and holds the datasets
logLikelihood / logFP / logBP / logProb1 / logProb2 / mixtureProb / viterbi in HBM.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import bptr, check, dptr, iptr, lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _cm(a, dtype=np.float64):
    """column-major (R / Armadillo) copy"""
    return np.asfortranarray(a, dtype=dtype)


class Session:
    def __init__(self, key: str, M: int, K: int, device: int = 0):
        self._h = C.c_void_p()
        check(lib.ehmm_session_open(str(key).encode(), int(M), int(K), int(device), C.byref(self._h)))
        self.key, self.M, self.K, self.device = str(key), int(M), int(K), int(device)
        self.N = 0
        self.B = 0
        self.G = 0
        self._glm_cache = {}

    # -- lifetime --------------------------------------------------------------------------
    def close(self):
        if self._h:
            check(lib.ehmm_session_close(self._h))
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        p = C.c_void_p()
        check(lib.ehmm_session_stream(self._h, C.byref(p)))
        return p.value or 0

    def sync(self):
        check(lib.ehmm_session_sync(self._h))

    def set_block(self, m0: int, Mtot: int):
        check(lib.ehmm_session_set_block(self._h, int(m0), int(Mtot)))

    # -- data ------------------------------------------------------------------------------
    def set_data(self, counts, offsets, controls=None):
        counts = _cm(counts, np.int32).reshape(self.M, -1, order="F")
        N = counts.shape[1]
        offsets = _cm(offsets).reshape(self.M, N, order="F")
        ctrl = None if controls is None else _cm(controls).reshape(self.M, N, order="F")
        check(lib.ehmm_set_data(self._h, N, iptr(counts), dptr(offsets), dptr(ctrl)))
        self.N = N

    def set_data_device(self, N, counts_ptr, offsets_ptr, controls_ptr=None):
        check(lib.ehmm_set_data_device(self._h, int(N), C.c_void_p(counts_ptr), C.c_void_p(offsets_ptr),
                                       C.c_void_p(controls_ptr) if controls_ptr else None))
        self.N = int(N)

    def controls_first(self) -> float:
        """controls[1] of the long table: the one control value the reference's emission uses
        (R/base-loglikelihood.R:113; see include/ehmm.h)"""
        v = C.c_double()
        check(lib.ehmm_controls_first(self._h, C.byref(v)))
        return v.value

    def set_controls_first(self, value: float):
        check(lib.ehmm_set_controls_first(self._h, float(value)))

    def build_groups(self) -> int:
        """dt$Group + dtUnique on the device (after set_data / set_design); returns G"""
        G = C.c_int64()
        check(lib.ehmm_build_groups(self._h, C.byref(G)))
        self.G = G.value
        return self.G

    def groups_fetch(self, with_group=True):
        """dict(group, u_chip, u_offsets, u_controls, u_dsg) as synth.make_groups returns them"""
        G = self.G
        grp = np.empty((self.M, self.N), dtype=np.int32, order="F") if with_group else None
        uc, uo = np.empty(G), np.empty(G)
        out = dict(G=G, group=grp, u_chip=uc, u_offsets=uo, u_controls=None, u_dsg=None)
        check(lib.ehmm_groups_fetch(self._h, iptr(grp), dptr(uc), dptr(uo), None, None))
        try:
            uctrl = np.empty(G)
            check(lib.ehmm_groups_fetch(self._h, None, None, None, dptr(uctrl), None))
            out["u_controls"] = uctrl
        except _lib.EhmmError:
            pass
        if self.B:
            try:
                ud = np.empty((G, self.B), dtype=np.uint8, order="F")
                check(lib.ehmm_groups_fetch(self._h, None, None, None, None, bptr(ud)))
                out["u_dsg"] = ud
            except _lib.EhmmError:
                pass
        return out

    def groups_first_cell(self):
        """column-major cell index (sample * M + bin) of every group's first appearance (build_groups)"""
        out = np.empty(self.G, dtype=np.int64)
        check(lib.ehmm_groups_first_cell(self._h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    def groups_remap(self, new_id, u_chip, u_offsets, u_controls=None, u_dsg=None):
        """rewrite dt$Group in place: old id g -> new_id[g - 1]; dtUnique is replaced by the given columns"""
        nid = np.ascontiguousarray(new_id, dtype=np.int32)
        assert nid.size == self.G
        uc, uo = _f64(u_chip), _f64(u_offsets)
        uctrl = None if u_controls is None else _f64(u_controls)
        udsg = None if u_dsg is None else np.asfortranarray(u_dsg, dtype=np.uint8)
        check(lib.ehmm_groups_remap(self._h, iptr(nid), uc.size, dptr(uc), dptr(uo), dptr(uctrl), bptr(udsg)))
        self.G = uc.size

    def set_data_encoded(self, group, u_chip, u_offsets, u_controls=None, u_dsg=None, design=None):
        """upload the table as dt$Group + dtUnique only (the counts/offsets matrices stay on the host)"""
        group = np.asarray(group)
        narrow = group.dtype == np.uint16      # 16-bit ids: half the upload, widened on the device
        g = _cm(group, np.uint16 if narrow else np.int32)
        g = g.reshape(self.M, -1, order="F")
        N = g.shape[1]
        if design is not None:
            d = np.ascontiguousarray(design, dtype=np.uint8)
            check(lib.ehmm_set_design_n(self._h, N, d.shape[0], bptr(d)))
            self.B = d.shape[0]
        uc, uo = _f64(u_chip), _f64(u_offsets)
        uctrl = None if u_controls is None else _f64(u_controls)
        udsg = None if u_dsg is None else np.asfortranarray(u_dsg, dtype=np.uint8)
        if narrow:
            check(lib.ehmm_set_data_encoded16(self._h, N, g.ctypes.data_as(C.POINTER(C.c_uint16)), uc.size, dptr(uc),
                                              dptr(uo), dptr(uctrl), bptr(udsg)))
        else:
            check(lib.ehmm_set_data_encoded(self._h, N, iptr(g), uc.size, dptr(uc), dptr(uo), dptr(uctrl), bptr(udsg)))
        self.N, self.G = N, uc.size

    def prefetch_groups(self, group):
        """start the upload of the NEXT table's Group column (pinned memory) on the copy stream; the next
        set_data_encoded(group, ...) with the same array takes the staged copy"""
        group = np.asarray(group)
        g = _cm(group, np.uint16 if group.dtype == np.uint16 else np.int32)
        check(lib.ehmm_prefetch_groups(self._h, g.size // self.M, g.ctypes.data, g.itemsize))

    def set_design(self, design):
        d = np.ascontiguousarray(design, dtype=np.uint8)
        assert d.ndim == 2 and d.shape[1] == self.N, "design must be B x N"
        check(lib.ehmm_set_design(self._h, d.shape[0], bptr(d)))
        self.B = d.shape[0]

    def set_groups(self, group, u_chip, u_offsets, u_controls=None, u_dsg=None):
        g = _cm(group, np.int32).reshape(self.M, self.N, order="F")
        uc, uo = _f64(u_chip), _f64(u_offsets)
        G = uc.size
        uctrl = None if u_controls is None else _f64(u_controls)
        udsg = None if u_dsg is None else np.asfortranarray(u_dsg, dtype=np.uint8)
        check(lib.ehmm_set_groups(self._h, iptr(g), G, dptr(uc), dptr(uo), dptr(uctrl), bptr(udsg)))
        self.G = G

    # -- emission --------------------------------------------------------------------------
    def loglik_consensus(self, psi, minZero):
        t1, t2 = _f64(psi[0]), _f64(psi[1])
        check(lib.ehmm_loglik_consensus(self._h, dptr(t1), dptr(t2), t1.size, float(minZero)))

    def loglik_consensus_zinb(self, psi, minZero):
        """dist = 'zinb': psi[0] = (lambda0[, lambda_ctrl], beta0[, beta_ctrl], size), psi[1] as for 'nb'"""
        t1, t2 = _f64(psi[0]), _f64(psi[1])
        check(lib.ehmm_loglik_consensus_zinb(self._h, dptr(t1), dptr(t2), t1.size, float(minZero)))

    def loglik_init(self, psi):
        t1, t2 = _f64(psi[0]), _f64(psi[1])
        check(lib.ehmm_loglik_init(self._h, dptr(t1), dptr(t2)))

    def loglik_differential_hmm(self, delta, psi, minZero):
        d, p = _f64(delta), _f64(np.concatenate([np.ravel(x) for x in psi]) if isinstance(psi, (list, tuple)) else psi)
        check(lib.ehmm_loglik_differential_hmm(self._h, dptr(d), dptr(p), float(minZero)))

    def loglik_differential_hmm_zinb(self, delta, psi, minZero):
        """dist = 'zinb': psi = (lambda, b_bg, l_bg, db, dl)"""
        d, p = _f64(delta), _f64(np.concatenate([np.ravel(x) for x in psi]) if isinstance(psi, (list, tuple)) else psi)
        check(lib.ehmm_loglik_differential_hmm_zinb(self._h, dptr(d), dptr(p), float(minZero)))

    def inner_expstep_zinb(self, delta, psi, minZero):
        d, p = _f64(delta), _f64(np.concatenate([np.ravel(x) for x in psi]) if isinstance(psi, (list, tuple)) else psi)
        check(lib.ehmm_inner_expstep_zinb(self._h, dptr(d), dptr(p), float(minZero)))

    def inner_expstep(self, delta, psi, minZero):
        d, p = _f64(delta), _f64(np.concatenate([np.ravel(x) for x in psi]) if isinstance(psi, (list, tuple)) else psi)
        check(lib.ehmm_inner_expstep(self._h, dptr(d), dptr(p), float(minZero)))

    # -- E-step / M-step -------------------------------------------------------------------
    def expstep(self, pi, gamma, logf=None):
        pi, gamma = _f64(pi), _cm(gamma)
        lf = None if logf is None else _cm(logf)
        if lf is not None:
            assert lf.shape == (self.M, self.K), "logf must be M x K"
        check(lib.ehmm_expstep(self._h, dptr(pi), dptr(gamma), dptr(lf)))

    def expstep_device(self, pi, gamma, logf_ptr: int):
        pi, gamma = _f64(pi), _cm(gamma)
        check(lib.ehmm_expstep_device(self._h, dptr(pi), dptr(gamma), C.c_void_p(logf_ptr)))

    def expstep_reduce(self, pi, gamma, logf=None):
        pi, gamma = _f64(pi), _cm(gamma)
        lf = None if logf is None else _cm(logf)
        op = np.empty(self.K * self.K + 1)
        check(lib.ehmm_expstep_reduce(self._h, dptr(pi), dptr(gamma), dptr(lf), dptr(op)))
        return op

    def expstep_apply(self, fwd, bwd):
        fwd, bwd = _f64(fwd), _f64(bwd)
        out = np.empty(self.K + 1)
        check(lib.ehmm_expstep_apply(self._h, dptr(fwd), dptr(bwd), dptr(out)))
        return out

    def expstep_apply_ops(self, ops, rank):
        ops = _f64(ops)
        check(lib.ehmm_expstep_apply_ops(self._h, dptr(ops), int(ops.shape[0]), int(rank)))

    def estep_stats(self):
        st = np.empty(self.K * self.K + 1)
        check(lib.ehmm_estep_stats(self._h, dptr(st)))
        return st

    def fetch_logprob1_row0(self):
        r = np.empty(self.K)
        check(lib.ehmm_estep_row0(self._h, dptr(r)))
        return r

    def estep_set_stats(self, stats, lp1_row0, loglik):
        check(lib.ehmm_estep_set_stats(self._h, dptr(_f64(stats)), dptr(_f64(lp1_row0)), float(loglik)))

    def loglik(self) -> float:
        v = C.c_double()
        check(lib.ehmm_loglik_value(self._h, C.byref(v)))
        return v.value

    def maxstepprob(self):
        pi = np.empty(self.K)
        gamma = np.empty((self.K, self.K), order="F")
        check(lib.ehmm_maxstepprob(self._h, dptr(pi), dptr(gamma)))
        return {"pi": pi, "gamma": gamma}

    def qfunction(self, pi, gamma) -> float:
        v = C.c_double()
        check(lib.ehmm_qfunction(self._h, dptr(_f64(pi)), dptr(_cm(gamma)), C.byref(v)))
        return v.value

    def bic(self, numPar, numSamples) -> float:
        v = C.c_double()
        check(lib.ehmm_bic(self._h, int(numPar), int(numSamples), C.byref(v)))
        return v.value

    def inner_maxstepprob(self):
        d = np.empty(self.B)
        check(lib.ehmm_inner_maxstepprob(self._h, dptr(d)))
        return d

    def marginal_probability(self):
        out = np.empty((self.M, self.K), order="F")
        check(lib.ehmm_marginal_probability(self._h, dptr(out)))
        return out

    def viterbi(self, pi, gamma, fetch=True):
        out = np.empty(self.M) if fetch else None
        check(lib.ehmm_viterbi(self._h, dptr(_f64(pi)), dptr(_cm(gamma)), dptr(out)))
        return out

    # -- callPeaks --------------------------------------------------------------------------
    def call_peaks(self, method="viterbi", exact=False):
        """R/callPeaks.R:61-62: boolean peak index per window; method 'viterbi' or an FDR level"""
        calls = np.empty(self.M, dtype=np.uint8)
        n = C.c_int64()
        if method == "viterbi":
            check(lib.ehmm_call_peaks_viterbi(self._h, calls.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(n)))
        else:
            cut = C.c_double()
            check(lib.ehmm_call_peaks_exact(self._h, int(bool(exact))))
            check(lib.ehmm_call_peaks(self._h, float(method), calls.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(n), C.byref(cut)))
            self.peaks_cutoff = cut.value
        assert int(calls.sum()) == n.value
        return calls.astype(bool)

    def call_peaks_path(self):
        p = C.c_int()
        check(lib.ehmm_call_peaks_path(self._h, C.byref(p)))
        return ("fp64 prefix sum, margins certified", "emulated long double (forced)", "emulated long double (margin too small)")[p.value]

    def peak_runs(self, chrom_starts=()):
        """(starts, ends), inclusive window indices of the merged peaks of the last call_peaks"""
        b = np.ascontiguousarray(chrom_starts, dtype=np.int64)
        bp = b.ctypes.data_as(C.POINTER(C.c_int64)) if b.size else None
        n = C.c_int64()
        check(lib.ehmm_peak_runs(self._h, bp, b.size, 0, None, None, C.byref(n)))
        st, en = np.empty(n.value, dtype=np.int64), np.empty(n.value, dtype=np.int64)
        if n.value:
            check(lib.ehmm_peak_runs(self._h, bp, b.size, n.value, st.ctypes.data_as(C.POINTER(C.c_int64)),
                                     en.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(n)))
        return st, en

    def flush(self):
        """every dataset of the fit, one device -> host pass (the end of the fit, R/base-core.R:212-226)"""
        out, args = {}, []
        for nm, cols, bit in (("logLikelihood", self.K, True), ("logFP", self.K, True), ("logBP", self.K, True), ("logProb1", self.K, True),
                              ("logProb2", self.K * self.K, True), ("mixtureProb", self.B, self.B > 0), ("viterbi", 1, True)):
            try:
                self.dims(nm)
                out[nm] = np.empty((self.M, cols), order="F")
                args.append(dptr(out[nm]))
            except _lib.EhmmError:
                args.append(None)
        check(lib.ehmm_flush(self._h, *args))
        return out

    def expstep_mode(self, lean):
        """1: E-steps inside an EM loop write the posteriors and the statistics only; logFP / logBP / logProb1 /
        logProb2 of the last E-step are produced when somebody asks for them.  0: the reference's contract."""
        check(lib.ehmm_expstep_mode(self._h, int(bool(lean))))

    def materialize(self):
        check(lib.ehmm_materialize(self._h))

    def viterbi_mode(self, mode):
        """0: parallel scan (default), 1: the sequential kernel; the states are bit-identical"""
        check(lib.ehmm_viterbi_mode(self._h, int(mode)))

    def viterbi_info(self):
        info = (C.c_int64 * 4)()
        check(lib.ehmm_viterbi_info(self._h, info))
        return {"path": ("scan", "sequential", "scan abandoned")[info[0]], "break_tiles": int(info[1]),
                "halfway_tiles": int(info[2]), "tiles": int(info[3])}

    def viterbi_forward(self, pi, gamma, v_in=None):
        v_out = np.empty(self.K)
        check(lib.ehmm_viterbi_forward(self._h, dptr(_f64(pi)), dptr(_cm(gamma)), dptr(None if v_in is None else _f64(v_in)),
                                       dptr(v_out)))
        return v_out

    def viterbi_backtrace(self, state_last=-1, fetch=True):
        out = np.empty(self.M) if fetch else None
        prev = C.c_int()
        check(lib.ehmm_viterbi_backtrace(self._h, int(state_last), C.byref(prev), dptr(out)))
        return out, prev.value

    # -- rejection control -----------------------------------------------------------------
    def rng_set_state(self, state625):
        st = np.ascontiguousarray(state625, dtype=np.uint32)
        assert st.size == 625
        check(lib.ehmm_rng_set_state(self._h, st.ctypes.data_as(C.POINTER(C.c_uint32))))

    def rng_get_state(self):
        st = np.empty(625, dtype=np.uint32)
        check(lib.ehmm_rng_get_state(self._h, st.ctypes.data_as(C.POINTER(C.c_uint32))))
        return st

    def rc_hint(self, p):
        """announce the next rejection-control call's cut (lets the E-step / innerExpStep count for it)"""
        check(lib.ehmm_rc_hint(self._h, float(p)))

    def rc_consensus(self, p, uniforms=None) -> int:
        u = None if uniforms is None else _f64(uniforms)
        n = C.c_int64()
        check(lib.ehmm_rc_consensus(self._h, float(p), dptr(u), 0 if u is None else u.size, C.byref(n)))
        self.rc_columns = self.K
        return n.value

    def rc_differential(self, p, uniforms=None) -> int:
        u = None if uniforms is None else _f64(uniforms)
        n = C.c_int64()
        check(lib.ehmm_rc_differential(self._h, float(p), self.N, dptr(u), 0 if u is None else u.size, C.byref(n)))
        self.rc_columns = self.B + 2
        return n.value

    def rc_count(self, p, differential=False):
        cols = self.B + 2 if differential else self.K
        cnt = np.zeros(cols, dtype=np.int64)
        check(lib.ehmm_rc_count(self._h, 1 if differential else 0, float(p), cnt.ctypes.data_as(C.POINTER(C.c_int64))))
        return cnt

    def rc_apply_at(self, p, col_offsets, total, differential=False):
        off = np.ascontiguousarray(col_offsets, dtype=np.int64)
        check(lib.ehmm_rc_apply_at(self._h, 1 if differential else 0, float(p), off.ctypes.data_as(C.POINTER(C.c_int64)),
                                   int(total)))
        self.rc_columns = off.size

    def rc_buckets_device(self):
        """(device pointer, element count, is_integer) of the rejection tables' accumulators: exact
        uint64 limbs (is_integer) or, for cuts below 2^-9, fp64 sums"""
        p, n, it = C.c_void_p(), C.c_int64(), C.c_int()
        check(lib.ehmm_rc_buckets_device(self._h, C.byref(p), C.byref(n), C.byref(it)))
        return p.value, n.value, bool(it.value)

    def rc_finalize(self):
        check(lib.ehmm_rc_finalize(self._h))

    def inner_mstep_sums(self):
        out = np.empty(self.B + 1)
        check(lib.ehmm_inner_mstep_sums(self._h, dptr(out)))
        return out

    def rc_table(self, column):
        """(rows x 2) matrix of (weight sum, group id): one element of the reference's field<mat>."""
        n = C.c_int64()
        check(lib.ehmm_rc_table_rows(self._h, int(column), C.byref(n)))
        sums, ids = np.empty(n.value), np.empty(n.value)
        check(lib.ehmm_rc_table_fetch(self._h, int(column), dptr(sums), dptr(ids)))
        return np.column_stack([sums, ids])

    def rc_buckets(self, column):
        b = np.empty(self.G + 1)
        check(lib.ehmm_rc_buckets_fetch(self._h, int(column), dptr(b)))
        return b

    # -- GLM -------------------------------------------------------------------------------
    def glm_nb(self, column, par, grad=True):
        par = _f64(par)
        v = C.c_double()
        g = np.empty(par.size) if grad else None
        check(lib.ehmm_glm_nb(self._h, int(column), dptr(par), par.size - 1, C.byref(v), dptr(g)))
        return v.value, g

    def glm_zinb(self, column, par, minZero, grad=True):
        """glmZINB/derivZINB: par = (beta[P], phi, lambda[P])"""
        par = _f64(par)
        v = C.c_double()
        g = np.empty(par.size) if grad else None
        check(lib.ehmm_glm_zinb(self._h, int(column), dptr(par), (par.size - 1) // 2, float(minZero), C.byref(v), dptr(g)))
        return v.value, g

    def glm_nb_batch(self, columns, pars):
        """value and gradient for several (column, par) pairs in one launch"""
        pars = np.ascontiguousarray(pars, dtype=np.float64)
        n, P1 = pars.shape
        key = (tuple(columns), P1)
        c = self._glm_cache.get(key)
        if c is None:   # the optimiser calls this in a tight loop: keep the small arrays and their pointers
            cols = np.ascontiguousarray(columns, dtype=np.int32)
            v, g = np.empty(n), np.empty((n, P1))
            c = self._glm_cache[key] = (cols, v, g, cols.ctypes.data_as(C.POINTER(C.c_int)), dptr(v), dptr(g))
        cols, v, g, pc, pv, pg = c
        rc = lib.ehmm_glm_nb_batch(self._h, n, pc, dptr(pars), P1 - 1, pv, pg)
        if rc:
            check(rc)
        return v.copy(), g.copy()

    def glm_mix_nb(self, par, minZero, grad=True):
        par = _f64(par)
        v = C.c_double()
        g = np.empty(4) if grad else None
        check(lib.ehmm_glm_mix_nb(self._h, dptr(par), float(minZero), C.byref(v), dptr(g)))
        return v.value, g

    def glm_mix_zinb(self, par, minZero, grad=True):
        """glmMixZINB/derivMixZINB: par = (lambda, b_bg, l_bg, b_enr, l_enr)"""
        par = _f64(par)
        v = C.c_double()
        g = np.empty(5) if grad else None
        check(lib.ehmm_glm_mix_zinb(self._h, dptr(par), float(minZero), C.byref(v), dptr(g)))
        return v.value, g

    # -- launch accounting -----------------------------------------------------------------
    def profile(self, enable=True):
        check(lib.ehmm_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        """{kernel name: (launches, summed CUDA-event ms)} since the last profile() call"""
        n = lib.ehmm_profile_kernels()
        cnt = np.zeros(n, dtype=np.int64)
        ms = np.zeros(n)
        check(lib.ehmm_profile_read(self._h, cnt.ctypes.data_as(C.POINTER(C.c_int64)), dptr(ms)))
        return {lib.ehmm_profile_name(i).decode(): (int(cnt[i]), float(ms[i])) for i in range(n)}

    # -- datasets --------------------------------------------------------------------------
    def dims(self, name):
        r, c = C.c_int64(), C.c_int64()
        check(lib.ehmm_dataset_dims(self._h, name.encode(), C.byref(r), C.byref(c)))
        return r.value, c.value

    def fetch(self, name):
        r, c = self.dims(name)
        out = np.empty((r, c), order="F")
        check(lib.ehmm_fetch(self._h, name.encode(), dptr(out), r * c))
        return out

    def store(self, name, a):
        a = _cm(a)
        if a.ndim == 1:
            a = a.reshape(-1, 1, order="F")
        check(lib.ehmm_store(self._h, name.encode(), dptr(a), a.shape[0], a.shape[1]))
        if name == "mixtureProb":
            self.B = a.shape[1]

    def device_ptr(self, name) -> int:
        p = C.c_void_p()
        check(lib.ehmm_device_ptr(self._h, name.encode(), C.byref(p)))
        return p.value

    def datasets(self):
        """all datasets the reference's HDF5 file would hold at this point"""
        out = {}
        for n in ("logLikelihood", "logFP", "logBP", "logProb1", "logProb2", "mixtureProb", "viterbi"):
            try:
                out[n] = self.fetch(n)
            except _lib.EhmmError:
                pass
        return out


def reweight(x, p, uniforms, device=0):
    """src/reweight.cpp:16-22 on the device; returns (weights, uniforms consumed)."""
    x = _f64(x).copy()
    u = _f64(uniforms)
    n = C.c_int64()
    check(lib.ehmm_reweight(int(device), dptr(x), x.size, float(p), dptr(u), u.size, C.byref(n)))
    return x, n.value


def aggregate(x, f, device=0):
    """src/aggregate.cpp:20-45 on the device; (rows x 2) of (sum, id) with non-zero sum."""
    x = _f64(x)
    f = np.ascontiguousarray(f, dtype=np.int32)
    G = int(f[: x.size].max()) if x.size else 0
    b = np.zeros(G + 1)
    check(lib.ehmm_aggregate(int(device), dptr(x), x.size, iptr(f), G, dptr(b)))
    ids = np.nonzero(b)[0]
    return np.column_stack([b[ids], ids.astype(float)])
