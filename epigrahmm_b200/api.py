"""The reference's Rcpp exports, by name and argument order (R/RcppExports.R:4-116).

Each function takes the ``hdf5`` path string the reference passes around as its only cross-call
handle; here it keys a device-resident :class:`~epigrahmm_b200.session.Session`.  ``expStep``
creates (or replaces) the session, exactly as the reference's ``expStep`` (re)writes the datasets
with ``hdf5_opts::replace`` (src/expStep.cpp:116-122); every other function fails with
``EhmmError`` when no session exists, where the reference would fail to open the file.

Rejection control consumes the module-level R-compatible stream (``epigrahmm_b200.rng``,
``set_seed`` = R's ``set.seed``), mirroring ``Rcpp::RNGScope``.
"""
from __future__ import annotations

import numpy as np

from . import rng as _rng
from ._lib import EhmmError
from .session import Session, aggregate as _aggregate, reweight as _reweight

_sessions: dict[str, Session] = {}


def _key(hdf5) -> str:
    # Rcpp::StringVector: only element 0 is used (src/expStep.cpp:50-51)
    if isinstance(hdf5, (list, tuple, np.ndarray)):
        hdf5 = hdf5[0]
    return str(hdf5)


def session(hdf5) -> Session:
    k = _key(hdf5)
    if k not in _sessions:
        raise EhmmError(3, "no session for '%s' (the reference would fail to read the HDF5 file)" % k)
    return _sessions[k]


def open_session(hdf5, M: int, K: int, device: int = 0) -> Session:
    k = _key(hdf5)
    old = _sessions.pop(k, None)
    if old is not None:
        old.close()
    s = Session(k, M, K, device)
    _sessions[k] = s
    return s


def close_session(hdf5) -> None:
    """the counterpart of deleting the HDF5 file (R/base-core.R:61)"""
    s = _sessions.pop(_key(hdf5), None)
    if s is not None:
        s.close()


def set_seed(seed: int) -> None:
    _rng.set_seed(seed)


# ---- src/*.cpp exports -------------------------------------------------------------------------

def expStep(pi, gamma, logf, hdf5) -> None:
    """E-step of HMM (src/expStep.cpp:45-123)."""
    logf = np.asarray(logf, dtype=np.float64)
    if logf.ndim != 2:
        raise EhmmError(1, "logf must be a matrix")
    M, K = logf.shape
    pi = np.asarray(pi, dtype=np.float64).ravel()
    if pi.size != K or np.shape(gamma) != (K, K):
        raise EhmmError(1, "pi, gamma and logf disagree on the number of states")
    k = _key(hdf5)
    s = _sessions.get(k)
    if s is None or s.M != M or s.K != K:
        s = open_session(k, M, K)
    s.expstep(pi, gamma, logf)


def maxStepProb(hdf5) -> dict:
    """M-step w.r.t. initial and transition probabilities (src/maxStepProb.cpp:37-81)."""
    return session(hdf5).maxstepprob()


def computeViterbiSequence(hdf5, pi, gamma) -> np.ndarray:
    """src/computeViterbiSequence.cpp:12-54; also stores the 'viterbi' dataset."""
    return session(hdf5).viterbi(pi, gamma)


def computeQFunction(hdf5, pi, gamma) -> float:
    return session(hdf5).qfunction(pi, gamma)


def computeBIC(hdf5, numPar, numSamples) -> float:
    return session(hdf5).bic(numPar, numSamples)


def innerMaxStepProb(hdf5) -> np.ndarray:
    return session(hdf5).inner_maxstepprob()


def getMarginalProbability(hdf5) -> np.ndarray:
    return session(hdf5).marginal_probability()


def _ensure_groups(s: Session, f) -> None:
    f = np.asarray(f).ravel(order="F").astype(np.int32)
    if s.N == 0 or f.size % s.M != 0:
        raise EhmmError(3, "rejection control needs the count data of the session (Session.set_data)")
    if s.G == 0:
        G = int(f.max())
        s.set_groups(f.reshape(s.M, -1, order="F"), np.zeros(G), np.zeros(G))


def _with_stream(s: Session, call):
    s.rng_set_state(_rng.GLOBAL.state)
    n = call()
    _rng.GLOBAL.state[:] = s.rng_get_state()
    return n


def consensusRejectionControlled(hdf5, f, p) -> list:
    """src/consensusRejectionControlled.cpp:14-31 -> list of K (sum, id) tables."""
    s = session(hdf5)
    _ensure_groups(s, f)
    _with_stream(s, lambda: s.rc_consensus(p))
    return [s.rc_table(k) for k in range(s.K)]


def differentialRejectionControlled(hdf5, f, p, N) -> list:
    """src/differentialRejectionControlled.cpp:14-48 -> list of B+2 (sum, id) tables."""
    s = session(hdf5)
    if int(N) != s.N:
        raise EhmmError(1, "N=%d does not match the session's data (N=%d)" % (N, s.N))
    _ensure_groups(s, f)
    _with_stream(s, lambda: s.rc_differential(p))
    return [s.rc_table(c) for c in range(s.B + 2)]


def reweight(x, p) -> np.ndarray:
    """src/reweight.cpp:16-22."""
    x = np.asarray(x, dtype=np.float64).ravel()
    need = int(np.count_nonzero((x < p) & (x / p != 0))) if p > 0 else 0
    st = _rng.GLOBAL.state.copy()
    u = _rng.GLOBAL.runif(need)
    out, n = _reweight(x, p, u)
    if n != need:  # keep the stream exact even if the device disagrees on the count
        _rng.GLOBAL.state[:] = st
        _rng.GLOBAL.runif(n)
    return out


def rbinomVectorized(prob) -> np.ndarray:
    """src/rbinomVectorized.cpp:15-19: rbinom(1, p_i) in order == reweight(prob, 1)."""
    prob = np.asarray(prob, dtype=np.float64).ravel()
    if np.any((prob < 0) | (prob > 1)):
        raise EhmmError(1, "probabilities must lie in [0, 1]")
    return reweight(prob, 1.0)


def aggregate(x, f) -> np.ndarray:
    """src/aggregate.cpp:20-45."""
    return _aggregate(x, f)
