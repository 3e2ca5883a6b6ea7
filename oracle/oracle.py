"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(epigrahmm_b200) never does.  Function names mirror the reference's Rcpp
exports (R/RcppExports.R) and R helpers; see oracle/ehmm_oracle.c for the
file:line each one follows.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libehmm_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ehmm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.eo_dnbinom_mu.restype = C.c_double
        _lib.eo_dnbinom_mu.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
        _lib.eo_digamma.restype = C.c_double
        _lib.eo_digamma.argtypes = [C.c_double]
        _lib.eo_unif_rand.restype = C.c_double
        _lib.eo_qfunction.restype = C.c_double
        _lib.eo_bic.restype = C.c_double
        _lib.eo_glm_nb.restype = C.c_double
        _lib.eo_glm_zinb.restype = C.c_double
        _lib.eo_glm_mix_zinb.restype = C.c_double
        _lib.eo_glm_mix_nb.restype = C.c_double
        for n in ("eo_reweight", "eo_reweight_buf", "eo_aggregate_compact", "eo_consensus_rc", "eo_differential_rc", "eo_fdr_control"):
            getattr(_lib, n).restype = C.c_int64
    return _lib


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _colmajor(a):
    """Return a Fortran-ordered float64 copy (R / Armadillo layout)."""
    return np.asfortranarray(a, dtype=np.float64)


def _i32cm(a):
    return np.asfortranarray(a, dtype=np.int32)


class RNG(C.Structure):
    """R's Mersenne-Twister state: .Random.seed[2] = mti, [3:627] = mt."""

    _fields_ = [("mti", C.c_uint32), ("mt", C.c_uint32 * 624)]

    @classmethod
    def set_seed(cls, seed: int) -> "RNG":
        g = cls()
        lib().eo_set_seed(C.byref(g), C.c_uint32(seed & 0xFFFFFFFF))
        return g

    def unif_rand(self) -> float:
        return lib().eo_unif_rand(C.byref(self))

    def runif(self, n: int) -> np.ndarray:
        u = np.empty(n, dtype=np.float64)
        lib().eo_unif_fill(C.byref(self), _p(u), C.c_int64(n))
        return u

    def state(self) -> np.ndarray:
        """(625,) uint32: mti followed by the 624 state words."""
        return np.concatenate([[self.mti], np.ctypeslib.as_array(self.mt)]).astype(np.uint32)

    def copy(self) -> "RNG":
        g = RNG()
        C.memmove(C.byref(g), C.byref(self), C.sizeof(RNG))
        return g


def dnbinom(x, size, mu, log=False):
    x, size, mu = np.broadcast_arrays(np.asarray(x, float), np.asarray(size, float), np.asarray(mu, float))
    out = np.empty(x.shape)
    L = lib()
    for idx in np.ndindex(x.shape):
        out[idx] = L.eo_dnbinom_mu(x[idx], size[idx], mu[idx], int(log))
    return out


def digamma(x):
    x = np.asarray(x, float)
    out = np.empty(x.shape)
    L = lib()
    for idx in np.ndindex(x.shape):
        out[idx] = L.eo_digamma(x[idx])
    return out


# ---- native functions (src/*.cpp) ------------------------------------------------

def expStep(pi, gamma, logf):
    """src/expStep.cpp:45-123 -> dict of the five HDF5 datasets (column-major M x K)."""
    logf = _colmajor(logf)
    M, K = logf.shape
    pi = _f64(pi)
    gamma = _colmajor(gamma)
    out = {n: np.empty((M, K), order="F") for n in ("logFP", "logBP", "logProb1")}
    out["logProb2"] = np.empty((M, K * K), order="F")
    lib().eo_expstep(C.c_int64(M), K, _p(pi), _p(gamma), _p(logf), _p(out["logFP"]), _p(out["logBP"]),
                     _p(out["logProb1"]), _p(out["logProb2"]))
    out["logLikelihood"] = logf
    return out


def expStepTruth(pi, gamma, logf):
    """NOT the reference: the scaled, long-double forward-backward used as the accuracy yardstick.
    Returns (posteriors M x K, log-likelihood)."""
    logf = _colmajor(logf)
    M, K = logf.shape
    P1 = np.empty((M, K), order="F")
    f = lib().eo_expstep_truth
    f.restype = C.c_double
    ll = f(C.c_int64(M), K, _p(_f64(pi)), _p(_colmajor(gamma)), _p(logf), _p(P1))
    return P1, float(ll)


def maxStepProb(h):
    """src/maxStepProb.cpp:37-81 -> {'pi','gamma'}."""
    L1, L2 = _colmajor(h["logProb1"]), _colmajor(h["logProb2"])
    M, K = L1.shape
    pi = np.empty(K)
    gamma = np.empty((K, K), order="F")
    lib().eo_maxstepprob(C.c_int64(M), K, _p(L1), _p(L2), _p(pi), _p(gamma))
    return {"pi": pi, "gamma": gamma}


def computeQFunction(h, pi, gamma):
    lf, L1, L2 = _colmajor(h["logLikelihood"]), _colmajor(h["logProb1"]), _colmajor(h["logProb2"])
    M, K = L1.shape
    return lib().eo_qfunction(C.c_int64(M), K, _p(lf), _p(L1), _p(L2), _p(_f64(pi)), _p(_colmajor(gamma)))


def computeBIC(h, numPar, numSamples):
    F = _colmajor(h["logFP"])
    M, K = F.shape
    return lib().eo_bic(C.c_int64(M), K, _p(F), int(numPar), int(numSamples))


def innerMaxStepProb(h):
    eta, L1 = _colmajor(h["mixtureProb"]), _colmajor(h["logProb1"])
    M, B = eta.shape
    d = np.empty(B)
    lib().eo_inner_maxstepprob(C.c_int64(M), B, _p(eta), _p(L1), _p(d))
    return d


def computeViterbiSequence(h, pi, gamma):
    F = _colmajor(h["logFP"])
    M, K = F.shape
    S = np.empty(M)
    lib().eo_viterbi(C.c_int64(M), K, _p(F), _p(_f64(pi)), _p(_colmajor(gamma)), _p(S))
    return S


def getMarginalProbability(h):
    return np.exp(_colmajor(h["logProb1"]))


def reweight(x, p, rng: RNG | None = None, uniforms=None):
    """src/reweight.cpp:16-22.  Returns (weights, n_uniforms_consumed)."""
    x = _f64(x).copy()
    if uniforms is not None:
        u = _f64(uniforms)
        n = lib().eo_reweight_buf(_p(x), C.c_int64(x.size), C.c_double(p), _p(u), C.c_int64(u.size))
    else:
        n = lib().eo_reweight(_p(x), C.c_int64(x.size), C.c_double(p), C.byref(rng))
    return x, int(n)


def aggregate(x, f):
    """src/aggregate.cpp:20-45 -> (s x 2) matrix of (sum, id) for non-zero buckets."""
    x = _f64(x)
    f = np.ascontiguousarray(f, dtype=np.int32)
    G = int(f[: x.size].max()) if x.size else 0
    bucket = np.empty(G + 1)
    lib().eo_aggregate_dense(_p(x), C.c_int64(x.size), _p(f, C.c_int32), C.c_int64(G), _p(bucket))
    return _compact(bucket)


def _compact(bucket):
    ids = np.nonzero(bucket)[0]
    return np.column_stack([bucket[ids], ids.astype(float)])


def consensusRejectionControlled(h, f, p, rng: RNG | None = None, uniforms=None, dense=False):
    """src/consensusRejectionControlled.cpp:14-31.  Returns (list of K tables, n_uniforms)."""
    L1 = _colmajor(h["logProb1"])
    M, K = L1.shape
    f = np.ascontiguousarray(np.asarray(f).ravel(order="F"), dtype=np.int32)
    G = int(f.max())
    b = np.empty((K, G + 1))
    u = None if uniforms is None else _f64(uniforms)
    n = lib().eo_consensus_rc(C.c_int64(M), K, _p(L1), _p(f, C.c_int32), C.c_int64(G), C.c_double(p),
                              C.byref(rng) if rng is not None else None, _p(u),
                              C.c_int64(0 if u is None else u.size), _p(b))
    return (b if dense else [_compact(b[k]) for k in range(K)]), int(n)


def differentialRejectionControlled(h, f, p, N, rng: RNG | None = None, uniforms=None, dense=False):
    """src/differentialRejectionControlled.cpp:14-48.  Returns (list of B+2 tables, n_uniforms)."""
    L1, eta = _colmajor(h["logProb1"]), _colmajor(h["mixtureProb"])
    M, B = eta.shape
    f = np.ascontiguousarray(np.asarray(f).ravel(order="F"), dtype=np.int32)
    G = int(f.max())
    b = np.empty((B + 2, G + 1))
    u = None if uniforms is None else _f64(uniforms)
    n = lib().eo_differential_rc(C.c_int64(M), B, int(N), _p(L1), _p(eta), _p(f, C.c_int32), C.c_int64(G),
                                 C.c_double(p), C.byref(rng) if rng is not None else None, _p(u),
                                 C.c_int64(0 if u is None else u.size), _p(b))
    return (b if dense else [_compact(b[k]) for k in range(B + 2)]), int(n)


# ---- R-side numeric functions -------------------------------------------------------

def loglikConsensusHMM(y, offsets, psi, minZero, controls=None):
    """R/base-loglikelihood.R:104-122 (NB).  psi = [theta1, theta2]; controls already log1p'd."""
    y, off = _i32cm(y), _colmajor(offsets)
    M, N = y.shape
    ctrl = None if controls is None else _colmajor(controls)
    th1, th2 = _f64(psi[0]), _f64(psi[1])
    LL = np.empty((M, 2), order="F")
    lib().eo_loglik_consensus(C.c_int64(M), N, _p(y, C.c_int32), _p(off), _p(ctrl), _p(th1), _p(th2),
                              int(th1.size), C.c_double(minZero), _p(LL))
    return LL


def loglikConsensusHMMzinb(y, offsets, psi, minZero, controls=None):
    """R/base-loglikelihood.R:123-141 (dist = 'zinb').  psi[0] = (lambda0[, lambda_ctrl], beta0[, beta_ctrl],
    size) for the zero-inflated background, psi[1] = (beta0[, beta_ctrl], size)."""
    y, off = _i32cm(y), _colmajor(offsets)
    M, N = y.shape
    ctrl = None if controls is None else _colmajor(controls)
    th1, th2 = _f64(psi[0]), _f64(psi[1])
    assert th1.size == (3 if ctrl is None else 5) and th2.size == (2 if ctrl is None else 3)
    LL = np.empty((M, 2), order="F")
    lib().eo_loglik_consensus_zinb(C.c_int64(M), N, _p(y, C.c_int32), _p(off), _p(ctrl), _p(th1), _p(th2),
                                   int(th1.size), C.c_double(minZero), _p(LL))
    return LL


def loglikInit(y, offsets, psi):
    """R/base-EM.R:24-30 (single sample, dnbinom(log=TRUE))."""
    y, off = _i32cm(np.asarray(y).reshape(-1, 1)), _colmajor(np.asarray(offsets).reshape(-1, 1))
    M = y.shape[0]
    LL = np.empty((M, 2), order="F")
    lib().eo_loglik_init(C.c_int64(M), _p(y, C.c_int32), _p(off), _p(_f64(psi[0])), _p(_f64(psi[1])), _p(LL))
    return LL


def _design(design):
    return np.ascontiguousarray(design, dtype=np.uint8)  # (B, N) row-major


def loglikDifferential(y, offsets, design, delta, psi):
    y, off, d = _i32cm(y), _colmajor(offsets), _design(design)
    M, N = y.shape
    B = d.shape[0]
    ll = np.empty((M, B), order="F")
    lib().eo_loglik_differential(C.c_int64(M), N, B, _p(y, C.c_int32), _p(off), _p(d, C.c_uint8),
                                 _p(_f64(delta)), _p(_f64(psi)), _p(ll))
    return ll


def loglikDifferentialHMM(y, offsets, design, delta, psi, minZero):
    y, off, d = _i32cm(y), _colmajor(offsets), _design(design)
    M, N = y.shape
    B = d.shape[0]
    LL = np.empty((M, 3), order="F")
    lib().eo_loglik_differential_hmm(C.c_int64(M), N, B, _p(y, C.c_int32), _p(off), _p(d, C.c_uint8),
                                     _p(_f64(delta)), _p(_f64(psi)), C.c_double(minZero), _p(LL))
    return LL


def innerExpStep(y, offsets, design, delta, psi, minZero):
    y, off, d = _i32cm(y), _colmajor(offsets), _design(design)
    M, N = y.shape
    B = d.shape[0]
    eta = np.empty((M, B), order="F")
    lib().eo_inner_estep(C.c_int64(M), N, B, _p(y, C.c_int32), _p(off), _p(d, C.c_uint8),
                         _p(_f64(delta)), _p(_f64(psi)), C.c_double(minZero), _p(eta))
    return eta


def loglikDifferentialZINB(y, offsets, design, delta, psi, minZero):
    """R/base-loglikelihood.R:23-36; psi = (lambda, b_bg, l_bg, db, dl)"""
    y, off, d = _i32cm(y), _colmajor(offsets), _design(design)
    M, N = y.shape
    B = d.shape[0]
    ll = np.empty((M, B), order="F")
    lib().eo_loglik_differential_zinb(C.c_int64(M), N, B, _p(y, C.c_int32), _p(off), _p(d, C.c_uint8),
                                      _p(_f64(delta)), _p(_f64(psi)), C.c_double(minZero), _p(ll))
    return ll


def loglikDifferentialHMMzinb(y, offsets, design, delta, psi, minZero):
    """R/base-loglikelihood.R:72-96"""
    y, off, d = _i32cm(y), _colmajor(offsets), _design(design)
    M, N = y.shape
    B = d.shape[0]
    LL = np.empty((M, 3), order="F")
    lib().eo_loglik_differential_hmm_zinb(C.c_int64(M), N, B, _p(y, C.c_int32), _p(off), _p(d, C.c_uint8),
                                          _p(_f64(delta)), _p(_f64(psi)), C.c_double(minZero), _p(LL))
    return LL


def innerExpStepZINB(y, offsets, design, delta, psi, minZero):
    y, off, d = _i32cm(y), _colmajor(offsets), _design(design)
    M, N = y.shape
    B = d.shape[0]
    eta = np.empty((M, B), order="F")
    lib().eo_inner_estep_zinb(C.c_int64(M), N, B, _p(y, C.c_int32), _p(off), _p(d, C.c_uint8),
                              _p(_f64(delta)), _p(_f64(psi)), C.c_double(minZero), _p(eta))
    return eta


def glmMixZINB(par, id, weights, y, offsets, dsg, maxid, minZero, grad=True):
    """R/base-optim.R:261-362 -> (value, gradient); par = (lambda, b_bg, l_bg, b_enr, l_enr)."""
    dsg = np.asfortranarray(dsg, dtype=np.uint8)
    n, B = dsg.shape
    g = np.empty(5)
    idv = np.ascontiguousarray(id, dtype=np.int32)
    v = lib().eo_glm_mix_zinb(C.c_int64(n), B, int(maxid), _p(_f64(par)), _p(idv, C.c_int32), _p(_f64(weights)),
                              _p(_f64(y)), _p(_f64(offsets)), _p(dsg, C.c_uint8), C.c_double(minZero),
                              _p(g) if grad else None)
    return v, g


def glmNB(par, y, x, offset, weights, grad=True):
    """R/base-optim.R:109-130 -> (value, gradient)."""
    x = _colmajor(np.asarray(x, float).reshape(len(y), -1) if len(y) else np.zeros((0, 1)))
    G, P = x.shape
    g = np.empty(P + 1)
    v = lib().eo_glm_nb(C.c_int64(G), P, _p(_f64(par)), _p(_f64(y)), _p(x), _p(_f64(offset)), _p(_f64(weights)),
                        _p(g) if grad else None)
    return v, g


def glmZINB(par, y, x, offset, weights, minZero, grad=True):
    """R/base-optim.R:136-181 -> (value, gradient); par = (beta[P], phi, lambda[P])."""
    x = _colmajor(np.asarray(x, float).reshape(len(y), -1) if len(y) else np.zeros((0, 1)))
    G, P = x.shape
    g = np.empty(2 * P + 1)
    v = lib().eo_glm_zinb(C.c_int64(G), P, _p(_f64(par)), _p(_f64(y)), _p(x), _p(_f64(offset)), _p(_f64(weights)),
                          C.c_double(minZero), _p(g) if grad else None)
    return v, g


def glmMixNB(par, id, weights, y, offsets, dsg, maxid, minZero, grad=True):
    """R/base-optim.R:187-255 -> (value, gradient).  dsg: n x B 0/1."""
    dsg = np.asfortranarray(dsg, dtype=np.uint8)
    n, B = dsg.shape
    g = np.empty(4)
    idv = np.ascontiguousarray(id, dtype=np.int32)
    v = lib().eo_glm_mix_nb(C.c_int64(n), B, int(maxid), _p(_f64(par)), _p(idv, C.c_int32), _p(_f64(weights)),
                            _p(_f64(y)), _p(_f64(offsets)), _p(dsg, C.c_uint8), C.c_double(minZero),
                            _p(g) if grad else None)
    return v, g


def fdrControl(prob, fdr=0.05):
    """R/base-utils.R:11-25 with R's long double cumsum: boolean vector in window order."""
    prob = _f64(prob)
    out = np.empty(prob.size, dtype=np.uint8)
    n = lib().eo_fdr_control(C.c_int64(prob.size), _p(prob), C.c_double(fdr), out.ctypes.data_as(C.POINTER(C.c_ubyte)))
    if n < 0:
        raise ValueError("Posterior probabilities must be between 0 and 1")
    return out.astype(bool)
