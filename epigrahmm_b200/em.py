"""EM drivers: the loops of R/base-EM.R over a device-resident session.

``emConsensus``, ``emInitialization``, ``outerEMDifferential`` and ``innerEMDifferential`` keep the
reference's iteration order, bookkeeping (``controlList`` / ``getError`` / ``checkConvergence``,
R/base-utils.R:31-45,326-341; R/base-checkers.R:11-17) and parameterisations; every numeric step
inside the loop body is one call into the CUDA library through ``backend`` (a
:class:`~epigrahmm_b200.session.Session`, or any object with the same methods -- the parity tests
drive the CPU oracle through this same loop).  The NB-GLM maximisation uses L-BFGS-B with R's
``optim`` defaults (lmm=5, factr=1e7, pgtol=0, maxit=100) over the library's value+gradient
reductions (R/base-optim.R:11-103).
"""
from __future__ import annotations

import copy
import math
import time

import numpy as np
from scipy.optimize import minimize

_EPS = np.finfo(np.float64).eps
XMIN = 2.2250738585072014e-308


def controlEM(epsilonEM=None, maxIterEM=500, minIterEM=3, gapIterEM=3, maxCountEM=3, maxDisp=1000, criterion="all",
              minZero=XMIN, probCut=0.05, quiet=True, maxIterInnerEM=5, epsilonInnerEM=1e-3, trimOffset=3,
              pattern=None, tempDir=None, fileName="epigraHMM", pruningThreshold=None, quietPruning=True):
    """R/controlEM.R:44-98 (same 18 fields, same validation)."""
    eps = {"MRCPE": 1e-3, "MACPE": 1e-3, "ARCEL": 1e-3}
    if epsilonEM is not None:
        for k, v in dict(epsilonEM).items():
            if k in eps:
                eps[k] = v
    if not all(v > 0 for v in eps.values()):
        raise ValueError("epsilonEM must be a positive numerical value (or vector)")
    if maxIterEM % 1 or maxIterEM <= 0 or minIterEM % 1 or minIterEM <= 0:
        raise ValueError("maxIterEM and minIterEM must be a positive integer")
    if gapIterEM % 1 or gapIterEM <= 0 or gapIterEM > minIterEM:
        raise ValueError("value of 'gapIterEM' must be a positive integer <= minIterEM")
    if maxCountEM % 1 or maxCountEM <= 0:
        raise ValueError("value of 'maxCountEM' must be a positive integer")
    if not maxDisp > 0:
        raise ValueError("value of 'maxDisp' must be > 0")
    if criterion not in ("MRCPE", "MACPE", "ARCEL", "ACC", "all"):
        raise ValueError("value of 'criterion' must be 'MRCPE', 'MACPE', 'ARCEL', 'ACC', or 'all'")
    if not minZero > 0:
        raise ValueError("'minZero' must be a positive numeric value")
    if not 0 < probCut < 1:
        raise ValueError("'probCut' must be a value between 0 and 1")
    if maxIterInnerEM % 1 or maxIterInnerEM <= 0:
        raise ValueError("value of 'maxIterInnerEM' must be a positive integer")
    if not epsilonInnerEM > 0:
        raise ValueError("'epsilonInnerEM' must be a positive value")
    if pattern is not None and not isinstance(pattern, list):
        raise ValueError("'pattern' must be NULL or a list'")
    if pruningThreshold is not None and not 0 < pruningThreshold < 1:
        raise ValueError("value of 'pruningThreshold' must be between (0,1)")
    return dict(epsilonEM=eps, minZero=minZero, maxIterEM=maxIterEM, minIterEM=minIterEM, gapIterEM=gapIterEM,
                maxCountEM=maxCountEM, maxDisp=maxDisp, criterion=criterion, probCut=probCut, quiet=quiet,
                maxIterInnerEM=maxIterInnerEM, epsilonInnerEM=epsilonInnerEM, trimOffset=trimOffset, pattern=pattern,
                tempDir=tempDir, fileName=fileName, pruningThreshold=pruningThreshold, quietPruning=quietPruning)


_CRIT = ("MRCPE", "MACPE", "ARCEL")


def controlList():
    """R/base-utils.R:326-341"""
    return dict(time=0.0, iteration=0, convergence=0, bic=0.0, q=0.0, error=np.zeros(3), count=np.zeros(3))


def unlist(theta) -> np.ndarray:
    """unlist(theta): pi, gamma (column-major), [delta,] psi -- the order parHist rows have in R"""
    parts = [np.ravel(theta["pi"]), np.ravel(np.asarray(theta["gamma"]), order="F")]
    if "delta" in theta and theta["delta"] is not None:
        parts.append(np.ravel(theta["delta"]))
    parts.extend(np.ravel(p) for p in theta["psi"])
    return np.concatenate(parts)


def checkConvergence(controlHist, control) -> float:
    cnt = controlHist[-1]["count"]
    if control["criterion"] == "all":
        return float(np.max(cnt))
    return float(cnt[_CRIT.index(control["criterion"])])


def getError(controlHist, parHist, control) -> np.ndarray:
    """R/base-utils.R:31-45"""
    it = controlHist[-1]["iteration"]
    if it <= 1:
        return np.zeros(3)
    gap = control["gapIterEM"] if it > control["minIterEM"] else 1
    old, new = unlist(parHist[it - gap - 1]), unlist(parHist[it - 1])
    q_old, q_new = controlHist[it - gap - 1]["q"], controlHist[it - 1]["q"]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.array([np.max(np.abs((new - old) / old)), np.max(np.abs(new - old)), abs((q_new - q_old) / q_old)])


def invertDispersion(par, model="nb"):
    """R/base-utils.R:170-181.  'nb': (beta..., size) <-> (beta..., 1/size), an involution.  'zinb': the model's
    (lambda[P], beta[P], size) -> the optimiser's (beta[P], 1/size, lambda[P])."""
    par = np.asarray(par, dtype=np.float64)
    if model == "nb":
        return np.concatenate([par[:-1], [1.0 / par[-1]]])
    P = (par.size - 1) // 2
    return np.concatenate([par[P:2 * P], [1.0 / par[2 * P]], par[:P]])


def _zinb_from_optim(x):
    """optimNB's back-transformation for dist = 'zinb' (R/base-optim.R:43-47)"""
    x = np.asarray(x, dtype=np.float64)
    P = (x.size - 1) // 2
    return np.concatenate([x[P + 1:], x[:P], [1.0 / x[P]]])


def _lbfgsb_minimize(fun, x0, lower, upper):
    """L-BFGS-B through scipy.optimize.minimize (the portable route)."""
    bounds = [(None if not np.isfinite(lo) else lo, None if not np.isfinite(hi) else hi) for lo, hi in zip(lower, upper)]
    res = minimize(fun, np.asarray(x0, dtype=np.float64), jac=True, method="L-BFGS-B", bounds=bounds,
                   options=dict(maxcor=5, ftol=1e7 * _EPS, gtol=0.0, maxiter=100, maxfun=15000, maxls=20))
    return res.x, {0: 0, 1: 1}.get(res.status, 52)


class _Setulb:
    """One L-BFGS-B instance driven through scipy's reverse-communication routine `setulb`
    (the optimiser core of scipy.optimize.minimize, without its per-call Python scaffolding:
    identical iterates, ~4x less host time).  request() runs the routine until it needs f and g at a
    point (returned) or has finished (None); feed() supplies them."""

    _lib = None

    @staticmethod
    def bounds(lower, upper):
        """(nbd, low, upp, lower, upper) in setulb's form; shared by instances with the same bounds"""
        lower, upper = np.asarray(lower, dtype=np.float64), np.asarray(upper, dtype=np.float64)
        lo_f, hi_f = np.isfinite(lower), np.isfinite(upper)
        nbd = np.where(lo_f & hi_f, 2, np.where(lo_f, 1, np.where(hi_f, 3, 0))).astype(_lbfgsb_int_dtype())
        return nbd, np.where(lo_f, lower, 0.0), np.where(hi_f, upper, 0.0), lower, upper

    def __init__(self, x0, lower=None, upper=None, bounds=None):
        if _Setulb._lib is None:
            from scipy.optimize import _lbfgsb
            _Setulb._lib = _lbfgsb
        m, n = 5, len(x0)
        self.nbd, self.low, self.upp, lower, upper = bounds if bounds is not None else _Setulb.bounds(lower, upper)
        idt = self.nbd.dtype
        self.x = np.clip(np.array(x0, dtype=np.float64), lower, upper)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros(n, dtype=np.float64)
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self.iwa = np.zeros(3 * n, dtype=idt)
        self.task, self.ln_task, self.lsave = np.zeros(2, dtype=idt), np.zeros(2, dtype=idt), np.zeros(4, dtype=idt)
        self.isave, self.dsave = np.zeros(44, dtype=idt), np.zeros(29, dtype=np.float64)
        self.nit = self.nfev = 0
        self.code = None

    def request(self):
        """-> the point (a view of the internal x: copy it before the next call) or None when finished"""
        setulb, task = self._lib.setulb, self.task
        while True:
            setulb(5, self.x, self.low, self.upp, self.nbd, self.f, self.g, 1e7, 0.0, self.wa, self.iwa,
                   task, self.lsave, self.isave, self.dsave, 20, self.ln_task)
            t = task[0]
            if t == 3:
                return self.x
            if t == 1:
                self.nit += 1
                if self.nit >= 100:
                    task[0], task[1] = 5, 504
                elif self.nfev > 15000:
                    task[0], task[1] = 5, 502
                continue
            break
        if t == 4:
            self.code = 0
        elif self.nit >= 100 or self.nfev > 15000:
            self.code = 1
        else:
            self.code = 52
        return None

    def feed(self, fv, gv):
        self.nfev += 1
        self.f[...] = fv
        self.g[:] = gv


def _lbfgsb_setulb(fun, x0, lower, upper):
    st = _Setulb(x0, lower, upper)
    while True:
        x = st.request()
        if x is None:
            return st.x, st.code
        st.feed(*fun(x.copy()))


def _lbfgsb_int_dtype():
    try:
        from scipy.optimize._lbfgsb_py import HAS_ILP64
        return np.int64 if HAS_ILP64 else np.int32
    except Exception:
        return np.int32


_FAST_LBFGSB = None


def _lbfgsb(fun, x0, lower, upper):
    """stats::optim(method='L-BFGS-B') defaults: lmm=5, factr=1e7, pgtol=0, maxit=100.
    Returns (par, convergence) with R's codes (0 ok, 1 maxit, 52 error from the routine)."""
    global _FAST_LBFGSB

    def f(x):
        v, g = fun(x)
        if not np.isfinite(v):
            raise FloatingPointError("L-BFGS-B needs finite values of 'fn'")
        return v, g

    if _FAST_LBFGSB is None:
        # use the direct driver only if this scipy exposes the routine with the expected signature
        try:
            def quad(x):
                return float(((x - 1.0) ** 2).sum()), 2.0 * (x - 1.0)
            a, ca = _lbfgsb_setulb(quad, np.array([0.3, 2.0]), [-np.inf, 1.5], [np.inf, np.inf])
            b, cb = _lbfgsb_minimize(quad, np.array([0.3, 2.0]), [-np.inf, 1.5], [np.inf, np.inf])
            _FAST_LBFGSB = bool(np.array_equal(a, b) and ca == cb)
        except Exception:
            _FAST_LBFGSB = False
    return (_lbfgsb_setulb if _FAST_LBFGSB else _lbfgsb_minimize)(f, x0, lower, upper)


def optimNB(par, backend, column, control):
    """R/base-optim.R:11-27 (dist='nb'): par = (beta..., size) -> optimise over (beta..., phi=1/size)."""
    par = np.asarray(par, dtype=np.float64)
    start = invertDispersion(par)
    P = par.size - 1
    try:
        lower = np.concatenate([np.full(P, -np.inf), [1.0 / control["maxDisp"]]])
        x, conv = _lbfgsb(lambda x: backend.glm_nb(column, x), start, lower, np.full(P + 1, np.inf))
    except Exception:  # tryCatch(error=): keep the starting value, convergence 99
        x, conv = start, 99
    return dict(par=invertDispersion(x), convergence=conv)


def optimZINB(par, backend, column, control):
    """R/base-optim.R:28-48 (dist='zinb'): par = (lambda[P], beta[P], size) -> optimise (beta, phi, lambda)."""
    par = np.asarray(par, dtype=np.float64)
    start = invertDispersion(par, "zinb")
    P = (par.size - 1) // 2
    try:
        lower = np.concatenate([np.full(P, -np.inf), [1.0 / control["maxDisp"]], np.full(P, -np.inf)])
        x, conv = _lbfgsb(lambda x: backend.glm_zinb(column, x, control["minZero"]), start, lower, np.full(2 * P + 1, np.inf))
    except Exception:
        x, conv = start, 99
    return dict(par=_zinb_from_optim(x), convergence=conv)


def optimNB_all(pars, backend, control, dist="nb"):
    """optimNB for every state (R/base-EM.R:131-141 runs them one after the other with lapply).
    With dist = 'zinb' the first state is the zero-inflated one (l.140) and is optimised on its own.
    The K problems are independent, so when the backend can evaluate several (state, parameter)
    pairs in one launch (glm_nb_batch) their L-BFGS-B instances advance in lock-step: each one sees
    exactly the evaluations it would see alone, and the host waits for K times fewer launches."""
    if dist == "zinb":
        return [optimZINB(pars[0], backend, 0, control)] + [optimNB(p, backend, k, control) for k, p in enumerate(pars) if k > 0]
    if not (hasattr(backend, "glm_nb_batch") and _fast_lbfgsb_available()):
        return [optimNB(p, backend, k, control) for k, p in enumerate(pars)]
    starts = [invertDispersion(p) for p in pars]
    P = len(starts[0]) - 1
    lower = np.concatenate([np.full(P, -np.inf), [1.0 / control["maxDisp"]]])
    upper = np.full(P + 1, np.inf)
    out = [None] * len(pars)
    states, req = {}, {}
    bounds = _Setulb.bounds(lower, upper)
    for k, x0 in enumerate(starts):
        try:
            states[k] = _Setulb(x0, bounds=bounds)
            req[k] = states[k].request()
        except Exception:
            out[k] = dict(par=invertDispersion(x0), convergence=99)
    X = np.empty((len(pars), P + 1))   # the batch of points of one round (request() returns views)
    while True:
        ks = []
        for k, st in states.items():
            if out[k] is None:
                if req[k] is None:
                    out[k] = dict(par=invertDispersion(st.x), convergence=st.code)
                else:
                    X[len(ks)] = req[k]
                    ks.append(k)
        if not ks:
            break
        try:
            vals, grads = backend.glm_nb_batch(ks, X[:len(ks)])
        except Exception:
            for k in ks:
                out[k] = dict(par=invertDispersion(starts[k]), convergence=99)
            continue
        for i, k in enumerate(ks):
            if not math.isfinite(vals[i]):  # optim(): "L-BFGS-B needs finite values of 'fn'" -> tryCatch(error=)
                out[k] = dict(par=invertDispersion(starts[k]), convergence=99)
                continue
            try:
                st = states[k]
                st.feed(vals[i], grads[i])
                req[k] = st.request()
            except Exception:
                out[k] = dict(par=invertDispersion(starts[k]), convergence=99)
    return out


def _fast_lbfgsb_available():
    if _FAST_LBFGSB is None:
        _lbfgsb(lambda x: (float(((x - 1.0) ** 2).sum()), 2.0 * (x - 1.0)), np.array([0.5]), [-np.inf], [np.inf])
    return bool(_FAST_LBFGSB)


def optimDifferential(par, backend, control):
    """R/base-optim.R:58-79 (dist='nb'): par = unlist(psi) = (b, l, db, dl); optimiser order (b, db, l, dl)."""
    par = np.asarray(par, dtype=np.float64)
    perm = [0, 2, 1, 3]
    try:
        lower = np.full(4, math.log(control["minZero"]))
        upper = np.array([np.inf, np.inf, math.log(control["maxDisp"]), math.log(control["maxDisp"])])
        x, conv = _lbfgsb(lambda x: backend.glm_mix_nb(x, control["minZero"]), par[perm], lower, upper)
    except Exception:
        x, conv = par, 99  # the reference permutes the untouched par as well (R/base-optim.R:73-77)
    return dict(par=np.asarray(x)[perm], convergence=conv)


def optimDifferentialZINB(par, backend, control):
    """R/base-optim.R:80-101 (dist='zinb'): par = unlist(psi) = (lambda, b, l, db, dl); the optimiser sees
    (lambda, b, l, b + db, l + dl)."""
    par = np.asarray(par, dtype=np.float64)
    newpar = np.array([par[0], par[1], par[2], par[1] + par[3], par[2] + par[4]])
    try:
        lower = np.full(5, math.log(control["minZero"]))
        upper = np.array([np.inf, np.inf, math.log(control["maxDisp"]), np.inf, math.log(control["maxDisp"])])
        x, conv = _lbfgsb(lambda x: backend.glm_mix_zinb(x, control["minZero"]), newpar, lower, upper)
    except Exception:
        x, conv = par, 99  # the reference back-transforms the untouched par as well (R/base-optim.R:98-100)
    x = np.asarray(x)
    return dict(par=np.array([x[0], x[1], x[2], x[3] - x[1], x[4] - x[2]]), convergence=conv)


def _rejection_cut(iteration, control):
    rejectionCut = max(0.9 ** iteration, control["probCut"])
    return (control["probCut"] > 0) * rejectionCut


def _hint(backend, p):
    """the iteration's rejection cut is known before its posteriors are computed (it depends on the iteration number
    only, R/base-EM.R:119-121): announce it, so that the kernels producing the posteriors count for it on the way"""
    if hasattr(backend, "rc_hint"):
        backend.rc_hint(p)


def _lean(backend, on):
    """inside the EM loops nothing reads logFP, logBP or logProb2 (R/base-EM.R:113-158, 191-225): the E-steps there
    write the posteriors and the statistics only; the datasets of the last E-step appear when they are asked for"""
    if hasattr(backend, "expstep_mode"):
        backend.expstep_mode(on)


def _update_hist(controlHist, parHist, theta_new, control, init_time, bic, q, conv):
    cur = controlHist[-1]
    parHist.append(copy.deepcopy(theta_new))
    eps = np.array([control["epsilonEM"][c] for c in _CRIT])
    cur["count"] = (cur["error"] <= eps) * (cur["iteration"] > control["minIterEM"]) * (cur["count"] + 1) + 0.0
    cur["time"] = (time.time() - init_time) / 60.0
    cur["bic"] = bic
    cur["q"] = q
    cur["convergence"] = conv
    cur["error"] = getError(controlHist, parHist, control)
    controlHist.append(copy.deepcopy(cur))


def _continue(controlHist, control):
    return checkConvergence(controlHist, control) < control["maxCountEM"] and \
        controlHist[-1]["iteration"] < control["maxIterEM"]


def emConsensus(theta_old, backend, N, control, dist="nb", callback=None):
    """R/base-EM.R:98-178.  theta = dict(pi, gamma, psi=[theta1, theta2]); psi_k = (beta0[, beta_ctrl], size);
    with dist = 'zinb' theta1 = (lambda0[, lambda_ctrl], beta0[, beta_ctrl], size)."""
    if dist not in ("nb", "zinb"):
        raise ValueError("dist must be 'nb' or 'zinb'")
    loglik = backend.loglik_consensus if dist == "nb" else backend.loglik_consensus_zinb
    controlHist, parHist = [controlList()], []
    theta_old = copy.deepcopy(theta_old)
    theta_new = copy.deepcopy(theta_old)
    numPar = unlist(theta_old).size - 3
    _lean(backend, True)
    t0 = time.time()
    while _continue(controlHist, control):
        cur = controlHist[-1]
        cur["iteration"] += 1
        _hint(backend, _rejection_cut(cur["iteration"], control))
        loglik(theta_old["psi"], control["minZero"])
        backend.expstep(theta_old["pi"], theta_old["gamma"])
        theta_new.update(backend.maxstepprob())
        backend.rc_consensus(_rejection_cut(cur["iteration"], control))
        optimMLE = optimNB_all(theta_old["psi"], backend, control, dist)
        theta_new["psi"] = [o["par"] for o in optimMLE]
        bic = backend.bic(numPar, N)
        q = backend.qfunction(theta_old["pi"], theta_old["gamma"])
        conv = 1 * any(o["convergence"] != 0 for o in optimMLE)
        _update_hist(controlHist, parHist, theta_new, control, t0, bic, q, conv)
        theta_old = copy.deepcopy(theta_new)
        if callback:
            callback(controlHist[-2], theta_new)
    _lean(backend, False)
    return dict(theta_new=theta_new, controlHist=controlHist, parHist=parHist)


def emInitialization(theta_old, backend, control, callback=None):
    """R/base-EM.R:12-92 (one sample, K = 2, dnbinom(log=TRUE) emission)."""
    controlHist, parHist = [controlList()], []
    theta_old = copy.deepcopy(theta_old)
    theta_new = copy.deepcopy(theta_old)
    numPar = unlist(theta_old).size - 3
    _lean(backend, True)
    t0 = time.time()
    while _continue(controlHist, control):
        cur = controlHist[-1]
        cur["iteration"] += 1
        _hint(backend, _rejection_cut(cur["iteration"], control))
        backend.loglik_init(theta_old["psi"])
        backend.expstep(theta_old["pi"], theta_old["gamma"])
        theta_new.update(backend.maxstepprob())
        backend.rc_consensus(_rejection_cut(cur["iteration"], control))
        optimMLE = optimNB_all(theta_old["psi"], backend, control)
        theta_new["psi"] = [o["par"] for o in optimMLE]
        # (the reference updates count/convergence/error before bic/q here; same values result)
        bic = backend.bic(numPar, 1)
        q = backend.qfunction(theta_old["pi"], theta_old["gamma"])
        conv = 1 * any(o["convergence"] != 0 for o in optimMLE)
        _update_hist(controlHist, parHist, theta_new, control, t0, bic, q, conv)
        theta_old = copy.deepcopy(theta_new)
        if callback:
            callback(controlHist[-2], theta_new)
    _lean(backend, False)
    return dict(theta_new=theta_new, controlHist=controlHist, parHist=parHist)


def innerEMDifferential(theta_old, backend, probCut, control, dist="nb"):
    """R/base-EM.R:184-228.  psi = [(b, l), (db, dl)]; with dist = 'zinb' [(lambda, b, l), (db, dl)]."""
    inner_expstep = backend.inner_expstep if dist == "nb" else backend.inner_expstep_zinb
    nfirst = 2 if dist == "nb" else 3
    count, error = 0, 1.0
    tmp_old = copy.deepcopy(theta_old)
    tmp_new = copy.deepcopy(theta_old)
    optimMLE = None
    _hint(backend, probCut)
    while count < control["maxIterInnerEM"] and error > control["epsilonInnerEM"]:
        psi_flat = np.concatenate([np.ravel(p) for p in tmp_old["psi"]])
        inner_expstep(tmp_old["delta"], psi_flat, control["minZero"])
        tmp_new["delta"] = backend.inner_maxstepprob()
        backend.rc_differential(probCut)
        optimMLE = (optimDifferential if dist == "nb" else optimDifferentialZINB)(psi_flat, backend, control)
        tmp_new["psi"] = [optimMLE["par"][:nfirst].copy(), optimMLE["par"][nfirst:].copy()]
        new = np.concatenate([tmp_new["delta"], np.concatenate(tmp_new["psi"])])
        old = np.concatenate([np.ravel(tmp_old["delta"]), psi_flat])
        with np.errstate(divide="ignore", invalid="ignore"):
            error = float(np.max(np.abs((new - old) / old)))
        tmp_old = copy.deepcopy(tmp_new)
        count += 1
    return dict(theta_new=dict(delta=tmp_new["delta"], psi=tmp_new["psi"]), optimMLE=optimMLE)


def outerEMDifferential(theta_old, backend, N, control, dist="nb", callback=None):
    """R/base-EM.R:234-296.  theta = dict(pi, gamma, delta, psi=[(b, l), (db, dl)])."""
    if dist not in ("nb", "zinb"):
        raise ValueError("dist must be 'nb' or 'zinb'")
    loglik = backend.loglik_differential_hmm if dist == "nb" else backend.loglik_differential_hmm_zinb
    controlHist, parHist = [controlList()], []
    theta_old = copy.deepcopy(theta_old)
    theta_new = copy.deepcopy(theta_old)
    npat = len(control["pattern"]) if control.get("pattern") else 0
    numPar = unlist(theta_old).size - npat - 3
    _lean(backend, True)
    t0 = time.time()
    while _continue(controlHist, control):
        cur = controlHist[-1]
        cur["iteration"] += 1
        psi_flat = np.concatenate([np.ravel(p) for p in theta_old["psi"]])
        _hint(backend, _rejection_cut(cur["iteration"], control))
        loglik(theta_old["delta"], psi_flat, control["minZero"])
        backend.expstep(theta_old["pi"], theta_old["gamma"])
        theta_new.update(backend.maxstepprob())
        inner = innerEMDifferential(theta_old, backend, _rejection_cut(cur["iteration"], control), control, dist)
        theta_new["delta"] = inner["theta_new"]["delta"]
        theta_new["psi"] = inner["theta_new"]["psi"]
        bic = backend.bic(numPar, N)
        q = backend.qfunction(theta_old["pi"], theta_old["gamma"])
        _update_hist(controlHist, parHist, theta_new, control, t0, bic, q, inner["optimMLE"]["convergence"])
        theta_old = copy.deepcopy(theta_new)
        if callback:
            callback(controlHist[-2], theta_new)
    _lean(backend, False)
    return dict(theta_new=theta_new, controlHist=controlHist, parHist=parHist)


# ---- callers either side of the loop (SURVEY.md 8f "next" rows; host-side, small) -----------------

def checkProbabilities(P):
    """R/base-checkers.R:34-40"""
    P = np.clip(np.asarray(P, dtype=np.float64), 0.0, 1.0)
    P[P == 0] = XMIN
    return P / P.sum(axis=1, keepdims=True)


def estimateTransitionProb(chain, numStates):
    """R/estimateTransitionProb.R:27-37 (chain values 0..numStates-1 after rescaling)."""
    chain = np.asarray(chain, dtype=np.float64)
    nm1 = numStates - 1
    if chain.max() > nm1:
        chain = np.round(nm1 * (chain - chain.max()) / (chain.max() - chain.min()) + nm1)
    c = chain.astype(np.int64)
    levels_from, levels_to = np.unique(c[:-1]), np.unique(c[1:])
    tab = np.zeros((levels_from.size, levels_to.size))
    np.add.at(tab, (np.searchsorted(levels_from, c[:-1]), np.searchsorted(levels_to, c[1:])), 1.0)
    return checkProbabilities(tab / tab.sum(axis=1, keepdims=True))


def fdrControl(prob, fdr=0.05):
    """R/base-utils.R:11-25: windows whose running mean of sorted (1 - prob) stays below fdr."""
    prob = np.asarray(prob, dtype=np.float64)
    if not np.all((prob >= 0) & (prob <= 1)):
        raise ValueError("Posterior probabilities must be between 0 and 1")
    if not 0 < fdr < 1:
        raise ValueError("fdr must be between 0 and 1")
    notpp = 1.0 - prob
    order = np.argsort(notpp, kind="stable")
    # R's cumsum adds in long double and stores doubles (src/main/cum.c); numpy's longdouble is that type on x86
    csum = np.cumsum(notpp[order].astype(np.longdouble)).astype(np.float64)
    ok = csum / np.arange(1, prob.size + 1) < fdr
    out = np.zeros(prob.size, dtype=bool)
    out[order] = ok
    return out


def callPeaks(backend, method="viterbi"):
    """R/callPeaks.R:61-62 without the GRanges post-processing: boolean peak index per bin."""
    if hasattr(backend, "call_peaks"):      # sort, prefix sum and comparison on the device
        return backend.call_peaks(method)
    if method == "viterbi":
        return backend.fetch("viterbi")[:, 0] == 1
    return fdrControl(np.exp(backend.fetch("logProb1")[:, 1]), fdr=float(method))


def callPatterns(backend, peaks, patterns, type="all", fdr=None, pattern=None, ranges=None):
    """R/callPaterns.R:78-112 without the GRanges bookkeeping.  `peaks` (and `ranges`) are boolean masks
    over the bins (the windows overlapping the peak / range set); `patterns` names the mixture
    components (differentialHMM's mixturePatterns, or getPatterns' condition names).  Returns the bin
    indices kept and, per type: 'all' / 'ranges' the M' x B posterior matrix, 'fdr' a boolean vector for
    `pattern`, 'max' the name of the most probable pattern per window (which.max: first of ties)."""
    post = backend.fetch("mixtureProb")
    peaks = np.asarray(peaks, dtype=bool)
    if peaks.shape != (post.shape[0],):
        raise ValueError("peaks must flag every bin")
    if len(patterns) != post.shape[1]:
        raise ValueError("one name per mixture component")
    idx = np.nonzero(peaks)[0]
    post = post[idx]
    if type == "all":
        return idx, post
    if type == "fdr":
        if pattern not in list(patterns):
            raise ValueError("unknown pattern %r" % (pattern,))
        return idx, fdrControl(post[:, list(patterns).index(pattern)], fdr=fdr)
    if type == "max":
        names = np.asarray(list(patterns), dtype=object)
        return idx, names[np.argmax(post, axis=1)]
    if type == "ranges":
        keep = np.asarray(ranges, dtype=bool)[idx]
        return idx[keep], post[keep]
    raise ValueError("type must be 'all', 'fdr', 'max' or 'ranges'")


def initial_psi_single_sample(counts, offsets, maxDisp, z):
    """initializerHMM's moment estimates (R/base-core.R:41-49) for the two groups of z (1/2)."""
    r = (np.asarray(counts, dtype=np.float64) + 1.0) / np.exp(offsets)
    psi = []
    for x in (1, 2):
        v = r[z == x]
        mu, s2 = v.mean(), v.var(ddof=1)
        psi.append(np.array([math.log(mu), min(mu ** 2 / max(0.0, s2 - mu) if s2 > mu else np.inf, maxDisp)]))
    return psi


def initializerHMM(counts, offsets, backend, control, callback=None):
    """R/base-core.R:11-63 for one sample: quantile split, moment initialisation, EM, Viterbi."""
    counts = np.asarray(counts).ravel()
    offsets = np.asarray(offsets, dtype=np.float64).ravel()
    score = np.log1p(counts / np.exp(offsets))
    score = (score - score.mean()) / score.std(ddof=1)
    z = np.where(score > np.quantile(score, 0.75), 2, 1)  # cut(..., right = TRUE)
    theta = dict(pi=np.array([0.999, 0.001]), gamma=estimateTransitionProb(z, 2),
                 psi=initial_psi_single_sample(counts, offsets, control["maxDisp"], z))
    fit = emInitialization(theta, backend, control, callback=callback)
    th = fit["theta_new"]
    return backend.viterbi(th["pi"], th["gamma"]), fit
