"""Callers either side of the EM loop: R/base-core.R (initializerHMM, consensusHMM,
differentialHMM) and the initialisation helpers of R/base-utils.R, on plain arrays.

These run once per fit on the host (SURVEY.md section 8f ranks them "next" after the loop body);
the loop body itself is :mod:`epigrahmm_b200.em` over a :class:`~epigrahmm_b200.session.Session`.
``counts`` / ``offsets`` are M x N matrices whose columns are sorted by condition, ``condition`` the
per-column condition labels (``colData(object)$condition``).
"""
from __future__ import annotations

import math

import numpy as np

from . import em
from .session import Session


def _pattern_table(n_cond):
    """ref of enumeratePatterns (R/base-utils.R:196-204): expand.grid of 0/1 ordered by row sum"""
    grid = np.array([[(r >> c) & 1 for c in range(n_cond)] for r in range(2 ** n_cond)], dtype=np.uint8)
    return grid[np.argsort(grid.sum(1), kind="stable")]


def enumeratePatterns(peaks, condition):
    """Group (1-based row of the pattern table) of every window from the initial peak calls."""
    cond = list(dict.fromkeys(condition))
    chain = np.column_stack([(peaks[:, [i for i, c in enumerate(condition) if c == u]].sum(1) > 0) for u in cond])
    ref = _pattern_table(len(cond))
    code = {tuple(r): i + 1 for i, r in enumerate(ref.tolist())}
    group = np.array([code[tuple(r)] for r in chain.astype(np.uint8).tolist()])
    return group, ref


def estimateCoefficients(z, counts, offsets, control, type_, dist="nb"):
    """R/base-utils.R:282-320"""
    r = (np.asarray(counts, dtype=np.float64) + 1.0) / np.exp(offsets)
    par = []
    for x in (z.min(), z.max()):
        v = r[z == x].ravel()
        mu, s2 = v.mean(), v.var(ddof=1)
        disp = control["maxDisp"] if s2 <= mu else min(mu ** 2 / (s2 - mu), control["maxDisp"])
        zip_ = max(0.01, (s2 - mu) / (s2 + mu ** 2 - mu))
        par.append((mu, disp, zip_))
    if type_ == "differential":
        p1 = np.log(np.array(par[0][:2]))
        p2 = np.log(np.array(par[1][:2])) - p1
        return [p1, p2] if dist == "nb" else [np.concatenate([[math.log(par[0][2])], p1]), p2]
    first = [math.log(par[0][0]), par[0][1]]
    if dist == "zinb":
        first = [math.log(par[0][2])] + first
    return [np.array(first), np.array([math.log(par[1][0]), par[1][1]])]


def initializer(counts, offsets, control, device=0, key="initializer"):
    """initializer() (R/initializer.R:42-60): one 2-state fit per sample; returns the M x N 0/1
    'peaks' assay.  The N fits are independent (one Session each, reusing the hot path with N=1)."""
    counts = np.asarray(counts)
    M, N = counts.shape
    peaks = np.zeros((M, N))
    for i in range(N):
        with Session("%s.%d" % (key, i), M, 2, device) as s:
            c, o = counts[:, [i]], np.asarray(offsets)[:, [i]]
            s.set_data(c, o)
            s.build_groups()          # dt$Group / dtUnique on the device (R/base-core.R:30-35)
            s.rng_set_state(_stream().state)
            peaks[:, i], _ = em.initializerHMM(c, o, s, control)
            _stream().state[:] = s.rng_get_state()
    return peaks


def _stream():
    from . import rng
    return rng.GLOBAL


def consensusHMM(counts, offsets, peaks, condition, control, controls=None, device=0, key=None, session=None, dist="nb"):
    """R/base-core.R:163-228.  Returns dict(session, theta, history)."""
    counts = np.asarray(counts)
    M, N = counts.shape
    z, _ = enumeratePatterns(np.asarray(peaks), condition)
    ctrl = None if controls is None else np.log1p(np.asarray(controls, dtype=np.float64))
    s = session or Session(key or "epigraHMM.h5", M, 2, device)
    s.set_data(counts, offsets, ctrl)
    s.build_groups()                  # R/base-core.R:184-198 on the device
    psi = estimateCoefficients(z, counts, offsets, control, "consensus", dist)
    if ctrl is not None:   # a zero coefficient after every intercept (R/base-utils.R:305-316, ifControls)
        psi = [np.concatenate([np.ravel([[v, 0.0] for v in p[:-1]]), p[-1:]]) for p in psi]
    theta = dict(pi=np.array([0.999, 0.001]), gamma=em.estimateTransitionProb(z, 2), psi=psi)
    s.rng_set_state(_stream().state)
    fit = em.emConsensus(theta, s, N, control, dist=dist)
    _stream().state[:] = s.rng_get_state()
    th = fit["theta_new"]
    s.viterbi(th["pi"], th["gamma"], fetch=False)
    return dict(session=s, theta=th, history=fit)


def differentialHMM(counts, offsets, peaks, condition, control, device=0, key=None, session=None, dist="nb"):
    """R/base-core.R:69-157 (control$pattern = NULL -> all 2^G - 2 patterns)."""
    counts = np.asarray(counts)
    M, N = counts.shape
    cond = list(dict.fromkeys(condition))
    nG = len(cond)
    z, ref = enumeratePatterns(np.asarray(peaks), condition)
    if control.get("pattern"):
        want = []
        for pat in control["pattern"]:  # 1-based condition indices, as in R
            row = np.zeros(nG, dtype=np.uint8)
            row[[p - 1 for p in pat]] = 1
            want.append(int(np.where((ref == row).all(1))[0][0]) + 1)
        zseq = want
    else:
        zseq = list(range(2, 2 ** nG))
    B = len(zseq)
    cond_of = np.array([cond.index(c) for c in condition])
    design = np.stack([ref[zq - 1][cond_of] for zq in zseq]).astype(np.uint8)  # B x N
    s = session or Session(key or "epigraHMM.h5", M, 3, device)
    s.set_data(counts, offsets)
    s.set_design(design)
    s.build_groups()                  # R/base-core.R:97-104 on the device
    zmin, zmax = z.min(), z.max()
    chain = np.where(z == zmin, 0, np.where(z == zmax, 2, 1))
    diff = z[(z != 1) & (z != zmax)]
    sub = diff[np.isin(diff, zseq)]
    delta = np.array([(sub == zq).sum() / max(sub.size, 1) for zq in zseq], dtype=np.float64)
    theta = dict(pi=np.array([0.999, 0.0005, 0.0005]), gamma=em.estimateTransitionProb(chain, 3), delta=delta,
                 psi=estimateCoefficients(z, counts, offsets, control, "differential", dist))
    s.rng_set_state(_stream().state)
    fit = em.outerEMDifferential(theta, s, N, control, dist=dist)
    _stream().state[:] = s.rng_get_state()
    th = fit["theta_new"]
    s.viterbi(th["pi"], th["gamma"], fetch=False)
    patterns = ["".join("E" if v else "B" for v in ref[zq - 1]) for zq in zseq]
    return dict(session=s, theta=th, history=fit, mixturePatterns=patterns, design=design)


def _components(fit):
    """metadata(object)$components (R/base-core.R:145-151): per mixture component the 1-based conditions it
    marks as enriched and its posterior proportion (the last delta)."""
    enr = [[i + 1 for i, ch in enumerate(p) if ch == "E"] for p in fit["mixturePatterns"]]
    return enr, np.asarray(fit["theta"]["delta"], dtype=np.float64)


def prunePatterns(counts, offsets, peaks, condition, control, dist="nb", **kw):
    """prunePatterns() (R/base-utils.R:398-447): fit the full differential model, then drop the rarest
    combinatorial pattern and refit while any posterior proportion is below control$pruningThreshold.
    Returns the last fit with `pruned` = the patterns removed, in order, and `fullBIC`."""
    threshold = control["pruningThreshold"]
    if threshold is None:
        raise ValueError("control['pruningThreshold'] is NULL")
    control = dict(control, pruningThreshold=None)
    fit = differentialHMM(counts, offsets, peaks, condition, control, dist=dist, **kw)
    full_bic = fit["history"]["controlHist"][-2]["bic"]
    pruned = []
    enr, post = _components(fit)
    while np.any(post < threshold):
        fit["session"].close()                       # unlink(metadata(fitFull)$output)
        idx = int(np.argmin(post))                   # which.min: the first of the rarest
        pruned.append(enr[idx])
        control = dict(control, pattern=[e for i, e in enumerate(enr) if i != idx])
        fit = differentialHMM(counts, offsets, peaks, condition, control, dist=dist, **kw)
        enr, post = _components(fit)
    fit["pruned"], fit["fullBIC"] = pruned, full_bic
    fit["control"] = dict(control, pruningThreshold=threshold)
    return fit


def epigraHMM(counts, offsets, peaks, condition, control, type="consensus", dist="nb", **kw):
    """epigraHMM() (R/epigraHMM.R:39-90) on arrays."""
    if dist not in ("nb", "zinb"):
        raise ValueError("dist must be 'nb' or 'zinb'")
    if type == "consensus":
        return consensusHMM(counts, offsets, peaks, condition, control, dist=dist, **kw)
    if type == "differential":
        if control.get("pruningThreshold") is not None:   # R/epigraHMM.R:82-84
            return prunePatterns(counts, offsets, peaks, condition, control, dist=dist, **kw)
        return differentialHMM(counts, offsets, peaks, condition, control, dist=dist, **kw)
    raise ValueError("type must be 'consensus' or 'differential'")
