"""Synthetic genome-wide count matrices: a K-state Markov chain over
bins with negative-binomial counts per sample and offsets rounded to 3 decimals."""
from __future__ import annotations

import numpy as np

GAMMA2 = np.array([[0.999, 0.001], [0.005, 0.995]])
GAMMA3 = np.array([[0.998, 0.001, 0.001], [0.004, 0.992, 0.004], [0.002, 0.002, 0.996]])

# fixed parameters for the timed iteration
THETA_CONSENSUS = dict(pi=np.array([0.999, 0.001]), gamma=np.array([[0.998, 0.002], [0.01, 0.99]]),
                       psi=[np.array([np.log(2.0), 3.0]), np.array([np.log(12.0), 4.0])])
THETA_DIFFERENTIAL = dict(pi=np.array([0.999, 0.0005, 0.0005]),
                          gamma=np.array([[0.998, 0.001, 0.001], [0.005, 0.99, 0.005], [0.001, 0.001, 0.998]]),
                          psi=np.array([np.log(2.0), np.log(3.0), np.log(6.0), np.log(4.0 / 3.0)]))


def enumerate_patterns(n_cond: int) -> np.ndarray:
    """0/1 patterns over conditions ordered as R's expand.grid + order(rowSums), dropping all-0
    and all-1 (R/base-utils.R:198-204, determineMixtures :228-229).  Returns (B, n_cond)."""
    grid = np.array([[(r >> c) & 1 for c in range(n_cond)] for r in range(2 ** n_cond)], dtype=np.uint8)
    order = np.argsort(grid.sum(1), kind="stable")
    return grid[order][1:-1]


def markov_chain(M: int, gamma: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """Hidden path by run-length sampling (geometric sojourn times), O(#runs)."""
    K = gamma.shape[0]
    z = np.empty(M, dtype=np.int8)
    pos, state = 0, 0
    while pos < M:
        stay = gamma[state, state]
        run = int(rng.geometric(1.0 - stay))
        z[pos:pos + run] = state
        pos += run
        p = gamma[state].copy()
        p[state] = 0.0
        state = int(rng.choice(K, p=p / p.sum()))
    return z


def make_consensus(M: int, N: int, seed: int, controls: bool = False):
    rng = np.random.default_rng(seed)
    z = markov_chain(M, GAMMA2, rng)
    mu = np.where(z == 1, 12.0, 2.0)[:, None]
    size = np.where(z == 1, 4.0, 3.0)[:, None]
    offsets = np.round(rng.normal(0.0, 0.1, (M, N)), 3)
    ctrl = None
    if controls:
        ctrl = np.log1p(rng.poisson(5.0, (M, N)).astype(np.float64))
    lam = rng.gamma(shape=np.broadcast_to(size, (M, N)), scale=np.broadcast_to(mu / size, (M, N)) * np.exp(offsets))
    counts = rng.poisson(lam).astype(np.int32)
    return dict(counts=np.asfortranarray(counts), offsets=np.asfortranarray(offsets), controls=ctrl, z=z)


def make_differential(M: int, n_cond: int, n_rep: int, seed: int):
    rng = np.random.default_rng(seed)
    N = n_cond * n_rep
    pats = enumerate_patterns(n_cond)
    B = pats.shape[0]
    z = markov_chain(M, GAMMA3, rng)
    comp = rng.integers(0, B, M)
    cond_of = np.repeat(np.arange(n_cond), n_rep)
    enriched = np.where((z == 2)[:, None], 1, np.where((z == 1)[:, None], pats[comp][:, cond_of], 0))
    mu = np.where(enriched == 1, 12.0, 2.0)
    size = np.where(enriched == 1, 4.0, 3.0)
    offsets = np.round(rng.normal(0.0, 0.1, (M, N)), 3)
    lam = rng.gamma(shape=size, scale=mu / size * np.exp(offsets))
    counts = rng.poisson(lam).astype(np.int32)
    design = pats[:, cond_of].astype(np.uint8)  # B x N
    return dict(counts=np.asfortranarray(counts), offsets=np.asfortranarray(offsets), design=design, z=z, B=B)


def make_groups(counts, offsets, controls=None, design=None):
    """dt$Group and dtUnique (R/base-core.R:97-104,187-198): 1-based ids in order of first
    appearance over the column-major long table, keyed by (ChIP[, controls], offsets[, Dsg.Mix*]).
    Each key column is ranked on its own and the ranks are packed into one int64 so that a single
    1-D sort finds the groups (M*N can be 1e8)."""
    counts = np.asarray(counts)
    M, N = counts.shape
    cols = [counts.ravel(order="F").astype(np.float64)]
    if controls is not None:
        cols.append(np.asarray(controls, dtype=np.float64).ravel(order="F"))
    cols.append(np.asarray(offsets, dtype=np.float64).ravel(order="F"))
    key = np.zeros(M * N, dtype=np.int64)
    for c in cols:
        u, inv = np.unique(c, return_inverse=True)
        if key.max(initial=0) > (2 ** 62) // max(u.size, 1):
            raise OverflowError("too many distinct key values to pack")
        key = key * u.size + inv.ravel()
    if design is not None:
        # the Dsg.Mix columns depend on the sample only: one code per distinct design column
        _, dcode = np.unique(np.asarray(design).T, axis=0, return_inverse=True)
        dcode = np.asarray(dcode).ravel()
        key = key * (int(dcode.max()) + 1) + np.repeat(dcode.astype(np.int64), M)
    _, first, inv = np.unique(key, return_index=True, return_inverse=True)
    inv = np.asarray(inv).ravel()
    order = np.argsort(first, kind="stable")          # groups numbered by first appearance (.GRP)
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    group = (rank[inv] + 1).astype(np.int32)
    rows = first[order]
    out = dict(group=np.asfortranarray(group.reshape(M, N, order="F")), G=int(order.size),
               u_chip=cols[0][rows].copy(), u_offsets=cols[-1][rows].copy())
    out["u_controls"] = cols[1][rows].copy() if controls is not None else None
    if design is not None:
        sample_of = rows // M
        out["u_dsg"] = np.asfortranarray(np.asarray(design, dtype=np.uint8).T[sample_of])
    else:
        out["u_dsg"] = None
    return out
