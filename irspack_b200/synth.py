"""Deterministic synthetic interaction matrices of the BASELINE.json shapes.

There is no network for MovieLens / Netflix, so the benchmark and the parity
tests run on seeded synthetic CSR matrices with the named shapes (SURVEY.md
section 8 d).  Host-side numpy only (data plumbing, not the hot path).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import scipy.sparse as sps

# name -> (n_users, n_items, nnz, K)       BASELINE.json `configs`, SURVEY.md 8
SHAPES: Dict[str, Tuple[int, int, int, int]] = {
    "ml1m": (6040, 3706, 1_000_209, 64),
    "ml20m": (138_493, 26_744, 20_000_263, 128),
    "netflix": (480_189, 17_770, 100_480_507, 256),
    "powerlaw1b": (10_000_000, 2_000_000, 1_000_000_000, 128),
}


def synth_csr(n_users: int, n_items: int, nnz: int, seed: int, law: str = "power",
              values: str = "ones") -> sps.csr_matrix:
    """CSR float32 with exactly ``nnz`` distinct (user, item) pairs, indices ascending.

    ``law="power"``: user activity ~ lognormal(sigma=1), item popularity
    ~ (rank + n_items/400)^-1 in a random item order (an ML-20M-like head: the
    most popular item holds ~0.3% of the interactions; some users end up empty).
    ``law="uniform"``: both uniform.  ``values``: "ones" or "counts" (1..5).
    """
    if nnz > n_users * n_items:
        raise ValueError("nnz exceeds the matrix size")
    rng = np.random.default_rng(seed)
    if law == "power":
        pu = rng.lognormal(0.0, 1.0, n_users)
        pi = 1.0 / (np.arange(n_items) + max(n_items / 400.0, 1.0))
        pi = pi[rng.permutation(n_items)]
    elif law == "uniform":
        pu = np.ones(n_users)
        pi = np.ones(n_items)
    else:
        raise ValueError("law must be 'power' or 'uniform'")
    cu = np.cumsum(pu / pu.sum())
    ci = np.cumsum(pi / pi.sum())
    cu[-1] = ci[-1] = 1.0
    keys = np.empty(0, dtype=np.int64)
    while keys.size < nnz:
        m = int((nnz - keys.size) * 1.25) + 1024
        rows = np.searchsorted(cu, rng.random(m), side="right").astype(np.int64)
        cols = np.searchsorted(ci, rng.random(m), side="right").astype(np.int64)
        keys = np.unique(np.concatenate([keys, rows * n_items + cols]))
    if keys.size > nnz:
        keep = np.sort(rng.choice(keys.size, nnz, replace=False))
        keys = keys[keep]
    rows = keys // n_items
    cols = (keys - rows * n_items).astype(np.int32)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n_users), out=indptr[1:])
    if values == "ones":
        data = np.ones(nnz, dtype=np.float32)
    elif values == "counts":
        data = rng.integers(1, 6, nnz).astype(np.float32)
    else:
        raise ValueError("values must be 'ones' or 'counts'")
    X = sps.csr_matrix((data, cols, indptr), shape=(n_users, n_items))
    X.has_sorted_indices = True
    return X


def init_factors(n: int, K: int, seed: int, init_std: float = 0.1) -> np.ndarray:
    """Initial factors set into BOTH implementations through the ``user`` / ``item``
    setters so that ``Solver::initialize`` is bypassed (SURVEY.md 8 d)."""
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, K), dtype=np.float32) * np.float32(init_std / np.sqrt(K)))


def holdout_split(X: sps.csr_matrix, test_fraction: float, seed: int
                  ) -> Tuple[sps.csr_matrix, sps.csr_matrix]:
    """Random per-interaction split into (train, test) CSR matrices of X's shape."""
    rng = np.random.default_rng(seed)
    X = sps.csr_matrix(X)
    to_test = rng.random(X.nnz) < test_fraction
    rows = np.repeat(np.arange(X.shape[0]), np.diff(X.indptr))

    def take(sel: np.ndarray) -> sps.csr_matrix:
        M = sps.csr_matrix((X.data[sel], (rows[sel], X.indices[sel])), shape=X.shape)
        M.sort_indices()
        return M

    return take(~to_test), take(to_test)
