#!/usr/bin/env python
"""Generates tests/golden/ials_small.npz: golden vectors for the iALS hot path.

The reference itself cannot be built or imported in this image (Eigen 5.0.1, nanobind, optuna
are absent), and it ships no fixture files for this path.  These vectors
therefore come from a SECOND, independent restatement of the reference's arithmetic -- plain
numpy float64, row-by-row Python loops, written from the formulas of
/root/reference/cpp_source/als/IALSTrainer.hpp (prepare_p :78-115, compute_reg :117-120,
step_cg :170-271, step_cholesky :273-331, step_ialspp :387-535, step :784-788, compute_loss
:836-940) and
/root/reference/cpp_source/evaluator.cpp:324-355 (top-k order) -- not from oracle/.  The
C++ oracle (float64 twin) must reproduce them to 1e-9, the float32 oracle and the CUDA path to
the float32 tolerances of tests/test_gpu_parity.py.

    python tests/golden/make_golden.py          # rewrites ials_small.npz (deterministic)
"""
import os

import numpy as np
import scipy.sparse as sps

HERE = os.path.dirname(os.path.abspath(__file__))
U, I, K = 90, 70, 12
ALPHA0, REG, NU = 0.1, 0.05, 1.0
CG_STEPS = 3
EPOCHS = 2


def make_inputs():
    rng = np.random.default_rng(20261017)
    dense = (rng.random((U, I)) < 0.12) * rng.integers(1, 6, size=(U, I))
    dense[7, :] = 0          # a user without interactions
    dense[:, 11] = 0         # an item without interactions
    X = sps.csr_matrix(dense.astype(np.float64))
    X.sort_indices()
    user = rng.standard_normal((U, K)) * (0.1 / np.sqrt(K))
    item = rng.standard_normal((I, K)) * (0.1 / np.sqrt(K))
    # the trainers take float32 inputs: round once so that every implementation starts equal
    return X, user.astype(np.float32).astype(np.float64), item.astype(np.float32).astype(np.float64)


def gram(Y):  # :78-115
    return ALPHA0 * (Y.T @ Y)


def reg_of(n_other, n_u):  # :117-120
    return REG * (ALPHA0 * n_other + n_u) ** NU


def solve_cg(target, X, other, bias):  # :170-271
    P = gram(other)
    out = target.copy()
    for u in range(X.shape[0]):
        s, e = X.indptr[u], X.indptr[u + 1]
        if e == s:
            out[u] = 0.0     # :207-210
            continue
        idx, c = X.indices[s:e], X.data[s:e]
        Y = other[idx]
        reg_u = reg_of(other.shape[0], e - s)
        x = out[u].copy()
        b = ((bias + c)[:, None] * Y).sum(axis=0)
        r = b - (P @ x + Y.T @ (c * (Y @ x)) + reg_u * x)
        p = r.copy()
        r2 = r @ r
        if r2 <= 1e-20:
            out[u] = x
            continue
        for _ in range(CG_STEPS):
            Ap = P @ p + Y.T @ (c * (Y @ p)) + reg_u * p
            den = p @ Ap
            assert den > 0
            alpha = r2 / den
            x = x + alpha * p
            r = r - alpha * Ap
            r2n = r @ r
            if r2n <= 1e-20:
                break
            p = r + (r2n / r2) * p
            r2 = r2n
        out[u] = x
    return out


def solve_cholesky(target, X, other, bias):  # :273-331
    P = gram(other)
    out = target.copy()
    for u in range(X.shape[0]):
        s, e = X.indptr[u], X.indptr[u + 1]
        if e == s:
            out[u] = 0.0
            continue
        idx, c = X.indices[s:e], X.data[s:e]
        Y = other[idx]
        A = P + (Y * c[:, None]).T @ Y + reg_of(other.shape[0], e - s) * np.eye(K)
        b = ((bias + c)[:, None] * Y).sum(axis=0)
        L = np.linalg.cholesky(A)
        out[u] = np.linalg.solve(L.T, np.linalg.solve(L, b))
    return out


IALSPP_SUBSPACE, IALSPP_ITERATIONS = 5, 2


def solve_ialspp(target, X, other, bias):  # :387-535 (_prediction, _step_dimrange, step_ialspp)
    P = gram(other)
    out = target.copy()
    for u in range(X.shape[0]):
        s, e = X.indptr[u], X.indptr[u + 1]
        idx, c = X.indices[s:e], X.data[s:e]
        Y = other[idx]
        reg_u = reg_of(other.shape[0], e - s)
        x = out[u].copy()
        for _ in range(IALSPP_ITERATIONS):
            pred = Y @ x                                     # :387-424
            for d0 in range(0, K, IALSPP_SUBSPACE):          # :526-533
                d1 = min(d0 + IALSPP_SUBSPACE, K)
                Ys = Y[:, d0:d1]
                A = P[d0:d1, d0:d1] + (Ys * c[:, None]).T @ Ys + reg_u * np.eye(d1 - d0)
                B = P[d0:d1, :] @ x + reg_u * x[d0:d1] + Ys.T @ (c * (pred - 1.0) - bias)
                delta = np.linalg.solve(A, B)                # :497-498
                x[d0:d1] -= delta                            # :499-500
                pred = pred - Ys @ delta                     # :502-508
        out[u] = x
    return out


def epochs(X, user, item, solver, bias, n):
    Xt = sps.csr_matrix(X.T)
    Xt.sort_indices()
    for _ in range(n):  # :784-788
        user = solver(user, X, item, bias)
        item = solver(item, Xt, user, bias)
    return user, item


def loss(X, user, item, bias):  # ials.py:252-258 (the model), IALSTrainer.hpp:836-940
    S = user @ item.T
    D = X.toarray()
    obs = D != 0
    w = (bias + D)[obs]
    n_u = np.diff(X.indptr)
    n_i = np.diff(sps.csr_matrix(X.T).indptr)
    val = 0.5 * (w * (S[obs] - 1.0) ** 2).sum() + 0.5 * ALPHA0 * (S ** 2).sum()
    if bias != 0:  # ORIGINAL: the observed cells carry alpha0 + c, not alpha0 on top of it
        val -= 0.5 * ALPHA0 * (S[obs] ** 2).sum()
    val += 0.5 * REG * (((ALPHA0 * I + n_u) ** NU) * (user ** 2).sum(axis=1)).sum()
    val += 0.5 * REG * (((ALPHA0 * U + n_i) ** NU) * (item ** 2).sum(axis=1)).sum()
    return val


def topk(user, item, X, k):  # evaluator.py:426-432 + evaluator.cpp:324-355
    S = user @ item.T
    S[X.nonzero()] = -np.inf
    out = np.full((U, k), -1, dtype=np.int32)
    for u in range(U):
        cand = [j for j in range(I) if S[u, j] != -np.inf]
        cand.sort(key=lambda j: (-S[u, j], j))
        out[u, :min(k, len(cand))] = cand[:k]
    return out


def main():
    X, u0, i0 = make_inputs()
    out = dict(indptr=X.indptr.astype(np.int64), indices=X.indices.astype(np.int32), data=X.data,
               shape=np.array([U, I, K]), hyper=np.array([ALPHA0, REG, NU, CG_STEPS, EPOCHS]),
               user0=u0, item0=i0)
    for name, solver in (("cg", solve_cg), ("chol", solve_cholesky)):
        for lt, bias in (("ialspp", 0.0), ("original", ALPHA0)):
            uu, ii = epochs(X, u0, i0, solver, bias, EPOCHS)
            out[f"user_{name}_{lt}"], out[f"item_{name}_{lt}"] = uu, ii
            out[f"loss_{name}_{lt}"] = np.array(loss(X, uu, ii, bias))
    uu, ii = out["user_cg_ialspp"], out["item_cg_ialspp"]
    out["gram_item0"] = gram(i0)
    out["scores_cg_ialspp"] = uu @ ii.T
    out["top10_cg_ialspp"] = topk(uu, ii, X, 10)
    np.savez_compressed(os.path.join(HERE, "ials_small.npz"), **out)
    print("wrote", os.path.join(HERE, "ials_small.npz"))
    # iALS++: a separate file, same inputs
    pp = dict(subspace=np.array(IALSPP_SUBSPACE), iterations=np.array(IALSPP_ITERATIONS))
    for lt, bias in (("ialspp", 0.0), ("original", ALPHA0)):
        uu, ii = epochs(X, u0, i0, solve_ialspp, bias, EPOCHS)
        pp[f"user_{lt}"], pp[f"item_{lt}"] = uu, ii
        pp[f"loss_{lt}"] = np.array(loss(X, uu, ii, bias))
    np.savez_compressed(os.path.join(HERE, "ialspp_small.npz"), **pp)
    print("wrote", os.path.join(HERE, "ialspp_small.npz"))


if __name__ == "__main__":
    main()
