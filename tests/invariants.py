"""The reference's closed-form iALS tests, restated once and run against both
back ends (tests/backends.py).  Each function cites the reference test it
restates (/root/reference/tests/recommenders/test_ials.py)."""
from __future__ import annotations

import math

import numpy as np
import pytest
import scipy.sparse as sps


def ials_grad(X, u, v, reg, alpha0, epsilon=None):
    """Brute-force gradient of the ORIGINAL loss (test_ials.py:19-51)."""
    weight = (lambda x: x) if epsilon is None else (lambda x: math.log(1 + x / epsilon))
    uv = u.dot(v.T)
    gu, gv = np.zeros_like(u), np.zeros_like(v)
    Xd = np.asarray(X.todense())
    for a in range(u.shape[0]):
        for b in range(v.shape[0]):
            x = Xd[a, b]
            sc = alpha0 * uv[a, b] if x == 0 else (alpha0 + weight(x)) * (uv[a, b] - 1)
            gu[a] += v[b] * sc
            gv[b] += u[a] * sc
    return gu + reg * u, gv + reg * v


def binarised(X):
    Xd = np.asarray(X.todense(), dtype=np.float64)
    Xd[Xd.nonzero()] = 1.0
    return Xd


def overfit_cholesky(Backend, X):  # test_ials.py:54-76
    t = Backend(X, K=4, alpha0=100, reg=1e-1, nu=0, loss_type="ORIGINAL", solver="CHOLESKY")
    for _ in range(100):
        t.step()
    u, i = t.transform_user(X), t.transform_item(X)
    np.testing.assert_allclose(u.dot(i.T), binarised(X), rtol=1e-2, atol=1e-2)


def overfit_cg(Backend, X):  # test_ials.py:516-548
    t = Backend(X, K=3, alpha0=100, reg=1e-1, nu=0, loss_type="ORIGINAL", solver="CG",
                max_cg_steps=3)
    for _ in range(100):
        t.step()
    u, i = t.transform_user(X), t.transform_item(X)
    np.testing.assert_allclose(u.dot(i.T), binarised(X), rtol=1e-2, atol=1e-2)
    np.testing.assert_allclose(u.dot(t.item.T), binarised(X), rtol=1e-2, atol=1e-2)
    with pytest.raises(ValueError):
        t.transform_item(sps.csr_matrix(X.T))
    np.testing.assert_allclose(t.user.dot(i.T), binarised(X), rtol=1e-2, atol=1e-2)


def overfit_ialspp(Backend, X, subspace_dimension, epochs=300):  # test_ials.py:573-599
    t = Backend(X, K=4, alpha0=100, reg=1.0, nu=0, loss_type="ORIGINAL", solver="IALSPP",
                subspace=subspace_dimension)
    for _ in range(epochs):
        t.step()
    np.testing.assert_allclose(t.user.dot(t.item.T), binarised(X), rtol=1e-2, atol=1e-2)


def loss_identity(Backend, X, loss_type, alpha0):  # test_ials.py:456-513
    reg = 1e-1
    t = Backend(X, K=2, alpha0=alpha0, reg=reg, nu=0, loss_type=loss_type, solver="CHOLESKY")
    t.step()
    t.step()
    u, v = t.user.astype(np.float64), t.item.astype(np.float64)
    ui = u.dot(v.T)
    row, col = X.nonzero()
    data = np.asarray(X[row, col]).ravel()
    if loss_type == "ORIGINAL":
        manual = (data + alpha0).dot((ui[row, col] - 1) ** 2)
        ui[row, col] = 0.0
        manual += alpha0 * ui.ravel().dot(ui.ravel())
    else:
        manual = data.dot((ui[row, col] - 1) ** 2)
        manual += alpha0 * ui.ravel().dot(ui.ravel())
    manual += reg * ((u ** 2).sum() + (v ** 2).sum())
    manual /= 2
    assert t.compute_loss() == pytest.approx(manual, rel=2e-6)


def user_scores_batching(Backend):  # test_ials.py:551-570
    rng = np.random.default_rng(0)
    n_users, n_items, K = 513, 257, 31
    X = sps.csr_matrix((n_users, n_items), dtype=np.float32)
    t = Backend(X, K=K, alpha0=0.1, reg=0.1, nu=1.0, loss_type="IALSPP", solver="CG")
    user = rng.standard_normal((n_users, K)).astype(np.float32)
    item = rng.standard_normal((n_items, K)).astype(np.float32)
    t.user, t.item = user, item
    for b, e in [(0, n_users), (17, 193), (n_users, n_users)]:
        got = t.user_scores(b, e)
        assert got.shape == (e - b, n_items)
        np.testing.assert_allclose(got, user[b:e] @ item.T, rtol=2e-5, atol=2e-5)


def cg_matches_cholesky(Backend, X):  # test_ials.py:627-661
    out = {}
    for solver, steps in (("CHOLESKY", 3), ("CG", 5)):
        t = Backend(X, K=4, alpha0=1 / 4.5, reg=3, nu=1.0, loss_type="IALSPP", solver=solver,
                    max_cg_steps=steps)
        for _ in range(5):
            t.step()
        out[solver] = (t.transform_user(X), t.transform_item(X))
    np.testing.assert_allclose(out["CHOLESKY"][0], out["CG"][0], atol=1e-3, rtol=1e-4)
    np.testing.assert_allclose(out["CHOLESKY"][1], out["CG"][1], atol=1e-3, rtol=1e-4)


def stationary_point_logscale(Backend, X, atol=1e-5):  # test_ials.py:664-697
    alpha0, reg, eps, K = 2.4, 1.1, 3.0, 5
    # nu = 0 and nu_star = 0: scaled_reg == reg (ials.py:412-418)
    t = Backend(X, K=K, alpha0=alpha0, reg=reg, nu=0, loss_type="ORIGINAL", solver="CHOLESKY",
                epsilon=eps)
    for _ in range(200):
        t.step()
    u = t.user.astype(np.float64)
    i_cold = t.transform_item(X).astype(np.float64)
    gu, gi = ials_grad(X, u, i_cold, reg, alpha0, eps)
    np.testing.assert_allclose(gi, np.zeros_like(gi), atol=atol)
    np.testing.assert_allclose(gu, np.zeros_like(gu), atol=atol)
