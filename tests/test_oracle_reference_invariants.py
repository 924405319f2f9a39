"""Pins the CPU oracle (oracle/) against the reference's own closed-form tests.

The reference holds no golden vectors for the iALS path and cannot be built here
(Eigen / nanobind absent), so these tolerance-level invariants -- restated from
/root/reference/tests/recommenders/test_ials.py and
/root/reference/tests/evaluation/test_evaluator.py -- are what anchors the oracle.
"""
import numpy as np
import pytest
import scipy.sparse as sps

import oracle
from backends import OracleBackend

import invariants as inv


def f64(*a, **k):
    return OracleBackend(*a, dtype=np.float64, **k)


@pytest.mark.parametrize("Backend", [OracleBackend, f64], ids=["f32", "f64"])
def test_overfit_cholesky(Backend, X_small):
    inv.overfit_cholesky(Backend, X_small)


@pytest.mark.parametrize("Backend", [OracleBackend, f64], ids=["f32", "f64"])
def test_overfit_cg(Backend, X_small):
    inv.overfit_cg(Backend, X_small)


@pytest.mark.parametrize("subspace_dimension", [1, 2, 3, 4])
def test_overfit_ialspp(X_small, subspace_dimension):  # test_ials.py:573-599
    inv.overfit_ialspp(OracleBackend, X_small, subspace_dimension)
    inv.overfit_ialspp(f64, X_small, subspace_dimension)


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-5), (np.float64, 1e-12)])
def test_ialspp_with_one_full_block_is_the_cholesky_step(dtype, tol):
    """With subspace >= K and one iteration, _step_dimrange solves the full normal equations:
    x - A^-1 (A x - b) = A^-1 b (IALSTrainer.hpp:474-500 against :296-324)."""
    from irspack_b200.synth import synth_csr

    X = synth_csr(300, 200, 6000, seed=5, values="counts")
    for loss in (oracle.LOSS_ORIGINAL, oracle.LOSS_IALSPP):
        a = oracle.OracleTrainer(X, 24, 0.1, 0.05, 1.0, loss, dtype=dtype)
        b = oracle.OracleTrainer(X, 24, 0.1, 0.05, 1.0, loss, dtype=dtype)
        a.ialspp_subspace_dimension = 64
        a.step(oracle.SOLVER_IALSPP)
        b.step(oracle.SOLVER_CHOLESKY)
        for x, y in ((a.user, b.user), (a.item, b.item)):
            assert np.abs(x - y).max() <= tol * np.abs(y).max()


def test_ialspp_block_sweeps_converge_to_the_exact_solve():
    """Block coordinate descent on a strictly convex quadratic: more sweeps, closer to A^-1 b."""
    from irspack_b200.synth import synth_csr

    X = synth_csr(300, 200, 6000, seed=5, values="counts")
    t = oracle.OracleTrainer(X, 24, 0.1, 0.05, 1.0, oracle.LOSS_ORIGINAL, dtype=np.float64)
    t.step(oracle.SOLVER_CHOLESKY)  # factors of realistic size, so that the blocks are coupled
    P = oracle.gram(t.item, 0.1)
    exact = np.zeros_like(t.user)
    oracle.step_cholesky(exact, t.X, t.item, P, 0.1, 0.05, 1.0, oracle.LOSS_ORIGINAL)
    errs = []
    for sweeps in (1, 10, 100):
        x = np.zeros_like(t.user)
        oracle.step_ialspp(x, t.X, t.item, P, 0.1, 0.05, 1.0, oracle.LOSS_ORIGINAL, 5, sweeps, 2)
        errs.append(np.abs(x - exact).max())
    assert errs[0] > errs[1] > errs[2] and errs[2] < 1e-2 * np.abs(exact).max()
    with pytest.raises(ValueError):
        oracle.step_ialspp(x, t.X, t.item, P, 0.1, 0.05, 1.0, oracle.LOSS_ORIGINAL, 0, 1, 1)


def _numpy_step_icd(target, X, other, P, alpha0, reg, nu, bias, iterations):
    """Solver::step_icd / _step_icd restated in numpy float64 (IALSTrainer.hpp:537-632):
    the path Solver::step takes for solver_type IALSPP with ialspp_subspace_dimension == 1
    (:671-676).  Dimension-outer, row-inner, prediction cache per stored entry."""
    x = target.copy()
    for _ in range(iterations):
        rows = np.repeat(np.arange(X.shape[0]), np.diff(X.indptr))
        pred = np.einsum("ij,ij->i", x[rows], other[X.indices])  # _prediction, :387-420
        for d in range(x.shape[1]):
            col = x[:, d].copy()
            for u in range(X.shape[0]):
                s, e = X.indptr[u], X.indptr[u + 1]
                reg_u = reg * (alpha0 * other.shape[0] + (e - s)) ** nu
                v = other[X.indices[s:e], d]
                c = X.data[s:e]
                B = P[d] @ x[u] + reg_u * col[u] + ((c * (pred[s:e] - 1) - bias) * v).sum()
                A = P[d, d] + (c * v * v).sum() + reg_u
                delta = B / A
                col[u] -= delta
                pred[s:e] -= delta * v
            x[:, d] = col  # :631: written back after every row of the dimension is done
    return x


@pytest.mark.parametrize("loss", [oracle.LOSS_ORIGINAL, oracle.LOSS_IALSPP])
def test_subspace_dimension_one_is_the_icd_solver(loss):
    """The reference dispatches IALSPP with a one-dimensional subspace to step_icd.  Rows are
    independent inside a half-step, so that loop is the block solver with S = 1: the oracle
    (and the CUDA path, tests/test_gpu_parity.py::test_ref_overfit_ialspp[1]) serve it from
    the block code."""
    from irspack_b200.synth import synth_csr

    X = synth_csr(60, 40, 500, seed=9, values="counts")
    t = oracle.OracleTrainer(X, 7, 0.1, 0.05, 1.0, loss, dtype=np.float64)
    P = oracle.gram(t.item, 0.1)
    bias = 0.1 if loss == oracle.LOSS_ORIGINAL else 0.0
    for iterations in (1, 3):
        want = _numpy_step_icd(t.user, t.X, t.item, P, 0.1, 0.05, 1.0, bias, iterations)
        got = t.user.copy()
        oracle.step_ialspp(got, t.X, t.item, P, 0.1, 0.05, 1.0, loss, 1, iterations, 2)
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("loss_type,alpha0", [("ORIGINAL", 0.1), ("IALSPP", 0.0), ("IALSPP", 0.1)])
def test_loss_identity(X_small, loss_type, alpha0):
    inv.loss_identity(OracleBackend, X_small, loss_type, alpha0)
    inv.loss_identity(f64, X_small, loss_type, alpha0)


def test_user_scores_batching():
    inv.user_scores_batching(OracleBackend)


@pytest.mark.parametrize("Backend", [OracleBackend, f64], ids=["f32", "f64"])
def test_cg_matches_cholesky(Backend, X_small):
    inv.cg_matches_cholesky(Backend, X_small)


def test_stationary_point_logscale(X_small):
    inv.stationary_point_logscale(OracleBackend, X_small)
    inv.stationary_point_logscale(f64, X_small, atol=1e-9)


def test_threads_agree():
    """n_threads only changes the Gram partial-sum order (IALSTrainer.hpp:95-112)."""
    rng = np.random.default_rng(3)
    X = sps.random(300, 200, density=0.05, random_state=3, format="csr", dtype=np.float32)
    X.data[:] = 1.0
    res = []
    for nt in (1, 4):
        t = oracle.OracleTrainer(X, 16, alpha0=0.1, reg=0.05, seed=7)
        for _ in range(3):
            t.step(oracle.SOLVER_CG, 3, n_threads=nt)
        res.append((t.user.copy(), t.item.copy()))
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-4, atol=1e-6)


def test_epoch_native_equals_python_orchestration():
    X = sps.random(120, 90, density=0.1, random_state=1, format="csr", dtype=np.float32)
    a = oracle.OracleTrainer(X, 8, seed=5)
    b = oracle.OracleTrainer(X, 8, seed=5)
    for solver in (oracle.SOLVER_CG, oracle.SOLVER_CHOLESKY):
        a.step(solver, 3)
        b.epoch_native(solver, 3)
        np.testing.assert_array_equal(a.user, b.user)
        np.testing.assert_array_equal(a.item, b.item)


def test_cg_singular_raises():
    """!(p.Ap > 0) -> runtime_error (IALSTrainer.hpp:249-254): negative confidences."""
    X = sps.csr_matrix(np.array([[-50.0, -50.0], [1.0, 0.0]], dtype=np.float32))
    t = oracle.OracleTrainer(X, 2, alpha0=0.0, reg=1e-3, nu=0.0, seed=0)
    with pytest.raises(RuntimeError, match="singular"):
        t.step(oracle.SOLVER_CG, 3)


def test_cholesky_failure_raises():
    """Non-positive pivot -> 'Cholesky decomposition failed.' (IALSTrainer.hpp:317-319)."""
    X = sps.csr_matrix(np.array([[1.0, 1.0], [1.0, 0.0]], dtype=np.float32))
    t = oracle.OracleTrainer(X, 2, alpha0=0.0, reg=-10.0, nu=0.0, seed=0)
    with pytest.raises(RuntimeError, match="Cholesky"):
        t.step(oracle.SOLVER_CHOLESKY, 3)


# ---- evaluator: restated from tests/evaluation/test_evaluator.py ----

@pytest.mark.parametrize("U,I,dtype", [(10, 5, "float32"), (10, 30, "float64"), (300, 5, "float32")])
def test_metrics_vs_sklearn(U, I, dtype):  # test_evaluator.py:19-49
    from sklearn.metrics import average_precision_score, ndcg_score

    rns = np.random.RandomState(42)
    scores = rns.randn(U, I).astype(dtype)
    X_gt = (rns.rand(U, I) >= 0.7).astype(np.float64)
    m, _, _ = oracle.topk_metrics(scores, sps.csr_matrix(X_gt), cutoff=I)
    d = m.as_dict()
    maps, ndcgs = [], []
    for i in range(U):
        if X_gt[i].sum() == 0:
            continue
        maps.append(average_precision_score(X_gt[i], scores[i]))
        ndcgs.append(ndcg_score(X_gt[i][None, :], scores[i][None, :]))
    assert d["map"] == pytest.approx(np.mean(maps), abs=1e-8)
    assert d["ndcg"] == pytest.approx(np.mean(ndcgs), abs=1e-8)


@pytest.mark.parametrize("U,I,C", [(10, 5, 5), (10, 30, 29)])
def test_metrics_with_cutoff(U, I, C):  # test_evaluator.py:87-152
    from sklearn.metrics import ndcg_score

    rns = np.random.RandomState(42)
    scores = rns.randn(U, I)
    X_gt = (rns.rand(U, I) >= 0.3).astype(np.float64)
    gt = sps.csr_matrix(X_gt)
    empty = sps.csr_matrix(X_gt.shape)
    d, _ = oracle.evaluate(lambda b, e: scores[b:e].copy(), empty, gt, cutoff=C)
    d1, _ = oracle.evaluate(lambda b, e: scores[b:e].copy(), empty, gt, cutoff=C, mb_size=1)
    for k in d:
        assert d1[k] == pytest.approx(d[k])
    ndcg = valid = map_ = prec = rec = 0.0
    cnt = np.zeros(I)
    for i in range(U):
        nzs = set(X_gt[i].nonzero()[0])
        if not nzs:
            continue
        valid += 1
        ndcg += ndcg_score(X_gt[[i]], scores[[i]], k=C)
        top = scores[i].argsort()[::-1][:C]
        denom = min(C, len(nzs))
        ap = hit = 0
        for r, it in enumerate(top):
            cnt[it] += 1
            if it in nzs:
                hit += 1
                ap += hit / float(r + 1)
        map_ += ap / denom
        rec += hit / denom
        prec += hit / C
    p = cnt / cnt.sum()
    entropy = -p.dot(np.log(p))
    lorentz = np.cumsum(np.sort(cnt) / cnt.sum())
    gini = sum((1 / I) * 2 * (((i + 1) / I) - lorentz[i]) for i in range(I))
    assert d["ndcg"] == pytest.approx(ndcg / valid)
    assert d["precision"] == pytest.approx(prec / valid, abs=1e-8)
    assert d["entropy"] == pytest.approx(entropy)
    assert d["gini_index"] == pytest.approx(gini)
    # the reference's own "map"/"recall" use n_gt unless recall_with_cutoff; here C >= most n_gt
    dc, _ = oracle.evaluate(lambda b, e: scores[b:e].copy(), empty, gt, cutoff=C,
                            recall_with_cutoff=True)
    assert dc["recall"] == pytest.approx(rec / valid, abs=1e-8)


def test_minus_inf_never_recommended():  # test_evaluator.py:358-368
    scores = np.array([[0.5, -np.inf, 0.1, -np.inf]], dtype=np.float32)
    gt = sps.csr_matrix(np.array([[1.0, 1.0, 0.0, 0.0]]))
    _, rec, cnt = oracle.topk_metrics(scores, gt, cutoff=4)
    assert cnt[0] == 2 and list(rec[0]) == [0, 2, -1, -1]


def test_tie_break_by_smaller_index():  # evaluator.cpp:329, 353-355
    scores = np.array([[1.0, 2.0, 2.0, 1.0, 2.0]], dtype=np.float32)
    gt = sps.csr_matrix(np.ones((1, 5)))
    _, rec, _ = oracle.topk_metrics(scores, gt, cutoff=4)
    assert list(rec[0]) == [1, 2, 4, 0]
