"""Feature-aware iALS in the ORACLE (the row SURVEY.md 8 f4 names after iALS++; the CUDA side
still raises NotImplementedError).  Restates the reference's own closed-form test,
/root/reference/tests/recommenders/test_ials.py:79-245
(test_feature_aware_ials_weighted_updates_objective_and_local_stability), against
oracle.OracleTrainer: step_with_prior for CG and Cholesky (IALSTrainer.hpp:170-271 with a prior,
:333-385, :634-662), the feature-weight ridge (:1095-1209), the feature-aware epoch (:758-783)
and the loss (:836-940)."""
import numpy as np
import pytest
import scipy.sparse as sps

import oracle

INTERACTION = np.array([[1, 0, 2, 1], [0, 3, 0, 0], [1, 1, 0, 4]], dtype=np.float64)
USER_F = np.array([[1, 0.2], [0.3, 1], [0.7, -0.2]], dtype=np.float32)
ITEM_F = np.array([[1, 0, 0.1], [0, 1, 0.2], [0.5, 0.2, 1], [-0.2, 0.8, 0.4]], dtype=np.float32)
ALPHA0, REG, NU, LAM_U, LAM_I = 0.7, 0.03, 0.6, 0.11, 0.17


def _train(solver, max_cg_steps, feature_type, dtype, epochs=500):
    uf, itf = (USER_F, ITEM_F) if feature_type == "dense" else (sps.csr_matrix(USER_F), sps.csr_matrix(ITEM_F))
    t = oracle.OracleTrainer(sps.csr_matrix(INTERACTION.astype(np.float32)), 3, ALPHA0, REG, NU,
                             oracle.LOSS_ORIGINAL, dtype=dtype, seed=0, user_features=uf, item_features=itf,
                             lambda_user_feature=LAM_U, lambda_item_feature=LAM_I)
    for _ in range(epochs):
        t.step(solver, max_cg_steps, n_threads=2)
    return t


@pytest.mark.parametrize("solver,max_cg_steps", [(oracle.SOLVER_CHOLESKY, 3), (oracle.SOLVER_CG, 0)])
@pytest.mark.parametrize("feature_type", ["dense", "sparse"])
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-9), (np.float32, 2e-5)])
def test_weighted_updates_objective_and_local_stability(solver, max_cg_steps, feature_type, dtype, tol):
    t = _train(solver, max_cg_steps, feature_type, dtype)
    user, item = t.user.astype(np.float64), t.item.astype(np.float64)
    uw, iw = t.user_feature_weight.astype(np.float64), t.item_feature_weight.astype(np.float64)
    uf, itf = USER_F.astype(np.float64), ITEM_F.astype(np.float64)
    user_reg = REG * (ALPHA0 * INTERACTION.shape[1] + np.count_nonzero(INTERACTION, axis=1)) ** NU
    item_reg = REG * (ALPHA0 * INTERACTION.shape[0] + np.count_nonzero(INTERACTION, axis=0)) ** NU

    # the weights are the weighted ridge fit of the factors (test_ials.py:137-148)
    want_uw = np.linalg.solve(uf.T @ (user_reg[:, None] * uf) + LAM_U * np.eye(2), uf.T @ (user_reg[:, None] * user))
    want_iw = np.linalg.solve(itf.T @ (item_reg[:, None] * itf) + LAM_I * np.eye(3), itf.T @ (item_reg[:, None] * item))
    np.testing.assert_allclose(uw, want_uw, rtol=max(tol, 2e-6), atol=max(tol, 2e-6))
    np.testing.assert_allclose(iw, want_iw, rtol=max(tol, 2e-6), atol=max(tol, 2e-6))

    def objective(values):  # test_ials.py:150-175
        u, i, wu, wi = values
        score = u @ i.T
        observed = INTERACTION.astype(bool)
        loss = ALPHA0 * np.square(score[~observed]).sum()
        loss += np.sum((INTERACTION[observed] + ALPHA0) * np.square(score[observed] - 1))
        loss += np.sum(user_reg[:, None] * np.square(u - uf @ wu)) + np.sum(item_reg[:, None] * np.square(i - itf @ wi))
        loss += LAM_U * np.square(wu).sum() + LAM_I * np.square(wi).sum()
        return float(loss / 2)

    params = [user, item, uw, iw]
    optimum = objective(params)
    np.testing.assert_allclose(t.compute_loss(), optimum, rtol=max(tol, 2e-6), atol=max(tol, 2e-6))

    def solve_embeddings(histories, other, prior, regs):  # test_ials.py:184-201
        out = []
        base = ALPHA0 * other.T @ other
        for row, prior_row, row_reg in zip(histories, prior, regs):
            lhs = base + row_reg * np.eye(other.shape[1])
            rhs = row_reg * prior_row
            for j, value in enumerate(row):
                if value:
                    lhs = lhs + value * np.outer(other[j], other[j])
                    rhs = rhs + (ALPHA0 + value) * other[j]
            out.append(np.linalg.solve(lhs, rhs))
        return np.asarray(out)

    X = sps.csr_matrix(INTERACTION.astype(np.float32))
    got_u = t.transform_user_with_feature(X, USER_F, solver, 0)  # prediction_time_max_cg_steps = 0 -> K steps
    got_i = t.transform_item_with_feature(X, ITEM_F, solver, 0)
    np.testing.assert_allclose(got_u, solve_embeddings(INTERACTION, item, uf @ uw, user_reg), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(got_i, solve_embeddings(INTERACTION.T, user, itf @ iw, item_reg), rtol=2e-5, atol=2e-5)

    if dtype == np.float64:  # a joint stationary point: no direction lowers the objective (:221-245)
        rng = np.random.default_rng(1)
        for radius in (1e-5, 1e-3):
            for _ in range(64):
                direction = [rng.standard_normal(v.shape) for v in params]
                norm = np.sqrt(sum(np.square(v).sum() for v in direction))
                for sign in (-1, 1):
                    moved = [v + sign * radius * d / norm for v, d in zip(params, direction)]
                    assert objective(moved) >= optimum - 5e-10


def test_without_features_the_trainer_is_unchanged():
    X = sps.csr_matrix(INTERACTION.astype(np.float32))
    a = oracle.OracleTrainer(X, 3, ALPHA0, REG, NU, oracle.LOSS_ORIGINAL, seed=0)
    b = oracle.OracleTrainer(X, 3, ALPHA0, REG, NU, oracle.LOSS_ORIGINAL, seed=0,
                             user_features=np.zeros((3, 0), np.float32), item_features=np.zeros((4, 0), np.float32))
    for _ in range(3):
        a.step(oracle.SOLVER_CG, 3)
        b.step(oracle.SOLVER_CG, 3)
    np.testing.assert_array_equal(a.user, b.user)
    np.testing.assert_array_equal(a.item, b.item)
    assert b.compute_loss() == pytest.approx(a.compute_loss(), rel=1e-6)


def test_warmup_epochs_and_argument_errors():
    X = sps.csr_matrix(INTERACTION.astype(np.float32))
    kw = dict(user_features=USER_F, item_features=ITEM_F, lambda_user_feature=LAM_U, lambda_item_feature=LAM_I)
    t = oracle.OracleTrainer(X, 3, ALPHA0, REG, NU, oracle.LOSS_ORIGINAL, seed=0, feature_warmup_epochs=2, **kw)
    plain = oracle.OracleTrainer(X, 3, ALPHA0, REG, NU, oracle.LOSS_ORIGINAL, seed=0)
    for _ in range(2):  # IALSTrainer.hpp:762: plain epochs until epoch_ reaches the warm-up
        t.step(oracle.SOLVER_CHOLESKY)
        plain.step(oracle.SOLVER_CHOLESKY)
    np.testing.assert_array_equal(t.user, plain.user)
    assert not t.user_feature_weight.any()
    t.step(oracle.SOLVER_CHOLESKY)
    assert t.user_feature_weight.any() and t.item_feature_weight.any()
    with pytest.raises(ValueError, match="IALSPP"):  # :759-761
        t.step(oracle.SOLVER_IALSPP)
    with pytest.raises(ValueError, match="row count"):  # :1006-1007
        oracle.OracleTrainer(X, 3, user_features=USER_F[:2], item_features=ITEM_F, lambda_user_feature=1.0,
                             lambda_item_feature=1.0)
    with pytest.raises(ValueError, match="must be positive"):  # :1008-1011
        oracle.OracleTrainer(X, 3, user_features=USER_F, item_features=ITEM_F, lambda_user_feature=0.0,
                             lambda_item_feature=1.0)
    with pytest.raises(ValueError, match="Shape mismatch"):  # :1016-1028
        t.transform_user_feature(np.zeros((2, 5), np.float32))
    # alpha0 = 0 and a vanishing regulariser leave an empty row's embedding undefined (:640-654)
    Xe = sps.csr_matrix(np.array([[1, 0], [0, 0]], dtype=np.float32))
    te = oracle.OracleTrainer(Xe, 2, 0.0, 0.0, 1.0, oracle.LOSS_IALSPP, seed=0, user_features=np.ones((2, 1), np.float32),
                              item_features=np.ones((2, 1), np.float32), lambda_user_feature=1.0, lambda_item_feature=1.0)
    with pytest.raises(ValueError, match="not uniquely defined"):
        te.step(oracle.SOLVER_CHOLESKY)


# ---- pinned to the reference's own code (oracle/_ref: IALSTrainer.hpp compiled where it lies) ----

def _random_problem(seed, sparse):
    rng = np.random.default_rng(seed)
    U, I, Fu, Fi = 60, 45, 4, 3
    X = sps.random(U, I, density=0.12, random_state=seed, format="csr", dtype=np.float32)
    X.data[:] = rng.choice([0.5, 1.0, 2.0], size=X.nnz).astype(np.float32)
    X = sps.csr_matrix(X.toarray() * (rng.random((U, 1)) > 0.1))  # some empty rows
    uf = (rng.standard_normal((U, Fu)) * (rng.random((U, Fu)) < 0.7)).astype(np.float32)
    itf = (rng.standard_normal((I, Fi)) * (rng.random((I, Fi)) < 0.7)).astype(np.float32)
    if sparse:
        uf, itf = sps.csr_matrix(uf), sps.csr_matrix(itf)
    return X, uf, itf


@pytest.mark.skipif(not oracle.RefTrainer.available(), reason="oracle/_ref is not built (needs /root/reference)")
@pytest.mark.parametrize("solver,steps", [(oracle.SOLVER_CG, 3), (oracle.SOLVER_CHOLESKY, 3)])
@pytest.mark.parametrize("sparse", [False, True])
@pytest.mark.parametrize("loss", [oracle.LOSS_IALSPP, oracle.LOSS_ORIGINAL])
def test_feature_aware_oracle_matches_the_references_own_trainer(solver, steps, sparse, loss):
    """Four epochs (one warm-up) of the feature-aware model: the oracle's restatement against the
    reference's IALSTrainer.hpp itself -- factors, both feature weights, loss, fold-in with features."""
    X, uf, itf = _random_problem(5 + sparse, sparse)
    K = 6
    kw = dict(user_features=uf, item_features=itf, lambda_user_feature=0.3, lambda_item_feature=0.2,
              feature_warmup_epochs=1)
    r = oracle.RefTrainer(X, K, 0.2, 0.05, 0.8, loss, random_seed=3, **kw)
    o = oracle.OracleTrainer(X, K, 0.2, 0.05, 0.8, loss, dtype=np.float32, **kw)
    o.user, o.item = r.user, r.item  # the reference's own initialisation
    ref_solver = {oracle.SOLVER_CG: 1, oracle.SOLVER_CHOLESKY: 0}[solver]
    for _ in range(4):
        r.step(ref_solver, steps)
        o.step(solver, steps)
    for a, b, what in ((o.user, r.user, "user"), (o.item, r.item, "item"),
                       (o.user_feature_weight, r.user_feature_weight, "user weight"),
                       (o.item_feature_weight, r.item_feature_weight, "item weight")):
        assert a.shape == b.shape, what
        assert np.abs(a - b).max() <= 5e-5 * (np.abs(b).max() + 1e-30), (what, np.abs(a - b).max())
    assert o.compute_loss() == pytest.approx(r.compute_loss(), rel=2e-5)
    Xn, ufn, _ = _random_problem(11, False)
    got = o.transform_user_with_feature(Xn[:20], ufn[:20], solver, 5)
    want = r.transform_with_feature(0, Xn[:20], ufn[:20], ref_solver, 5)
    assert np.abs(got - want).max() <= 5e-5 * (np.abs(want).max() + 1e-30)
