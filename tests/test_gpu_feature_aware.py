"""Feature-aware iALS on the CUDA path (SURVEY.md 8 f4 "after that"): the reference's own
closed-form test (/root/reference/tests/recommenders/test_ials.py:79-245) and its argument
errors restated on irspack_b200._ials_core.IALSTrainer, plus epoch-by-epoch agreement with the
oracle's feature-aware trainer (oracle.OracleTrainer; IALSTrainer.hpp:170-271 with a prior,
:333-385, :634-662, :758-789, :836-940, :1066-1209) on a larger random problem."""
import pickle

import numpy as np
import pytest
import scipy.sparse as sps

import oracle

pytestmark = pytest.mark.gpu

INTERACTION = np.array([[1, 0, 2, 1], [0, 3, 0, 0], [1, 1, 0, 4]], dtype=np.float64)
USER_F = np.array([[1, 0.2], [0.3, 1], [0.7, -0.2]], dtype=np.float32)
ITEM_F = np.array([[1, 0, 0.1], [0, 1, 0.2], [0.5, 0.2, 1], [-0.2, 0.8, 0.4]], dtype=np.float32)
ALPHA0, REG, NU, LAM_U, LAM_I = 0.7, 0.03, 0.6, 0.11, 0.17


@pytest.fixture(scope="module")
def core():
    from irspack_b200 import _ials_core as core

    return core


def _config(core, K, alpha0=ALPHA0, reg=REG, nu=NU, loss="ORIGINAL", warmup=0, lam_u=LAM_U, lam_i=LAM_I, seed=0):
    return (core.IALSModelConfigBuilder().set_K(K).set_alpha0(alpha0).set_reg(reg).set_nu(nu)
            .set_loss_type(getattr(core.LossType, loss)).set_random_seed(seed).set_lambda_user_feature(lam_u)
            .set_lambda_item_feature(lam_i).set_feature_warmup_epochs(warmup).build())


def _solver(core, solver, steps=3):
    return (core.IALSSolverConfigBuilder().set_solver_type(getattr(core.SolverType, solver))
            .set_max_cg_steps(steps).build())


@pytest.mark.parametrize("solver,max_cg_steps", [("CHOLESKY", 3), ("CG", 0)])
@pytest.mark.parametrize("feature_type", ["dense", "sparse"])
def test_weighted_updates_objective_and_fold_in(core, solver, max_cg_steps, feature_type):
    """test_ials.py:79-219 on the GPU: after many epochs the weights are the weighted ridge fit of the
    factors, compute_loss is the feature-aware objective, the fold-in with features is the
    closed-form embedding."""
    uf, itf = (USER_F, ITEM_F) if feature_type == "dense" else (sps.csr_matrix(USER_F), sps.csr_matrix(ITEM_F))
    X = sps.csr_matrix(INTERACTION.astype(np.float32))
    t = core.IALSTrainer(_config(core, 3), X, uf, itf)
    sc = _solver(core, solver, max_cg_steps)
    for _ in range(300):
        t.step(sc)
    user, item = t.user.astype(np.float64), t.item.astype(np.float64)
    uw, iw = t.user_feature_weight.astype(np.float64), t.item_feature_weight.astype(np.float64)
    assert uw.shape == (2, 3) and iw.shape == (3, 3)
    ufd, itfd = USER_F.astype(np.float64), ITEM_F.astype(np.float64)
    user_reg = REG * (ALPHA0 * INTERACTION.shape[1] + np.count_nonzero(INTERACTION, axis=1)) ** NU
    item_reg = REG * (ALPHA0 * INTERACTION.shape[0] + np.count_nonzero(INTERACTION, axis=0)) ** NU
    want_uw = np.linalg.solve(ufd.T @ (user_reg[:, None] * ufd) + LAM_U * np.eye(2), ufd.T @ (user_reg[:, None] * user))
    want_iw = np.linalg.solve(itfd.T @ (item_reg[:, None] * itfd) + LAM_I * np.eye(3), itfd.T @ (item_reg[:, None] * item))
    np.testing.assert_allclose(uw, want_uw, rtol=5e-5, atol=5e-5)
    np.testing.assert_allclose(iw, want_iw, rtol=5e-5, atol=5e-5)

    score = user @ item.T
    observed = INTERACTION.astype(bool)
    loss = ALPHA0 * np.square(score[~observed]).sum()
    loss += np.sum((INTERACTION[observed] + ALPHA0) * np.square(score[observed] - 1))
    loss += np.sum(user_reg[:, None] * np.square(user - ufd @ uw)) + np.sum(item_reg[:, None] * np.square(item - itfd @ iw))
    loss += LAM_U * np.square(uw).sum() + LAM_I * np.square(iw).sum()
    assert t.compute_loss(sc) == pytest.approx(loss / 2, rel=5e-5, abs=5e-5)

    def solve_embeddings(histories, other, prior, regs):
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

    fold = _solver(core, solver, 0)  # prediction_time_max_cg_steps = 0 -> K steps
    got_u = t.transform_user_with_feature(X, uf, fold)
    got_i = t.transform_item_with_feature(X, itf, fold)
    np.testing.assert_allclose(got_u, solve_embeddings(INTERACTION, item, ufd @ uw, user_reg), rtol=5e-5, atol=5e-5)
    np.testing.assert_allclose(got_i, solve_embeddings(INTERACTION.T, user, itfd @ iw, item_reg), rtol=5e-5, atol=5e-5)
    np.testing.assert_allclose(t.transform_user_feature(uf), ufd @ uw, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(t.transform_item_feature(itf), itfd @ iw, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("solver,K", [("CG", 32), ("CHOLESKY", 32), ("CG", 160), ("CHOLESKY", 136)])
@pytest.mark.parametrize("sparse", [False, True])
def test_epochs_match_the_oracle(core, solver, K, sparse):
    """Three feature-aware epochs (one of them a warm-up) on a random problem with empty rows and
    columns, both feature kinds: factors, weights and loss against the oracle's float32 trainer."""
    rng = np.random.default_rng(K + sparse)
    U, I, Fu, Fi = 300, 200, 7, 5
    X = sps.random(U, I, density=0.05, random_state=3, format="csr", dtype=np.float32)
    X.data[:] = rng.choice([0.5, 1.0, 2.0, 3.0], size=X.nnz).astype(np.float32)
    X = sps.csr_matrix(X.toarray() * (rng.random((U, 1)) > 0.05))  # some users without interactions
    uf = rng.standard_normal((U, Fu)).astype(np.float32) * (rng.random((U, Fu)) < 0.6)
    itf = rng.standard_normal((I, Fi)).astype(np.float32) * (rng.random((I, Fi)) < 0.6)
    ufm, itfm = (sps.csr_matrix(uf), sps.csr_matrix(itf)) if sparse else (uf, itf)
    from irspack_b200.synth import init_factors

    u0, i0 = init_factors(U, K, 1), init_factors(I, K, 2)
    g = core.IALSTrainer(_config(core, K, alpha0=0.1, reg=0.05, nu=1.0, loss="IALSPP", warmup=1, lam_u=0.3, lam_i=0.2),
                         X, ufm, itfm)
    g.user, g.item = u0, i0
    o = oracle.OracleTrainer(X, K, 0.1, 0.05, 1.0, oracle.LOSS_IALSPP, dtype=np.float32, user_features=ufm,
                             item_features=itfm, lambda_user_feature=0.3, lambda_item_feature=0.2,
                             feature_warmup_epochs=1)
    o.user, o.item = u0.copy(), i0.copy()
    sc = _solver(core, solver)
    osolver = oracle.SOLVER_CG if solver == "CG" else oracle.SOLVER_CHOLESKY
    for epoch in range(3):
        g.step(sc)
        o.step(osolver, 3)
        if epoch == 0:
            assert not g.user_feature_weight.any()  # still in the warm-up (IALSTrainer.hpp:762)
    for a, b, what in ((g.user, o.user, "user"), (g.item, o.item, "item"),
                       (g.user_feature_weight, o.user_feature_weight, "user weight"),
                       (g.item_feature_weight, o.item_feature_weight, "item weight")):
        err = np.abs(a - b).max() / (np.abs(b).max() + 1e-30)
        assert err <= 3e-4, (what, err)
    assert g.compute_loss(sc) == pytest.approx(o.compute_loss(), rel=2e-4)


def test_without_feature_columns_the_trainer_is_unchanged(core):
    X = sps.csr_matrix(INTERACTION.astype(np.float32))
    a = core.IALSTrainer(_config(core, 3), X)
    b = core.IALSTrainer(_config(core, 3), X, np.zeros((3, 0), np.float32), sps.csr_matrix((4, 0), dtype=np.float32))
    sc = _solver(core, "CG")
    for _ in range(3):
        a.step(sc)
        b.step(sc)
    np.testing.assert_array_equal(a.user, b.user)
    np.testing.assert_array_equal(a.item, b.item)
    assert b.user_feature_weight.shape == (0, 3) and a.user_feature_weight.shape == (0, 3)
    assert b.compute_loss(sc) == pytest.approx(a.compute_loss(sc), rel=1e-6)


def test_warmup_errors_and_pickle(core):
    X = sps.csr_matrix(INTERACTION.astype(np.float32))
    t = core.IALSTrainer(_config(core, 3, warmup=2), X, USER_F, ITEM_F)
    plain = core.IALSTrainer(_config(core, 3), X)
    chol = _solver(core, "CHOLESKY")
    for _ in range(2):  # plain epochs until epoch_ reaches the warm-up (IALSTrainer.hpp:762)
        t.step(chol)
        plain.step(chol)
    np.testing.assert_array_equal(t.user, plain.user)
    assert not t.user_feature_weight.any()
    t.step(chol)
    assert t.user_feature_weight.any() and t.item_feature_weight.any()
    with pytest.raises(ValueError, match="IALSPP"):  # :759-761
        t.step(_solver(core, "IALSPP"))
    with pytest.raises(ValueError, match="row count"):  # :1006-1007
        core.IALSTrainer(_config(core, 3, lam_u=1.0, lam_i=1.0), X, USER_F[:2], ITEM_F)
    with pytest.raises(ValueError, match="must be positive"):  # :1008-1011
        core.IALSTrainer(_config(core, 3, lam_u=0.0, lam_i=1.0), X, USER_F, ITEM_F)
    with pytest.raises(ValueError, match="Shape mismatch"):  # :1016-1028
        t.transform_user_feature(np.zeros((2, 5), np.float32))
    with pytest.raises(ValueError, match="not initialized"):  # a trainer without features
        plain.transform_user_feature(USER_F)
    with pytest.raises(TypeError):  # the overload takes both matrices (wrapper.cpp:133-136)
        core.IALSTrainer(_config(core, 3), X, user_feature=USER_F)
    # alpha0 = 0 and a vanishing regulariser leave an empty row's embedding undefined (:640-654)
    Xe = sps.csr_matrix(np.array([[1, 0], [0, 0]], dtype=np.float32))
    te = core.IALSTrainer(_config(core, 2, alpha0=0.0, reg=0.0, nu=1.0, loss="IALSPP", lam_u=1.0, lam_i=1.0), Xe,
                          np.ones((2, 1), np.float32), np.ones((2, 1), np.float32))
    with pytest.raises(ValueError, match="not uniquely defined"):
        te.step(chol)
    # the pickle tuple carries the weights (wrapper.cpp:162-181)
    t2 = pickle.loads(pickle.dumps(t))
    np.testing.assert_array_equal(t2.user_feature_weight, t.user_feature_weight)
    np.testing.assert_array_equal(t2.item_feature_weight, t.item_feature_weight)
    np.testing.assert_allclose(t2.transform_item_feature(ITEM_F), t.transform_item_feature(ITEM_F), rtol=1e-6)


def test_recommender_with_features(core):
    """IALSRecommender(X, user_features=..., item_features=...) (ials.py:383-436, 538-621)."""
    import irspack_b200

    rng = np.random.default_rng(0)
    X = sps.random(120, 80, density=0.1, random_state=1, format="csr", dtype=np.float32)
    X.data[:] = 1.0
    uf = rng.standard_normal((120, 4)).astype(np.float32)
    rec = irspack_b200.IALSRecommender(X, n_components=8, alpha0=0.1, reg=0.05, solver_type="CHOLESKY",
                                       train_epochs=4, user_features=uf, lambda_user_feature=0.5).learn()
    core_t = rec.trainer_as_ials.core_trainer
    assert core_t.user_feature_weight.shape == (4, 8) and core_t.item_feature_weight.shape == (0, 8)
    cold = rec.compute_user_embedding_from_features(uf[:5])
    assert cold.shape == (5, 8) and np.isfinite(cold).all() and np.abs(cold).max() > 0
    warm = rec.compute_user_embedding(X[:5], user_features=uf[:5])
    assert warm.shape == (5, 8) and not np.allclose(warm, cold)
    with pytest.raises(ValueError, match="IALSPP"):
        irspack_b200.IALSRecommender(X, solver_type="IALSPP", user_features=uf)


@pytest.mark.skipif(not oracle.RefTrainer.available(), reason="oracle/_ref is not built")
@pytest.mark.parametrize("solver", ["CG", "CHOLESKY"])
@pytest.mark.parametrize("sparse", [False, True])
def test_epochs_match_the_references_own_trainer(core, solver, sparse):
    """The CUDA path against the reference's IALSTrainer.hpp itself (oracle/_ref): four feature-aware
    epochs from the reference's initial factors, weights, loss and the fold-in with features."""
    rng = np.random.default_rng(17 + sparse)
    U, I, Fu, Fi, K = 90, 70, 5, 4, 24
    X = sps.random(U, I, density=0.1, random_state=4, format="csr", dtype=np.float32)
    X.data[:] = rng.choice([0.5, 1.0, 2.0], size=X.nnz).astype(np.float32)
    X = sps.csr_matrix(X.toarray() * (rng.random((U, 1)) > 0.1))
    uf = (rng.standard_normal((U, Fu)) * (rng.random((U, Fu)) < 0.7)).astype(np.float32)
    itf = (rng.standard_normal((I, Fi)) * (rng.random((I, Fi)) < 0.7)).astype(np.float32)
    ufm, itfm = (sps.csr_matrix(uf), sps.csr_matrix(itf)) if sparse else (uf, itf)
    r = oracle.RefTrainer(X, K, 0.2, 0.05, 0.8, oracle.LOSS_IALSPP, random_seed=3, user_features=ufm,
                          item_features=itfm, lambda_user_feature=0.3, lambda_item_feature=0.2,
                          feature_warmup_epochs=1)
    g = core.IALSTrainer(_config(core, K, alpha0=0.2, reg=0.05, nu=0.8, loss="IALSPP", warmup=1, lam_u=0.3, lam_i=0.2),
                         X, ufm, itfm)
    g.user, g.item = r.user, r.item
    sc = _solver(core, solver)
    for _ in range(4):
        r.step(0 if solver == "CHOLESKY" else 1, 3)
        g.step(sc)
    for a, b, what in ((g.user, r.user, "user"), (g.item, r.item, "item"),
                       (g.user_feature_weight, r.user_feature_weight, "user weight"),
                       (g.item_feature_weight, r.item_feature_weight, "item weight")):
        err = np.abs(a - b).max() / (np.abs(b).max() + 1e-30)
        assert err <= 2e-4, (what, err)
    assert g.compute_loss(sc) == pytest.approx(r.compute_loss(), rel=1e-4)
    got = g.transform_user_with_feature(X[:20], uf[:20], _solver(core, solver, 5))
    want = r.transform_with_feature(0, X[:20], uf[:20], 0 if solver == "CHOLESKY" else 1, 5)
    assert np.abs(got - want).max() <= 2e-4 * (np.abs(want).max() + 1e-30)
