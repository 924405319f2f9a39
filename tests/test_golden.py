"""Golden vectors (tests/golden/ials_small.npz, made by tests/golden/make_golden.py: an
independent numpy-float64 restatement of the reference's formulas) against the C++ oracle
(CPU, not gpu-marked) and against the CUDA path (gpu-marked)."""
import os

import numpy as np
import pytest
import scipy.sparse as sps

import oracle

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ials_small.npz"))
U, I, K = (int(v) for v in G["shape"])
ALPHA0, REG, NU, CG_STEPS, EPOCHS = (float(v) for v in G["hyper"])
X = sps.csr_matrix((G["data"], G["indices"], G["indptr"]), shape=(U, I))
CASES = [("cg", "ialspp"), ("cg", "original"), ("chol", "ialspp"), ("chol", "original")]


def _oracle(dtype, solver, loss):
    lt = oracle.LOSS_ORIGINAL if loss == "original" else oracle.LOSS_IALSPP
    o = oracle.OracleTrainer(X, K, ALPHA0, REG, NU, lt, dtype=dtype)
    o.user, o.item = G["user0"].astype(dtype), G["item0"].astype(dtype)
    st = oracle.SOLVER_CG if solver == "cg" else oracle.SOLVER_CHOLESKY
    for _ in range(int(EPOCHS)):
        o.step(st, int(CG_STEPS), 1)
    return o


@pytest.mark.parametrize("solver,loss", CASES)
def test_oracle_f64_reproduces_golden(solver, loss):
    o = _oracle(np.float64, solver, loss)
    np.testing.assert_allclose(o.user, G[f"user_{solver}_{loss}"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(o.item, G[f"item_{solver}_{loss}"], rtol=1e-9, atol=1e-12)
    assert o.compute_loss() == pytest.approx(float(G[f"loss_{solver}_{loss}"]), rel=1e-9)


@pytest.mark.parametrize("solver,loss", CASES)
def test_oracle_f32_within_float32_tolerance_of_golden(solver, loss):
    o = _oracle(np.float32, solver, loss)
    for got, want in ((o.user, G[f"user_{solver}_{loss}"]), (o.item, G[f"item_{solver}_{loss}"])):
        assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max()
    assert o.compute_loss() == pytest.approx(float(G[f"loss_{solver}_{loss}"]), rel=1e-4)


def test_oracle_gram_scores_topk_golden():
    np.testing.assert_allclose(oracle.gram(G["item0"].copy(), ALPHA0, 1), G["gram_item0"], rtol=1e-12)
    o = _oracle(np.float64, "cg", "ialspp")
    np.testing.assert_allclose(o.user_scores(0, U), G["scores_cg_ialspp"], rtol=1e-9, atol=1e-12)
    gt = sps.csr_matrix((np.ones(U), (np.arange(U), np.zeros(U, int))), shape=(U, I))
    _, rec = oracle.evaluate(lambda b, e: G["user_cg_ialspp"][b:e] @ G["item_cg_ialspp"].T, X, gt, cutoff=10)
    np.testing.assert_array_equal(rec, G["top10_cg_ialspp"])


@pytest.mark.gpu
@pytest.mark.parametrize("solver,loss", CASES)
def test_cuda_path_within_float32_tolerance_of_golden(solver, loss):
    import irspack_b200
    from irspack_b200 import _ials_core as core

    if irspack_b200.device_count() == 0:
        pytest.fail("GPU test selected but no CUDA device is visible")
    lt = core.LossType.ORIGINAL if loss == "original" else core.LossType.IALSPP
    cfg = (core.IALSModelConfigBuilder().set_K(K).set_alpha0(ALPHA0).set_reg(REG).set_nu(NU)
           .set_loss_type(lt).build())
    st = core.SolverType.CG if solver == "cg" else core.SolverType.CHOLESKY
    sc = core.IALSSolverConfigBuilder().set_solver_type(st).set_max_cg_steps(int(CG_STEPS)).build()
    g = core.IALSTrainer(cfg, X.astype(np.float32))
    g.user, g.item = G["user0"].astype(np.float32), G["item0"].astype(np.float32)
    for _ in range(int(EPOCHS)):
        g.step(sc)
    for got, want in ((g.user, G[f"user_{solver}_{loss}"]), (g.item, G[f"item_{solver}_{loss}"])):
        assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max()
    assert g.compute_loss(sc) == pytest.approx(float(G[f"loss_{solver}_{loss}"]), rel=1e-4)
    if (solver, loss) == ("cg", "ialspp"):
        # top-10 on the GOLDEN factors: identical lists (no near-ties at this size: checked below)
        h = core.IALSTrainer(cfg, X.astype(np.float32))
        h.user, h.item = G["user_cg_ialspp"].astype(np.float32), G["item_cg_ialspp"].astype(np.float32)
        got, _ = h.recommend(0, U, 10, mask="train")
        want = G["top10_cg_ialspp"]
        s = G["scores_cg_ialspp"]
        for r in np.flatnonzero((got != want).any(axis=1)):
            for a, b in zip(got[r], want[r]):
                assert a == b or abs(s[r, a] - s[r, b]) <= 1e-5 * np.abs(s[r]).max()


# ---- iALS++: tests/golden/ialspp_small.npz, same inputs as above ----
GP = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ialspp_small.npz"))
PP_S, PP_IT = int(GP["subspace"]), int(GP["iterations"])


def _oracle_ialspp(dtype, loss):
    lt = oracle.LOSS_ORIGINAL if loss == "original" else oracle.LOSS_IALSPP
    o = oracle.OracleTrainer(X, K, ALPHA0, REG, NU, lt, dtype=dtype)
    o.user, o.item = G["user0"].astype(dtype), G["item0"].astype(dtype)
    o.ialspp_subspace_dimension, o.ialspp_iteration = PP_S, PP_IT
    for _ in range(int(EPOCHS)):
        o.step(oracle.SOLVER_IALSPP, int(CG_STEPS), 1)
    return o


@pytest.mark.parametrize("loss", ["ialspp", "original"])
def test_oracle_ialspp_reproduces_golden(loss):
    o = _oracle_ialspp(np.float64, loss)
    np.testing.assert_allclose(o.user, GP[f"user_{loss}"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(o.item, GP[f"item_{loss}"], rtol=1e-9, atol=1e-12)
    assert o.compute_loss() == pytest.approx(float(GP[f"loss_{loss}"]), rel=1e-9)
    o = _oracle_ialspp(np.float32, loss)
    for got, want in ((o.user, GP[f"user_{loss}"]), (o.item, GP[f"item_{loss}"])):
        assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max()


@pytest.mark.gpu
@pytest.mark.parametrize("loss", ["ialspp", "original"])
def test_cuda_ialspp_within_float32_tolerance_of_golden(loss):
    import irspack_b200
    from irspack_b200 import _ials_core as core

    if irspack_b200.device_count() == 0:
        pytest.fail("GPU test selected but no CUDA device is visible")
    lt = core.LossType.ORIGINAL if loss == "original" else core.LossType.IALSPP
    cfg = (core.IALSModelConfigBuilder().set_K(K).set_alpha0(ALPHA0).set_reg(REG).set_nu(NU)
           .set_loss_type(lt).build())
    sc = (core.IALSSolverConfigBuilder().set_solver_type(core.SolverType.IALSPP)
          .set_ialspp_subspace_dimension(PP_S).set_ialspp_iteration(PP_IT).build())
    g = core.IALSTrainer(cfg, X.astype(np.float32))
    g.user, g.item = G["user0"].astype(np.float32), G["item0"].astype(np.float32)
    for _ in range(int(EPOCHS)):
        g.step(sc)
    for got, want in ((g.user, GP[f"user_{loss}"]), (g.item, GP[f"item_{loss}"])):
        assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max()
    assert g.compute_loss(sc) == pytest.approx(float(GP[f"loss_{loss}"]), rel=1e-4)
