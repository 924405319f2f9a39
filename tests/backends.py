"""Two interchangeable back ends for the restated reference invariants:

* ``OracleBackend``  -- oracle/ (CPU restatement; the checker), float32 or float64
* ``GpuBackend``     -- irspack_b200._ials_core.IALSTrainer (the product, CUDA)

Both expose the slice of ``_ials_core.IALSTrainer`` the reference's tests use.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sps

import oracle


def scale_log(X: sps.csr_matrix, epsilon):
    if epsilon is None:
        return X
    X = sps.csr_matrix(X, dtype=np.float64, copy=True)
    X.data = np.log(1 + X.data / epsilon)  # ials.py:437-446
    return X


class OracleBackend:
    name = "oracle"

    def __init__(self, X, K, alpha0, reg, nu, loss_type, solver, max_cg_steps=3,
                 pred_cg_steps=5, epsilon=None, dtype=np.float32, seed=42, subspace=64,
                 pred_ialspp_iteration=7):
        self.solver = {"CG": oracle.SOLVER_CG, "CHOLESKY": oracle.SOLVER_CHOLESKY,
                       "IALSPP": oracle.SOLVER_IALSPP}[solver]
        self.max_cg_steps, self.pred_cg_steps, self.epsilon = max_cg_steps, pred_cg_steps, epsilon
        self.pred_ialspp_iteration = pred_ialspp_iteration
        lt = oracle.LOSS_ORIGINAL if loss_type == "ORIGINAL" else oracle.LOSS_IALSPP
        self.t = oracle.OracleTrainer(scale_log(X, epsilon), K, alpha0, reg, nu, lt, dtype=dtype,
                                      seed=seed)
        self.t.ialspp_subspace_dimension = subspace

    user = property(lambda s: s.t.user, lambda s, v: setattr(s.t, "user", np.ascontiguousarray(v, s.t.dtype)))
    item = property(lambda s: s.t.item, lambda s, v: setattr(s.t, "item", np.ascontiguousarray(v, s.t.dtype)))

    def step(self):
        self.t.ialspp_iteration = 1  # ials.py:113-118: one sweep per training epoch
        self.t.step(self.solver, self.max_cg_steps)

    def transform_user(self, X):
        self.t.ialspp_iteration = self.pred_ialspp_iteration  # ials.py:130-138
        return self.t.transform_user(scale_log(X, self.epsilon), self.solver, self.pred_cg_steps)

    def transform_item(self, X):
        self.t.ialspp_iteration = self.pred_ialspp_iteration
        return self.t.transform_item(scale_log(X, self.epsilon), self.solver, self.pred_cg_steps)

    def compute_loss(self):
        return self.t.compute_loss()

    def user_scores(self, b, e):
        return self.t.user_scores(b, e)


class GpuBackend:
    name = "gpu"

    def __init__(self, X, K, alpha0, reg, nu, loss_type, solver, max_cg_steps=3,
                 pred_cg_steps=5, epsilon=None, dtype=np.float32, seed=42, subspace=64,
                 pred_ialspp_iteration=7):
        from irspack_b200 import _ials_core as core

        assert np.dtype(dtype) == np.float32
        self.core = core
        self.epsilon = epsilon
        cfg = (core.IALSModelConfigBuilder().set_K(K).set_alpha0(alpha0).set_reg(reg).set_nu(nu)
               .set_loss_type(getattr(core.LossType, loss_type)).set_random_seed(seed).build())
        st = getattr(core.SolverType, solver)
        self.sc = (core.IALSSolverConfigBuilder().set_solver_type(st).set_max_cg_steps(max_cg_steps)
                   .set_ialspp_subspace_dimension(subspace).set_ialspp_iteration(1).build())
        self.psc = (core.IALSSolverConfigBuilder().set_solver_type(st).set_max_cg_steps(pred_cg_steps)
                    .set_ialspp_subspace_dimension(subspace)
                    .set_ialspp_iteration(pred_ialspp_iteration).build())
        self.t = core.IALSTrainer(cfg, sps.csr_matrix(scale_log(X, epsilon)).astype(np.float32))
        # same deterministic start as the oracle backend
        rng = np.random.default_rng(seed)
        U, I = X.shape
        scale = 0.1 / math.sqrt(K)
        self.t.user = (rng.standard_normal((U, K)) * scale).astype(np.float32)
        self.t.item = (rng.standard_normal((I, K)) * scale).astype(np.float32)

    user = property(lambda s: s.t.user, lambda s, v: setattr(s.t, "user", v))
    item = property(lambda s: s.t.item, lambda s, v: setattr(s.t, "item", v))

    def step(self):
        self.t.step(self.sc)

    def transform_user(self, X):
        return self.t.transform_user(sps.csr_matrix(scale_log(X, self.epsilon)).astype(np.float32), self.psc)

    def transform_item(self, X):
        return self.t.transform_item(sps.csr_matrix(scale_log(X, self.epsilon)).astype(np.float32), self.psc)

    def compute_loss(self):
        return self.t.compute_loss(self.sc)

    def user_scores(self, b, e):
        return self.t.user_scores(b, e, self.sc)
