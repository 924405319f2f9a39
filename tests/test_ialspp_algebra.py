"""The reformulation behind csrc/ialspp_dense.cu, checked on the CPU against the f64 oracle (itself
pinned to the reference's own IALSTrainer.hpp, tests/test_oracle_vs_reference_trainer.py):

iALS++ (Solver::step_ialspp, IALSTrainer.hpp:387-535) keeps a prediction per stored entry and walks a
row's neighbours three times per block.  With G = sum c y y^T, b = sum (c + bias) y and
A = P + G + reg I its block step is  A_DD delta = (A x - b)_D,  x_D -= delta  -- block Gauss-Seidel
on the row's normal equations -- and the kernel solves each block by an elimination WITHOUT square
roots (unscaled pivot rows, the right-hand side riding along as an extra column) restricted to the
block's upper triangle.  This file replays exactly those operations in numpy."""
import numpy as np
import pytest
import scipy.sparse as sps

import oracle
from irspack_b200.synth import init_factors, synth_csr


def block_gauss_seidel(A, b, x, S, sweeps):
    K = x.size
    x = x.copy()
    for _ in range(sweeps):
        for d0 in range(0, K, S):
            Sd = min(S, K - d0)
            z = (A @ x - b)[d0:d0 + Sd].copy()          # residual of the block
            B = A[d0:d0 + Sd, d0:d0 + Sd].copy()
            for k in range(Sd):                           # right-looking, unscaled rows, upper triangle
                piv = B[k, k]
                assert piv > 0
                for j in range(k + 1, Sd):
                    akj = B[k, j] / piv
                    B[k + 1:j + 1, j] -= B[k, k + 1:j + 1] * akj
                    z[j] -= akj * z[k]
            d = np.zeros(Sd)
            for j in range(Sd - 1, -1, -1):               # backward substitution on the unscaled rows
                d[j] = z[j] / B[j, j]
                z[:j] -= B[:j, j] * d[j]
            x[d0:d0 + Sd] -= d
    return x


@pytest.mark.parametrize("K,S,sweeps,loss,bias", [(20, 6, 2, oracle.LOSS_IALSPP, 0.0),
                                                  (32, 8, 1, oracle.LOSS_ORIGINAL, 0.1),
                                                  (24, 24, 1, oracle.LOSS_ORIGINAL, 0.1),
                                                  (17, 1, 3, oracle.LOSS_IALSPP, 0.0)])
def test_block_gauss_seidel_on_the_normal_equations_is_ialspp(K, S, sweeps, loss, bias):
    X = sps.csr_matrix(synth_csr(90, 60, 1500, seed=3, values="counts"))
    U, I = X.shape
    alpha0, reg, nu = 0.1, 0.02, 1.0
    o = oracle.OracleTrainer(X, K, alpha0, reg, nu, loss, dtype=np.float64)
    u0, i0 = init_factors(U, K, 1).astype(np.float64), init_factors(I, K, 2).astype(np.float64)
    o.user, o.item = u0.copy(), i0.copy()
    o.ialspp_subspace_dimension, o.ialspp_iteration = S, sweeps
    o._solve(o.user, o.X, o.item, oracle.SOLVER_IALSPP, 3, 1)
    P = alpha0 * i0.T @ i0
    got = np.empty_like(u0)
    for u in range(U):
        idx = X.indices[X.indptr[u]:X.indptr[u + 1]]
        c = X.data[X.indptr[u]:X.indptr[u + 1]].astype(np.float64)
        Y = i0[idx]
        A = P + (Y * c[:, None]).T @ Y + reg * (alpha0 * I + idx.size) ** nu * np.eye(K)
        b = ((c + bias)[:, None] * Y).sum(axis=0)
        got[u] = block_gauss_seidel(A, b, u0[u], S, sweeps)
    assert np.abs(got - o.user).max() <= 1e-12 * np.abs(o.user).max()
