#!/usr/bin/env python
"""Diagnostic: the 3 x 4 matrix of tests/test_gpu_parity.py::test_empty_rows_and_columns, K = 8,
GPU against the f32 / f64 oracles, per solver and per half-epoch."""
import os
import sys

import numpy as np
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from irspack_b200 import _ials_core as core  # noqa: E402
from test_gpu_parity import make_pair, solver_cfg  # noqa: E402

np.set_printoptions(precision=7, linewidth=200, suppress=True)
X = sps.csr_matrix(np.array([[1, 0, 2, 0], [0, 0, 0, 0], [3, 0, 0, 0]], dtype=np.float32))
for solver in ("CG", "CHOLESKY"):
    g, o32, o64 = make_pair(core, X, 8)
    st = oracle.SOLVER_CG if solver == "CG" else oracle.SOLVER_CHOLESKY
    sc = solver_cfg(core, solver)
    for side, name in ((0, "user"), (1, "item")):
        g.half_step(side, sc)
        for o in (o32, o64):
            if side == 0:
                o._solve(o.user, o.X, o.item, st, 3, 1)
            else:
                o._solve(o.item, o.X_t, o.user, st, 3, 1)
        G = getattr(g, name)
        A, B = getattr(o32, name), getattr(o64, name)
        print(solver, name, "gpu-o32 %.3e  gpu-o64 %.3e  o32-o64 %.3e  scale %.3e" % (
            np.abs(G - A).max(), np.abs(G - B).max(), np.abs(A - B).max(), np.abs(A).max()))
        if np.abs(G - A).max() > 1e-5:
            print("gpu\n", G, "\no32\n", A, "\no64\n", B)
        P = g.gram(side)
        Y = getattr(o64, "item" if side == 0 else "user")
        Pref = 0.1 * Y.T @ Y
        print("   gram err %.3e (max %.3e)" % (np.abs(P - Pref).max(), np.abs(Pref).max()))
