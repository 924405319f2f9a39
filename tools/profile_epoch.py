#!/usr/bin/env python
"""Runs a few iALS epochs (or a recommend pass) of a BASELINE shape, for use under ncu.

    ncu ... python tools/profile_epoch.py --shape ml20m --epochs 3 [--solver CG|CHOLESKY] [--recommend]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from irspack_b200 import _ials_core as core  # noqa: E402
from irspack_b200.synth import SHAPES, init_factors, synth_csr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="ml20m")
ap.add_argument("--epochs", type=int, default=3)
ap.add_argument("--solver", default="CG")
ap.add_argument("--K", type=int, default=0)
ap.add_argument("--scale", type=float, default=1.0, help="shrink users/nnz by this factor")
ap.add_argument("--recommend", type=int, default=0, help="also run recommend() on this many users")
ap.add_argument("--dump", default="", help="save the factors after the last epoch to this .npz")
a = ap.parse_args()
U, I, nnz, K = SHAPES[a.shape]
U, nnz = int(U * a.scale), int(nnz * a.scale)
K = a.K or K
X = synth_csr(U, I, nnz, seed=1002)
cfg = core.IALSModelConfigBuilder().set_K(K).set_alpha0(0.1).set_reg(1e-3).build()
st = core.SolverType.CG if a.solver == "CG" else core.SolverType.CHOLESKY
sc = core.IALSSolverConfigBuilder().set_solver_type(st).set_max_cg_steps(3).build()
t = core.IALSTrainer(cfg, X)
t.user, t.item = init_factors(U, K, 1), init_factors(I, K, 2)
for _ in range(a.epochs):
    t.step(sc)
if a.dump:
    import numpy as np

    np.savez(a.dump, user=t.user, item=t.item)
if a.recommend:
    t.recommend(0, min(a.recommend, U), 10)
print("done")
