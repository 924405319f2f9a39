#!/usr/bin/env python
"""Times the epochs of one BASELINE.json configuration on cuda:0 and prints one JSON line.

    python tools/time_config.py --config c1|c2|c3 [--epochs N] [--cpu-epochs M]

c1: ML-1M shape, K=64, CG, 10 epochs (the reference's CPU-runnable case; the oracle port is
    timed beside it on all host threads).
c2: ML-20M shape, K=128, CG (bench.py's workload, here without the e2e / roofline legs).
c3: Netflix shape, K=256, Cholesky solver (SURVEY.md 8 d: ~16.2 TFLOP per epoch).
Not a bench line: bench.py owns the contract; this records the other configurations.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from irspack_b200 import _ials_core as core  # noqa: E402
from irspack_b200.synth import SHAPES, init_factors, synth_csr  # noqa: E402

CONFIGS = {"c1": ("ml1m", "CG", 10, 1001), "c2": ("ml20m", "CG", 5, 1002), "c3": ("netflix", "CHOLESKY", 2, 1003)}

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c1", choices=sorted(CONFIGS))
ap.add_argument("--epochs", type=int, default=0)
ap.add_argument("--cpu-epochs", type=int, default=-1, help="oracle epochs on the host (default: c1 only)")
ap.add_argument("--scale", type=float, default=1.0, help="shrink users/nnz (dry runs)")
ap.add_argument("--solver", default="", help="override the configuration's solver: CG | CHOLESKY | IALSPP")
ap.add_argument("--subspace", type=int, default=64, help="ialspp_subspace_dimension")
a = ap.parse_args()
shape, solver, epochs, seed = CONFIGS[a.config]
solver = a.solver or solver
epochs = a.epochs or epochs
U, I, nnz, K = SHAPES[shape]
U, nnz = int(U * a.scale), int(nnz * a.scale)
t0 = time.perf_counter()
X = synth_csr(U, I, nnz, seed=seed)
t_synth = time.perf_counter() - t0
u0, i0 = init_factors(U, K, 1), init_factors(I, K, 2)
cfg = core.IALSModelConfigBuilder().set_K(K).set_alpha0(0.1).set_reg(1e-3).build()
st = getattr(core.SolverType, solver)
sc = (core.IALSSolverConfigBuilder().set_solver_type(st).set_max_cg_steps(3)
      .set_ialspp_subspace_dimension(a.subspace).set_ialspp_iteration(1).build())
t = core.IALSTrainer(cfg, X)
t.user, t.item = u0, i0
t.step(sc)  # warm-up epoch (step() synchronises and checks the solver status)
t.user, t.item = u0, i0
t.set_profiling(True)
t0 = time.perf_counter()
for _ in range(epochs):
    t.step(sc)
dt = (time.perf_counter() - t0) / epochs
phase_ms, n_ep = t.get_timings()
t.set_profiling(False)
line = {"config": a.config, "shape": shape, "n_users": U, "n_items": I, "nnz": nnz, "K": K, "solver": solver,
        "epochs_timed": epochs, "ms_per_epoch": 1e3 * dt, "interactions_per_s": nnz / dt,
        "synth_s": round(t_synth, 1),
        "phases_ms_per_epoch": dict(zip(("gram_item", "users_heavy_gram", "users_heavy_dense", "users_rows",
                                         "gram_user", "items_heavy_gram", "items_heavy_dense", "items_rows"),
                                        (round(v / max(n_ep, 1), 3) for v in phase_ms)))}
if solver == "CHOLESKY":  # SURVEY.md 8 d: 2 nnz K(K+1) + (U+I)(K^3/3 + 2K^2) + 4 nnz K
    flops = 2.0 * nnz * K * (K + 1) + (U + I) * (K ** 3 / 3.0 + 2.0 * K * K) + 4.0 * nnz * K
    line["algorithmic_tflop_per_epoch"] = flops / 1e12
    line["tflops"] = flops / dt / 1e12
cpu_epochs = a.cpu_epochs if a.cpu_epochs >= 0 else (epochs if a.config == "c1" else 0)
if cpu_epochs:
    import oracle

    nt = oracle.hardware_threads()
    o = oracle.OracleTrainer(X, K, 0.1, 1e-3, 1.0, oracle.LOSS_IALSPP)
    o.user, o.item = u0.copy(), i0.copy()
    osolver = {"CG": oracle.SOLVER_CG, "CHOLESKY": oracle.SOLVER_CHOLESKY, "IALSPP": oracle.SOLVER_IALSPP}[solver]
    o.ialspp_subspace_dimension = a.subspace
    t0 = time.perf_counter()
    for _ in range(cpu_epochs):
        o.step(osolver, 3, nt)
    cdt = (time.perf_counter() - t0) / cpu_epochs
    g_u, g_i = t.user, t.item
    line["cpu_port"] = {"ms_per_epoch": 1e3 * cdt, "cores": nt, "epochs": cpu_epochs}
    if cpu_epochs == epochs:  # same trajectory: report the distance as well
        line["max_rel_diff_vs_oracle"] = float(max(np.abs(g_u - o.user).max() / np.abs(o.user).max(),
                                                   np.abs(g_i - o.item).max() / np.abs(o.item).max()))
print(json.dumps(line))
