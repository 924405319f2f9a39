#!/usr/bin/env python
"""Times the Evaluator's device pass (scores + seen mask + top-k) on a BASELINE shape.

    python tools/time_recommend.py [--shape ml20m] [--users N] [--k 10] [--epochs 1]

Prints one JSON line per path (fused tcgen05 kernel, and the three-kernel FP32 SIMT path with
IALS_SCORE=simt): wall milliseconds of IALSTrainer.recommend over all users (synchronous C-ABI
call, k indices + scores + counts copied back), the algorithmic GEMM rate 2*U*I*K / t, and
whether both paths return identical lists.
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

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="ml20m")
ap.add_argument("--users", type=int, default=0)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--epochs", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--skip-simt", action="store_true")
a = ap.parse_args()
U, I, nnz, K = SHAPES[a.shape]
X = synth_csr(U, I, nnz, seed=1002)
cfg = core.IALSModelConfigBuilder().set_K(K).set_alpha0(0.1).set_reg(1e-3).build()
sc = core.IALSSolverConfigBuilder().set_max_cg_steps(3).build()
t = core.IALSTrainer(cfg, X)
t.user, t.item = init_factors(U, K, 1), init_factors(I, K, 2)
for _ in range(a.epochs):
    t.step(sc)
n = a.users or U
res = {}
for path in (["tc"] if a.skip_simt else ["tc", "simt"]):
    if path == "simt":
        os.environ["IALS_SCORE"] = "simt"
    else:
        os.environ.pop("IALS_SCORE", None)
    t.recommend(0, min(n, 4096), a.k)  # warm-up (scratch allocation, sortedness check)
    best = float("inf")
    for _ in range(a.reps):
        t0 = time.perf_counter()
        idx, cnt = t.recommend(0, n, a.k)
        best = min(best, time.perf_counter() - t0)
    res[path] = idx
    print(json.dumps({"path": path, "users": n, "items": I, "K": K, "k": a.k, "ms": 1e3 * best,
                      "tflops_algorithmic": 2.0 * n * I * K / best / 1e12,
                      "users_per_s": n / best}), flush=True)
if len(res) == 2:
    same = (res["tc"] == res["simt"]).all(axis=1).mean()
    print(json.dumps({"identical_lists_fraction": float(same)}), flush=True)
