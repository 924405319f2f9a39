#!/usr/bin/env python
"""Times the Evaluator's device pass (scores + seen mask + top-k) on a BASELINE shape.

    python tools/time_recommend.py [--shape ml20m] [--users N] [--k 10] [--epochs 1]

Prints one JSON line per call pattern of the fused tcgen05 kernel: all users in one call, and the
Evaluator's pattern (blocks of --block users): wall milliseconds of IALSTrainer.recommend
(synchronous C-ABI call, k indices + counts copied back) and the algorithmic GEMM rate 2*U*I*K / t.
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
ap.add_argument("--block", type=int, default=4096)
ap.add_argument("--allow", type=float, default=0.5,
                help="fraction of the catalogue in the allow-list patterns (0: skip them)")
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
t.recommend(0, min(n, 4096), a.k)  # warm-up (scratch allocation, sortedness check)
for name, block in (("one call", n), (f"blocks of {a.block}", a.block)):
    best = float("inf")
    for _ in range(a.reps):
        t0 = time.perf_counter()
        for b in range(0, n, block):
            t.recommend(b, min(b + block, n), a.k)
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"pattern": name, "users": n, "items": I, "K": K, "k": a.k, "ms": 1e3 * best,
                      "tflops_algorithmic": 2.0 * n * I * K / best / 1e12,
                      "users_per_s": n / best}), flush=True)
if a.allow > 0:  # the Evaluator's recommendable items fused into the same kernel
    rng = np.random.default_rng(3)
    shared = np.sort(rng.choice(I, max(1, int(a.allow * I)), replace=False)).astype(np.int32)
    m = min(n, 32768)
    per_len = rng.integers(0, 1000, size=m).clip(0, I)  # candidate lists to re-rank
    ip = np.zeros(m + 1, np.int64)
    np.cumsum(per_len, out=ip[1:])
    flat = np.concatenate([np.sort(rng.choice(I, int(c), replace=False)) for c in per_len]).astype(np.int32)
    for name, rows, allowed in (("one call, shared allow-list", n, (1, np.array([0, shared.size]), shared)),
                                ("one call, per-user allow-lists", m, (m, ip, flat))):
        best = float("inf")
        for _ in range(a.reps):
            t0 = time.perf_counter()
            t.recommend(0, rows, a.k, allowed=allowed)
            best = min(best, time.perf_counter() - t0)
        print(json.dumps({"pattern": name, "users": rows, "items": I, "K": K, "k": a.k,
                          "allowed_entries": int(allowed[2].size), "ms": 1e3 * best,
                          "tflops_algorithmic": 2.0 * rows * I * K / best / 1e12,
                          "users_per_s": rows / best}), flush=True)
pick = np.random.default_rng(4).integers(0, U, size=min(n, 32768))
best = float("inf")
for _ in range(a.reps):  # serving: users picked by index (gathered on the device), own rows masked
    t0 = time.perf_counter()
    t.recommend_users(pick, a.k, mask="train", return_scores=True)
    best = min(best, time.perf_counter() - t0)
print(json.dumps({"pattern": "users picked by index, one call", "users": int(pick.size), "items": I, "K": K,
                  "k": a.k, "ms": 1e3 * best, "tflops_algorithmic": 2.0 * pick.size * I * K / best / 1e12,
                  "users_per_s": pick.size / best}), flush=True)
