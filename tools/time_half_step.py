#!/usr/bin/env python
"""Wall-clock of repeated user half-epochs (Gram + solve, synchronous) of configs[1]; used for
timing experiments whose arithmetic is deliberately wrong (the result is never fed back)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irspack_b200 import _ials_core as core  # noqa: E402
from irspack_b200.synth import SHAPES, init_factors, synth_csr  # noqa: E402

U, I, nnz, K = SHAPES["ml20m"]
X = synth_csr(U, I, nnz, seed=1002)
cfg = core.IALSModelConfigBuilder().set_K(K).set_alpha0(0.1).set_reg(1e-3).build()
sc = core.IALSSolverConfigBuilder().set_max_cg_steps(3).build()
t = core.IALSTrainer(cfg, X)
t.user, t.item = init_factors(U, K, 1), init_factors(I, K, 2)
for side in (0, 1):
    t.half_step(side, sc)
    t.user, t.item = init_factors(U, K, 1), init_factors(I, K, 2)
    best = 1e9
    for _ in range(8):
        t0 = time.perf_counter()
        t.half_step(side, sc)
        best = min(best, time.perf_counter() - t0)
        t.user, t.item = init_factors(U, K, 1), init_factors(I, K, 2)
    print(f"side {side}: {1e3 * best:.3f} ms per half-epoch", flush=True)
