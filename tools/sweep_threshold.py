#!/usr/bin/env python
"""Epoch time of the ML-20M-shaped workload for several heavy-row thresholds (env knobs of api.cu)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from irspack_b200 import _ials_core as core
from irspack_b200.synth import SHAPES, init_factors, synth_csr

U, I, nnz, K = SHAPES["ml20m"]
X = synth_csr(U, I, nnz, seed=1002)
u0, i0 = init_factors(U, K, 1), init_factors(I, K, 2)
cfg = core.IALSModelConfigBuilder().set_K(K).set_alpha0(0.1).set_reg(1e-3).build()
sc = core.IALSSolverConfigBuilder().set_max_cg_steps(3).build()
ref = None
for spec in (sys.argv[1:] or ["100000000", "384", "192", "96", "48", "24"]):
    # "T" or "T:M" (heavy threshold : mid threshold)
    thr = int(spec.split(":")[0])
    os.environ["IALS_HEAVY_THRESHOLD"] = str(thr)
    os.environ["IALS_MID_THRESHOLD"] = spec.split(":")[1] if ":" in spec else str(1 << 30)
    t = core.IALSTrainer(cfg, X)
    t.user, t.item = u0, i0
    for _ in range(2):
        t.step_async(sc)
    t.sync()
    t.set_profiling(True)
    for _ in range(5):
        t.step_async(sc)
    t.sync()
    ms, n = t.get_timings()
    t.set_profiling(False)
    # parity across thresholds: fresh start, 2 epochs
    t.user, t.item = u0, i0
    t.step(sc); t.step(sc)
    u = t.user
    if ref is None:
        ref = u
    err = np.abs(u - ref).max() / np.abs(ref).max()
    ph = [v / n for v in ms]
    print(f"threshold {spec:>10s}: gram {ph[0]:.3f}+{ph[4]:.3f}  users wgram/dense/light {ph[1]:.2f}/{ph[2]:.2f}/{ph[3]:.2f}  "
          f"items {ph[5]:.2f}/{ph[6]:.2f}/{ph[7]:.2f}  epoch {sum(ph):.3f} ms   rel diff vs first {err:.2e}", flush=True)
    del t
