#!/bin/bash
# r02ax: ranking metrics on the device (ials_metrics_accumulate): parity with the host bookkeeping, the evaluator
# tests, and the wall clock of a whole Evaluator pass over configs[1]'s users.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_evaluator_wide.py tests/test_score_tc.py tests/test_gpu_parity.py tests/test_host_api.py -m gpu -q -x \
  -k "metrics or evaluator or Evaluator or ndcg or recommender" > gpurun_out/t_ax.log 2>&1
echo "== tests rc=$?"; tail -n 6 gpurun_out/t_ax.log
timeout 300 python - <<'P'
import time, numpy as np, scipy.sparse as sps
from irspack_b200 import Evaluator, IALSRecommender
from irspack_b200.synth import SHAPES, synth_csr, init_factors
U, I, nnz, K = SHAPES["ml20m"]
X = synth_csr(U, I, nnz, seed=1002)
rng = np.random.default_rng(1)
te = sps.random(U, I, density=4e6 / (U * I), random_state=3, format="csr", dtype=np.float32); te.data[:] = 1
rec = IALSRecommender(X, n_components=K, alpha0=0.1, reg=1e-3, train_epochs=1).learn()
ev = Evaluator(te, cutoff=10)
ev.get_score(rec)
t0 = time.perf_counter(); d = ev.get_score(rec); dt = time.perf_counter() - t0
print({"evaluator_pass_ms": round(1e3 * dt, 1), "users": U, "items": I, "cutoff": 10, "ndcg": d["ndcg"]})
P
