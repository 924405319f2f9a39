#!/bin/bash
# r02c: after the variant clean-up -- whole GPU suite, bench (cg_rows on FFMA2), the 256-column
# Gram operator and the tensor-core Cholesky route, configs[3] on one GPU at 10 % with a sweep of
# the heavy-row threshold.
mkdir -p gpurun_out
TAIL=8 tools/gpu_check.sh tests
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-epochs 2 --c4 off > gpurun_out/bench.log 2>&1; echo "== bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench.log") if l.startswith("{")][-1]); r=d["roofline"]
print("ms/epoch", d["ms_per_step"], "frac", r["frac"], "whole", r["whole_solve"]["frac"], "e2e", d["e2e"]["ms_per_step"], "cpu", d["cpu_baseline"]["ms_per_epoch"])
print({k: round(v, 3) for k, v in r["phases_ms_per_epoch"].items()})
PY
IALS_CHOL=tc timeout 300 python tools/parity_chol256.py > gpurun_out/parity_chol_tc.log 2>&1; echo "== chol tc parity rc=$?"; tail -n 2 gpurun_out/parity_chol_tc.log | cut -c1-600
for thr in 2048 1024 512 256 128; do
  IALS_HEAVY_THRESHOLD=$thr timeout 600 python tools/time_c4.py --scale 0.1 --steps 2 --e2e-steps 1 --score-users 32768 > gpurun_out/c4_s01_thr$thr.log 2>&1
  echo "== c4 x0.1 thr=$thr rc=$?"
  python - gpurun_out/c4_s01_thr$thr.log <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(open(sys.argv[1]).read()[-800:])
else:
    d=json.loads(l[-1]); print("ms/epoch %.2f  G int/s %.3f  GB/s %.0f"%(d["ms_per_epoch"], d["interactions_per_s"]/1e9, d["achieved_gbs"]), {k: round(v,2) for k,v in d["phases_ms_max_over_ranks"].items()}, "e2e ms", d["e2e"]["ms_per_step"] if d["e2e"] else None, "topk users/s", round(d["score_topk"]["users_per_s"]) if d["score_topk"] else None)
PY
done
