#!/bin/bash
# A/B bench runs under different environment settings:  tools/gpu_ab.sh "NAME=VAL ..." "NAME=VAL ..." ...
# Each configuration runs bench.py (no CPU baseline) and prints ms/epoch + the phase table.
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  log=gpurun_out/ab_$i.log
  env $cfg timeout 600 python bench.py --steps 10 --warmup 3 --cpu-epochs 0 > $log 2>&1
  echo "== [$cfg] rc=$?"
  python - "$log" <<'PY'
import json, sys
lines = [l for l in open(sys.argv[1]) if l.startswith("{")]
if not lines:
    print(open(sys.argv[1]).read()[-1500:])
else:
    d = json.loads(lines[-1]); r = d["roofline"]
    print("ms/epoch %.3f  e2e %.3f  whole-solve frac %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], r["whole_solve"]["frac"]))
    print({k: round(v, 3) for k, v in r["phases_ms_per_epoch"].items()})
PY
done
