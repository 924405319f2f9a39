#!/bin/bash
# r02u: step_io, flags raised by the copy engine.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "step_io or empty_rows or half_steps" > gpurun_out/t_u.log 2>&1
echo "== tests rc=$?"; tail -n 4 gpurun_out/t_u.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err
echo "== bench rc=$?"; python - <<'P'
import json
for l in open('gpurun_out/bench_u.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], r['frac'], d['cpu_baseline']['ms_per_epoch'])
P
