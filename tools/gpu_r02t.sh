#!/bin/bash
# r02t: step_io with the user factors arriving in flagged chunks during the user half-epoch.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
echo "== all gpu tests rc=$?"; tail -n 8 gpurun_out/t_all.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 86 --launch-timeout 0 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_step_io_equals_set_step_get" > gpurun_out/sanitize_t_memcheck.log 2>&1
echo "== memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitize_t_memcheck.log | head -n 8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
echo "== bench rc=$?"; python - <<'P'
import json
for l in open('gpurun_out/bench_t.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], r['frac'], d['cpu_baseline']['ms_per_epoch'])
P
