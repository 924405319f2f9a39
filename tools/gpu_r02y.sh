#!/bin/bash
# r02y: cg_rows_kernel with the sweep over P shared by pairs of warps: parity, racecheck, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
echo "== all gpu tests rc=$?"; tail -n 8 gpurun_out/t_all.log
for tool in memcheck racecheck; do
timeout 600 compute-sanitizer --tool $tool --error-exitcode 86 --launch-timeout 0 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(test_half_steps and CG and 128) or test_cg_row_length_boundaries or test_empty_rows_and_columns or test_step_io" > gpurun_out/sanitize_y_$tool.log 2>&1
echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard" gpurun_out/sanitize_y_$tool.log | head -n 8
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err
echo "== bench rc=$?"; python - <<'P'
import json
for l in open('gpurun_out/bench_y.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], r['frac'], r['phases_ms_per_epoch'], d['cpu_baseline']['ms_per_epoch'])
P
