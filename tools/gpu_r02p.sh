#!/bin/bash
# r02p: wgram_kernel with two double-buffered producer groups (headline bench), scoring kernel with
# the lean single producer group + rank merge, everything through the whole GPU suite.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
echo "== all gpu tests rc=$?"; tail -n 8 gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err
echo "== bench rc=$?"; python - <<'P'
import json
for l in open('gpurun_out/bench_p.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], r['frac'], r['phases_ms_per_epoch'], d['cpu_baseline']['ms_per_epoch'])
P
timeout 300 python tools/time_recommend.py > gpurun_out/time_recommend6.log 2>&1; echo "== time_recommend rc=$?"; tail -n 4 gpurun_out/time_recommend6.log | cut -c1-300
timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_p.log 2>&1
echo "== c3 x 0.05 rc=$?"; tail -n 1 gpurun_out/c3_scaled_p.log | cut -c1-500
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 86 --launch-timeout 0 \
  python -m pytest tests/test_gpu_parity.py tests/test_wgram.py -m gpu -x -q -k "test_heavy_row_tensor_path_at_low_thresholds or test_gram or test_recommend_matches_reference_ordering" > gpurun_out/sanitize_p_memcheck.log 2>&1
echo "== memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitize_p_memcheck.log | head -n 8
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_tc -s 1 -c 1 -o gpurun_out/prof_score6 -f \
  python tools/time_recommend.py > gpurun_out/ncu_score6.log 2>&1
echo "== ncu score rc=$?"
