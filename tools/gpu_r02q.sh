#!/bin/bash
# r02q: the committed state (wgram_kernel with its 16 producer warps again, wgram256_kernel, ll v2, scoring
# with the lean producer group): whole GPU suite, headline bench + launch list, configs[2] full size.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
echo "== all gpu tests rc=$?"; tail -n 8 gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
echo "== bench rc=$?"; python - <<'P'
import json
for l in open('gpurun_out/bench_q.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], r['frac'], r['phases_ms_per_epoch'], d['cpu_baseline']['ms_per_epoch'])
P
timeout 900 python tools/time_c3_sharded.py --cpu-sample 0.02 > gpurun_out/c3_sharded_1q.log 2>&1
echo "== c3 sharded driver, 1 GPU rc=$?"; tail -n 1 gpurun_out/c3_sharded_1q.log | cut -c1-700
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_q_launches.csv \
  python bench.py --steps 2 --warmup 1 > gpurun_out/bench_q_ncu.log 2>&1
echo "== launch list rc=$?"
