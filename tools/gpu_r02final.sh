#!/bin/bash
# End-of-round record on one B200: whole GPU suite, smoke(), bench.py, the ncu launch list of the same
# bench command and one ncu --set full capture of cg_rows_kernel (users + items launch).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
echo "== all gpu tests rc=$?"; tail -n 5 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "== bench rc=$?"; python - <<'P'
import json
for l in open('gpurun_out/bench_final.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], r['frac'], r['phases_ms_per_epoch'], d['cpu_baseline']['ms_per_epoch'], d.get('c4_single_gpu',{}).get('ms_per_epoch'))
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 2 --warmup 1 > gpurun_out/launches_final.log 2>&1
echo "== launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cg_rows_kernel -s 2 -c 2 -f -o gpurun_out/prof_rows_final \
  python tools/profile_epoch.py --epochs 2 > gpurun_out/ncu_rows_final.log 2>&1
echo "== ncu rows rc=$?"
