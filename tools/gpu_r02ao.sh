#!/bin/bash
# r02ao: iALS++ tensor route, branch-free elimination: parity + timing + launch list + ncu of the dense kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -k "ialspp or IALSPP or golden" > gpurun_out/t_ao.log 2>&1
echo "== ialspp tests rc=$?"; tail -n 5 gpurun_out/t_ao.log
for chunk in 4096; do
  IALS_GS_CHUNK=$chunk timeout 300 python tools/time_config.py --config c2 --solver IALSPP --epochs 3 > gpurun_out/ialspp_c2_$chunk.log 2>&1
  echo "== c2 IALSPP chunk $chunk rc=$?"; tail -n 1 gpurun_out/ialspp_c2_$chunk.log | cut -c1-600
done
IALS_GS_CHUNK=4096 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ialspp.csv \
  python tools/time_config.py --config c2 --solver IALSPP --epochs 1 --scale 0.25 > gpurun_out/launches_ialspp.log 2>&1
echo "== launch list rc=$?"
python - <<'P'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/launches_ialspp.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split('(')[0][-40:]
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] in ('ns', 'nsecond') else v  # -> us
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, v) in agg.items():
    if v > 500: print(f"{k:42s} {n:5d} launches {v/1e3:10.3f} ms  {v/n:10.1f} us each")
P
IALS_GS_CHUNK=4096 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ialspp_dense_kernel -s 3 -c 1 -f -o gpurun_out/prof_gs \
  python tools/time_config.py --config c2 --solver IALSPP --epochs 1 --scale 0.25 > gpurun_out/ncu_gs.log 2>&1
echo "== ncu dense rc=$?"
