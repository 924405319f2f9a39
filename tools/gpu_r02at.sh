#!/bin/bash
# r02at: ialspp_dense, blocked substitutions, clean build: timing + launch list.
mkdir -p gpurun_out
timeout 300 python tools/time_config.py --config c2 --solver IALSPP --epochs 3 > gpurun_out/ialspp_c2_at.log 2>&1
echo "== c2 IALSPP rc=$?"; tail -n 1 gpurun_out/ialspp_c2_at.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ialspp.csv \
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
    a = agg.setdefault(name, [0, 0.0, []]); a[0] += 1; a[1] += v; a[2].append(round(v))
for k, (n, v, l) in agg.items():
    if v > 500: print(f"{k:42s} {n:5d} launches {v/1e3:10.3f} ms  {v/n:10.1f} us each", l[:12])
P
