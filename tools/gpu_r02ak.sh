#!/bin/bash
# r02ak: where the iALS++ tensor route spends its time: ncu launch list of two epochs at quarter size.
mkdir -p gpurun_out
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
for k, (n, v) in agg.items(): print(f"{k:42s} {n:5d} launches {v/1e3:10.3f} ms  {v/n:10.1f} us each")
P
