#!/bin/bash
# r02ar: launch lists of the iALS++ route at quarter size for S = 64 and S = 32 (does the block solver's time depend on S?)
mkdir -p gpurun_out
for S in 64 32; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ialspp_s$S.csv \
  python tools/time_config.py --config c2 --solver IALSPP --subspace $S --epochs 1 --scale 0.25 > gpurun_out/launches_ialspp_s$S.log 2>&1
echo "== launch list S=$S rc=$?"
python - $S <<'P'
import csv, collections, sys
rows = list(csv.reader(l for l in open(f'gpurun_out/launches_ialspp_s{sys.argv[1]}.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split('(')[0][-40:]
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] in ('ns', 'nsecond') else v  # -> us
    a = agg.setdefault(name, [0, 0.0, []]); a[0] += 1; a[1] += v; a[2].append(round(v))
for k, (n, v, l) in agg.items():
    if v > 500: print(f"{k:42s} {n:5d} launches {v/1e3:10.3f} ms  {v/n:10.1f} us each", l[:20])
P
done
