#!/bin/bash
# per-kernel durations of one TCW configuration (ncu launch list; cold-cache, serialised)
mkdir -p gpurun_out
T=${1:-4}; W=${2:-256}
timeout 600 python bench.py --fmt tcw --tc-threshold $T --tc-width $W --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/tcw_plain.log 2>&1 || { tail -5 gpurun_out/tcw_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/tcw_launches.csv \
  python bench.py --fmt tcw --tc-threshold $T --tc-width $W --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/tcw_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/tcw_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    a = agg.setdefault(r[ki][:70], [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items(): print('%-72s n=%3d avg=%10.1f us' % (k, n, t / n / 1000))
PY
