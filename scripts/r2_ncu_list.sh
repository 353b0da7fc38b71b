#!/bin/bash
# per-kernel durations of a sweep (ncu launch list; cold-cache serialised times: shares, not absolutes)
# usage: scripts/r2_ncu_list.sh <tag> <workload> <k> plans...
tag=$1; shift
STEPS=2 WARM=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_spmm -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
  python scripts/r2_sweep.py "$@" > gpurun_out/${tag}_ncu.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/${tag}_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    name = r[ki].split("(")[0][:60]
    v = float(r[vi].replace(",", "")); u = r[ui]
    us = v / 1000 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000)
    seq.append((name, us))
for name, us in seq:
    if "spmm" in name: print("%-62s %9.1f us" % (name, us))
PY
