#!/bin/bash
# per-kernel durations of ONE tile build (ncu launch list of the builder kernels): scripts/r2_build_list.sh <tag> <workload> <k>
tag=$1; shift
STEPS=1 WARM=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_tcw|k_heavy|k_fill|k_detect|k_scan|k_worklist|k_panel|k_pad|k_nodense" -c 80 --csv --log-file gpurun_out/${tag}_build.csv \
  python scripts/r2_sweep.py "$@" 4:256:224:1024 > gpurun_out/${tag}_build.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/${tag}_build.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
seq = [(r[ki].split("(")[0][:48], float(r[vi].replace(",", "")) / (1000 if r[ui] in ("ns", "nsecond") else 1)) for r in rows[hdr + 1:] if len(r) > vi]
# the last build of the log = the last occurrence of each kernel after the final k_tcw_pad
last = max(i for i, (n, _) in enumerate(seq) if "k_tcw_pad" in n)
tot = 0
for n, us in seq[last:]:
    print("%-50s %9.1f us" % (n, us)); tot += us
print("sum %.1f us" % tot)
PY
