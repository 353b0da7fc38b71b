#!/bin/bash
# DRAM bytes of one SpMM step (sum over its kernels) from ncu: scripts/r2_traffic.sh <workload> <k>  -> gpurun_out/r2_traffic_<workload>_<k>.csv
wl=$1; k=$2
STEPS=1 WARM=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none \
  -k regex:k_spmm -c 12 --csv --log-file gpurun_out/r2_traffic_${wl}_${k}.csv python scripts/r2_sweep.py $wl $k 4:256:224:1024 > gpurun_out/r2_traffic_${wl}_${k}.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_traffic_${wl}_${k}.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ki, mi, vi, ui, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
per = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    key = (int(r[ii]), r[ki].split("(")[0][-40:])
    per.setdefault(key, {})[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
def mb(x):
    v, u = x
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
# the LAST tcw step = the last launch of each of the three kernels
last = {}
for (i, name), m in per.items(): last[name] = m
tot = 0
for name, m in last.items():
    rd, wr = mb(m["dram__bytes_read.sum"]), mb(m["dram__bytes_write.sum"])
    t = m["gpu__time_duration.sum"]; us = t[0] / 1000 if t[1] in ("ns", "nsecond") else t[0]
    xb = mb(m["l1tex__m_xbar2l1tex_read_bytes.sum"])
    print("%-42s %8.1f us  dram rd %8.1f MB wr %8.1f MB  xbar->L1 %9.1f MB  L2 hit %5.1f %%" % (name, us, rd, wr, xb, m["lts__t_sector_hit_rate.pct"][0]))
    tot += rd + wr
print("TOTAL_DRAM_MB %.1f" % tot)
PY
