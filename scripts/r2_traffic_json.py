"""profiles/r2_traffic.json from the ncu captures of scripts/r2_traffic.sh (gpurun_out/r2_traffic_<workload>_<k>.csv):
DRAM bytes of ONE SpMM step per workload, tagged with the hash of the kernel sources they were taken on (bench.py reports
`roofline.traffic` only while that hash matches).  python scripts/r2_traffic_json.py reddit:128 yelp:128 amazon:128"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
out_path = os.path.join(ROOT, "profiles", "r2_traffic.json")
old = json.load(open(out_path)) if os.path.exists(out_path) else {}
out = {"kernels_sha": bench.kernels_sha(),
       "how": "scripts/r2_traffic.sh <workload> <k>: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the kernels of ONE SpMM "
              "step (FX_FMT_TCW default plan), summed; 1 x B200; scripts/r2_traffic_json.py wrote this file",
       "bytes": {}, "per_kernel_MB": {}}
for k in ("before_chunk_ordering", "note"):
    if k in old:
        out[k] = old[k]
for spec in sys.argv[1:]:
    wl, k = spec.split(":")
    src = os.path.join(ROOT, "gpurun_out", f"r2_traffic_{wl}_{k}.csv")
    rows = list(csv.reader(open(src)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ki, mi, vi, ui, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
    per = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            per.setdefault((int(r[ii]), r[ki].split("(")[0].split("::")[-1].split("<")[0]), {})[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
    last = collections.OrderedDict()
    for (_, name), m in per.items():  # the LAST launch of each kernel = the last (warm) step
        last[name] = m
    rec, tot = {}, 0.0
    for name, m in last.items():
        mb = lambda key: m[key][0] * UNIT.get(m[key][1], 1.0)
        t = m["gpu__time_duration.sum"]
        rec[name] = {"us": round(t[0] / 1000 if t[1] in ("ns", "nsecond") else t[0], 1), "dram_rd": round(mb("dram__bytes_read.sum"), 1),
                     "dram_wr": round(mb("dram__bytes_write.sum"), 1), "xbar_to_l1": round(mb("l1tex__m_xbar2l1tex_read_bytes.sum"), 1),
                     "l2_hit_pct": round(m["lts__t_sector_hit_rate.pct"][0], 1)}
        tot += rec[name]["dram_rd"] + rec[name]["dram_wr"]
    key = f"{wl}:{k}:tcw"
    out["bytes"][key] = int(round(tot * 1e6, -5))
    out["per_kernel_MB"][key] = rec
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(src) as f, open(os.path.join(ROOT, "profiles", f"r2_traffic_{wl}_{k}.csv"), "w") as g:
        g.write(f.read())
    print(key, "%.1f MB per step" % tot, rec)
json.dump(out, open(out_path, "w"), indent=1)
