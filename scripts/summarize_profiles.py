"""Turns gpurun_out/launches_*.csv and prof_*.ncu-rep into the markdown summaries kept under profiles/."""
import csv
import subprocess
import sys

launch_csv, ncu_rep, out_md, title = sys.argv[1:5]
rows = list(csv.reader(open(launch_csv)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
lines = [f"# {title}", "", "## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
         "| kernel | launches | mean µs | min µs |", "|---|---|---|---|"]
for k, v in agg.items():
    lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {min(v) / 1e3:.1f} |")
spmm = {k: sum(v) / len(v) for k, v in agg.items() if "spmm" in k}
tot = sum(spmm.values())
lines += ["", "SpMM step = " + " + ".join(f"`{k}` {v / 1e3:.0f} µs ({100 * v / tot:.0f} %)" for k, v in spmm.items()), ""]
raw = subprocess.run(["ncu", "-i", ncu_rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hh = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
lines += ["## `ncu --set full --clock-control none` (one launch each)", ""]
for r in rr[2:]:
    name = r[hh.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    lines += [f"### `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
    for w in want:
        if w in hh:
            lines.append(f"| {w} | {r[hh.index(w)]} | {rr[1][hh.index(w)]} |")
    lines.append("")
open(out_md, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
