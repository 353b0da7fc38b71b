"""After scripts/r2_final.sh: gpurun_out/r2_final/* and the traffic CSVs -> profiles/ (traffic JSON tagged with the kernel hash,
launch list, reference-arm line, ncu summary of the three SpMM kernels).  The one-GPU bench line is copied only if it carries the
traffic of the matching capture (otherwise run `python bench.py` once more after this script and copy that line)."""
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
O = os.path.join(ROOT, "gpurun_out", "r2_final")
P = os.path.join(ROOT, "profiles")
subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "r2_traffic_json.py"), "reddit:128", "yelp:128", "amazon:128"], check=True, stdout=subprocess.DEVNULL)
shutil.copy(os.path.join(O, "launches_reddit_k128.csv"), os.path.join(P, "r2_launches_reddit_k128.csv"))
shutil.copy(os.path.join(O, "bench_reference.json"), os.path.join(P, "r2_bench_reference_arm.json"))
if os.path.exists(os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")):
    shutil.copy(os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"), os.path.join(P, "r2_parity_report.jsonl"))
d = json.loads(open(os.path.join(O, "bench_1gpu.json")).read().strip().splitlines()[-1])
if d["roofline"]["traffic"] is not None:
    shutil.copy(os.path.join(O, "bench_1gpu.json"), os.path.join(P, "r2_bench_1gpu.json"))
parts = []
for i, kn in enumerate(("k_spmm_rows", "k_spmm_tc", "k_spmm_special_cta")):
    tmp = f"/tmp/r2_{kn}.md"
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "summarize_profiles.py"), os.path.join(O, "launches_reddit_k128.csv"),
                    os.path.join(O, f"{kn}_full.ncu-rep"), tmp, f"Round 2 — `{kn}`, default bench (Reddit-shape k=128, FX_FMT_TCW), 1 x B200"],
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    s = open(tmp).read()
    parts.append(s if i == 0 else s[s.index("### "):])
s = "\n".join(parts).replace("# Round 2 — `k_spmm_rows`", "# Round 2 — the three SpMM kernels")
tot = 0.0
for m in re.finditer(r"\| `(k_tcw_\w+|k_heavy|k_fill\w*|k_detect|k_pad_rowptr|k_scan_tcount|k_panel_lists|k_special_keys|k_worklist)` \| (\d+) \| ([0-9.]+)", s):
    n = int(m.group(2))
    tot += float(m.group(3)) * (n // 4 if n >= 4 else 1)
km = d["roofline"]["kernel_ms"]
rows = float(re.search(r"`k_spmm_rows<128[^`]*` \| \d+ \| ([0-9.]+)", s).group(1))
tc = float(re.search(r"`fxtc::k_spmm_tc<128>` \| \d+ \| ([0-9.]+)", s).group(1))
sp = float(re.search(r"`k_spmm_special_cta<128>` \| \d+ \| ([0-9.]+)", s).group(1))
T = rows + tc + sp
a = s.index("SpMM step = ")
b = s.index("\n", a)
s = s[:a] + (f"SpMM step (device-buffer path, the `<128>` kernels) = `k_spmm_tc<128>` {tc:.0f} µs ({100 * tc / T:.1f} %) + `k_spmm_special_cta<128>` {sp:.0f} µs ({100 * sp / T:.1f} %) + "
             f"`k_spmm_rows<128,…>` {rows:.0f} µs ({100 * rows / T:.1f} %) = {T:.0f} µs cold-cache and serialised; live in the same command's unprofiled run "
             f"(`roofline.kernel_ms` of its bench line): {km['k_spmm_tc']:.3f} + {km['k_spmm_special_cta']:.3f} + {km['k_spmm_rows']:.3f} ms with events between the kernels, "
             f"{d['ms_per_step']:.3f} ms per step without them. The `<64>` kernels are the end-to-end leg (`fx_spmm_host`): two 64-column chunks x four row groups, each launch covering "
             f"a quarter of the panels; the `k_tcw_*` … `k_worklist` rows are the tile builds of the command: {tot / 1000:.2f} ms of kernels per build.") + s[b:]
s = s.replace("## Launch list", "Command: `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-amazon` (scripts/r2_final.sh, collected by scripts/r2_final_collect.py); raw launch "
              "list: `profiles/r2_launches_reddit_k128.csv`; DRAM bytes per step and per kernel: `profiles/r2_traffic.json` (+ `r2_traffic_*_128.csv`).\n\n## Launch list", 1)
open(os.path.join(P, "r2_ncu_reddit_k128.md"), "w").write(s)
print("bench line: value %.0f ms %.4f e2e %.3f ms tPre %.3f traffic %s | build kernels %.2f ms" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["tPre_ms"], d["roofline"]["traffic"], tot / 1000))
