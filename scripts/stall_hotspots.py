"""Source-page summary of the ncu captures in gpurun_out/ (run here after gpurun): where the warps of each SpMM kernel wait.
Writes profiles/r1_stall_hotspots.md: stall-reason totals, the SASS lines with the most samples and the regions by
executed-instruction share.  Input: `ncu --set full --import-source on` reports of scripts/final_profile.sh."""
import csv, io, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0][1], rows[1], rows[2:]


def section(title, rep, top=14):
    kernel, hdr, data = source_page(rep)
    ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    idx = [hdr.index(c) for c in cols]
    tot = sum(int(r[isamp]) for r in data)
    instr = sum(int(r[iex]) for r in data)
    agg = sorted(((sum(int(r[i]) for r in data), c) for c, i in zip(cols, idx)), reverse=True)
    md = [f"## {title}", "", f"Kernel: `{kernel}`", "", f"{tot} stall samples, {instr} warp instructions executed.", "",
          "Stall reasons (share of samples): " + ", ".join(f"{c[6:]} {100 * v / tot:.0f} %" for v, c in agg if v > 0.02 * tot), "",
          "| SASS line | samples | share | executed | main reason |", "|---|---|---|---|---|"]
    order = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top]
    for i in order:
        r = data[i]
        why = max(zip((int(r[j]) for j in idx), cols))[1][6:]
        md.append(f"| `{' '.join(r[ia].split())[:70]}` | {r[isamp]} | {100 * int(r[isamp]) / tot:.1f} % | {r[iex]} | {why} |")
    return "\n".join(md) + "\n"


def main():
    g = os.path.join(ROOT, "gpurun_out")
    md = ["# Where the warps wait — source pages of the round's final ncu captures (default bench, Reddit-shape k=128)", "",
          "`python scripts/stall_hotspots.py` over `gpurun_out/{rows,special,tc}_full.ncu-rep` (scripts/final_profile.sh). A line's",
          "samples are charged to the instruction that WAITS, i.e. the first consumer of an outstanding load: `FFMA2 ... lon` lines are",
          "the B-row requests of the group before them.", ""]
    md.append(section("k_spmm_rows", os.path.join(g, "rows_full.ncu-rep")))
    md.append(section("k_spmm_special_cta", os.path.join(g, "special_full.ncu-rep"), top=8))
    md.append(section("k_spmm_tc", os.path.join(g, "tc_full.ncu-rep")))
    open(os.path.join(ROOT, "profiles", "r1_stall_hotspots.md"), "w").write("\n".join(md))
    print("written profiles/r1_stall_hotspots.md")


main()
