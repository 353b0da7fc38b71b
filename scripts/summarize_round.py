"""Builds profiles/r1_ncu_tcw_reddit_k128.md from the ncu reports in gpurun_out/ (run here, after gpurun)."""
import csv, io, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "l1tex__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    return d


def section(title, rep, note):
    d = raw(rep)
    lines = [f"## {title}", "", note, "", f"Kernel: `{d.get('Kernel Name', ('?',))[0]}`", "", "| metric | value | unit |", "|---|---|---|"]
    for m in WANT:
        if m in d:
            lines.append(f"| {m} | {d[m][0]} | {d[m][1]} |")
    return "\n".join(lines) + "\n"


def main():
    g = os.path.join(ROOT, "gpurun_out")
    md = ["# ncu summaries — default bench (Reddit-shape, k=128, FX_FMT_TCW), round 1 final", "",
          "Command: `ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 2 -c 1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline`",
          "(one capture per kernel, each after the same command exited 0 without ncu; numbers under ncu are cold-cache and",
          "serialised: use the shares, not the absolutes). Launch list of the same command: `profiles/r1_launches_reddit_k128.csv`;",
          "DRAM bytes per kernel (`--metrics dram__bytes_read.sum,dram__bytes_write.sum`): `profiles/r1_dram_reddit_k128.csv`.", ""]
    md.append(section("k_spmm_rows — remainder nz (row-grab kernel)", os.path.join(g, "rows_full.ncu-rep"),
                      "Bound by the L2->SM gather path (xbar2l1tex bytes = nz x 512 B): warps wait on B rows (long scoreboard 73 % of the stall samples, "
                      "56 % on the B requests, 15 % on the nz metadata); L1 data pipe (l1tex__data_pipe_lsu_wavefronts) 71 % busy, L2 hit 76 %, DRAM 2 TB/s."))
    md.append(section("k_spmm_special_cta — 512-nz chunks of long rows", os.path.join(g, "special_full.ncu-rep"),
                      "Same gather path, no per-row overhead: 18.2 TB/s through the L2->SM crossbar, L1 data pipe 84 % busy -- the rate the row kernel is measured against."))
    md.append(section("k_spmm_tc — tcgen05 kernel of the tensor windows", os.path.join(g, "tc_full.ncu-rep"),
                      "Tensor pipe active share = the 3xTF32 MMAs; the rest is staging (16-byte loads, tf32 split, shared-memory stores) and the tc_out write."))
    sass = subprocess.run("cuobjdump -sass %s | grep -oE 'UTCHMMA|UTCBAR[.A-Z0-9_]*|LDTM[.A-Za-z0-9_]*|UTCATOMSWS[.A-Z_]*|UBLKCP[.A-Z0-9_]*|SYNCS[.A-Z0-9_]*|FENCE.VIEW.ASYNC[.A-Z]*|FFMA2|LDS.128|LDG.E.128[.A-Z]*' | sort | uniq -c | sort -rn"
                          % os.path.join(ROOT, "flex_b200", "libflexb200.so"), shell=True, capture_output=True, text=True).stdout
    md += ["## SASS mnemonics in libflexb200.so (cuobjdump -sass, counts)", "",
           "tcgen05.mma -> `UTCHMMA`, tcgen05.commit -> `UTCBAR`, tcgen05.ld -> `LDTM`, tcgen05.alloc/dealloc -> `UTCATOMSWS`,",
           "cp.async.bulk -> `UBLKCP`, mbarrier -> `SYNCS`, fence.proxy.async -> `FENCE.VIEW.ASYNC`, fma.rn.f32x2 -> `FFMA2`.", "", "```", sass.rstrip(), "```", ""]
    open(os.path.join(ROOT, "profiles", "r1_ncu_tcw_reddit_k128.md"), "w").write("\n".join(md))
    print("written profiles/r1_ncu_tcw_reddit_k128.md")


main()
