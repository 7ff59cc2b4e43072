"""Compact per-kernel summary of an .ncu-rep (run where ncu is installed; no GPU needed): python tools/ncu_summary.py rep [rep...]"""
import csv, subprocess, sys, io
KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"), ("smsp__inst_executed.sum", "inst"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg_throttle"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_no_inst"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "st_sleep"),
        ("launch__occupancy_limit_registers", "occ_lim_regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"), ("launch__waves_per_multiprocessor", "waves")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        parts = []
        for k, short in KEYS:
            if k in hdr:
                v = r[hdr.index(k)]; u = units[hdr.index(k)]
                try: v = "%.4g" % float(v.replace(",", ""))
                except ValueError: pass
                parts.append("%s=%s%s" % (short, v, u if u in ("ms", "us", "Mbyte", "Gbyte", "Kbyte", "byte") else ""))
        print(name[:48], "|", " ".join(parts))
