"""Compact per-kernel summary of an .ncu-rep (reads `ncu --page raw --csv`)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h = r[0]
keys = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "dur"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf")]
idx = [(h.index(k), n) for k, n in keys if k in h]
units = r[1]
for row in r[2:]:
    parts = []
    for i, n in idx:
        v = row[i]
        if n == "kernel":
            v = v.split("(")[0].replace("void <unnamed>::", "")[:48]
        else:
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            if n in ("dur", "dram_rd", "dram_wr"):
                v += units[i]          # ncu picks the unit per report (us / ms, Mbyte / Gbyte)
        parts.append(f"{n}={v}")
    print(" ".join(parts))
