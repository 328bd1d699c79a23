"""ncu -i X.ncu-rep --page raw --csv  ->  compact per-kernel summary CSV for profiles/ (run where ncu exists)"""
import csv, subprocess, sys
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum", "dram__sectors_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = [hdr.index(k) for k in KEEP if k in hdr]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
print("wrote", out, len(rows) - 2, "kernels")
