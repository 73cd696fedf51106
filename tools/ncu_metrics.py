"""Pick the metrics quoted in profiles/README.md out of an `ncu -i X.ncu-rep --page raw --csv` export:
    ncu -i X.ncu-rep --page raw --csv | python tools/ncu_metrics.py > profiles/X_metrics.csv"""
import csv
import sys

KEEP = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
idx = [hdr.index(k) for k in KEEP if k in hdr]
name = hdr.index("Kernel Name")
w = csv.writer(sys.stdout)
w.writerow(["Kernel Name"] + [hdr[i] for i in idx])
w.writerow([""] + [units[i] for i in idx])
for r in rows[2:]:
    if len(r) > max(idx):
        w.writerow([r[name]] + [r[i] for i in idx])
