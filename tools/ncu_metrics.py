"""Print selected metrics from an `ncu --page raw --csv` dump:  python tools/ncu_metrics.py raw.csv [pattern...]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
pats = sys.argv[2:] or [
    r"^gpu__time_duration.sum$", r"^dram__bytes_(read|write).sum$", r"dram__throughput.avg.pct_of_peak_sustained_elapsed",
    r"^sm__throughput.avg.pct", r"sm__warps_active.avg.pct_of_peak_sustained_active", r"launch__registers_per_thread",
    r"launch__occupancy_limit", r"launch__waves", r"smsp__issue_active.avg.pct", r"smsp__average_warp.*stall",
    r"smsp__average_warps_issue_stalled_.*_per_issue_active", r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
    r"l1tex__t_sectors_pipe_lsu_mem_global_op_(st|ld).sum$", r"l1tex__t_requests_pipe_lsu_mem_global_op_(st|ld).sum$",
    r"lts__t_sectors_op_(read|write).sum$", r"lts__t_sector_hit_rate.pct", r"smsp__inst_executed.sum$",
    r"sm__cycles_elapsed.max", r"launch__grid_size", r"launch__block_size", r"dram__sectors_(read|write).sum",
]
kn = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
if kn is not None:
    print("kernels:", [r[kn][:50] for r in data])
for i, h in enumerate(hdr):
    if any(re.search(p, h) for p in pats):
        print(f"{h:90s} {units[i]:14s}", [r[i] for r in data])
