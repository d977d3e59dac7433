"""Transpose an `ncu -i X.ncu-rep --page raw --csv` export into the per-kernel summary kept under profiles/.

usage: ncu_summary.py <raw.csv> <out.csv>

Keeps the last captured launch of every distinct kernel (launches after the warm-up) and the metrics the roofline
discussion in DESIGN.md uses (DRAM bytes and throughput, durations, issue/pipe utilisation, occupancy limits,
LSU wavefronts, stall reasons)."""
import csv
import re
import sys

KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum|gpu__dram_throughput|gpu__time_duration\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$"
                  r"|l1tex__data_pipe_lsu_wavefronts|launch__(block_size|grid_size|occupancy_limit|registers_per_thread|shared_mem_per_block_static)"
                  r"|sm__inst_executed_pipe_xu\.avg\.pct|sm__inst_issued\.avg\.pct_of_peak_sustained_active|sm__pipe_(alu|fma|fmaheavy|fmalite)_cycles_active\.avg\.pct"
                  r"|sm__throughput\.avg\.pct|sm__warps_active\.avg\.pct|smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio"
                  r"|smsp__inst_executed\.sum$|smsp__thread_inst_executed_per_inst_executed\.ratio|lts__t_bytes\.sum$|lts__t_sector_hit_rate\.pct)")

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units, data = rows[hdr_i], rows[hdr_i + 1], rows[hdr_i + 2:]
ki = hdr.index("Kernel Name")
last = {}
for r in data:
    if len(r) == len(hdr):
        last[r[ki]] = r  # later launches overwrite earlier ones
cols = [i for i, h in enumerate(hdr) if KEEP.match(h)]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    short = lambda k: re.sub(r"\(.*", "", k).replace("hb::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")  # noqa: E731
    w.writerow(["metric", "unit"] + [short(k) for k in last])
    w.writerow(["Kernel Name", ""] + list(last))
    for i in cols:
        w.writerow([hdr[i], units[i]] + [last[k][i] for k in last])
print("kernels:", [re.sub(r"\(.*", "", k) for k in last], "metrics:", len(cols))
