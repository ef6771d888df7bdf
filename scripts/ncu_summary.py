#!/usr/bin/env python
"""Text summary of an .ncu-rep (ncu --set full): one column per profiled launch, the metrics the roofline discussion in
DESIGN.md uses. Usage: python scripts/ncu_summary.py report.ncu-rep [header line ...] > profiles/rNN_ncu_<kernel>.txt"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {n: k for k, n in enumerate(head)}
    for line in sys.argv[2:]:
        print(line)
    print("source: ncu -i %s --page raw --csv (%d launches)\n" % (rep.split("/")[-1], len(data)))
    names = [r[col["Kernel Name"]].split("(")[0].replace("void ", "")[:40] for r in data]
    w = max(44, max(len(n) for n in names) + 2)
    print("%-72s %-16s %s" % ("Kernel Name", "", " | ".join(n.ljust(w - 3) for n in names)))
    for m in METRICS:
        if m in col:
            print("%-72s %-16s %s" % (m, units[col[m]], " | ".join(r[col[m]].ljust(w - 3) for r in data)))
    print("\nwarp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active), values > 0.05:")
    for n, k in sorted(col.items()):
        if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and "not_issued" not in n:
            vals = [float(r[k].replace(",", "") or 0) for r in data]
            if max(vals) > 0.05:
                print("%-89s %s" % (n, " | ".join(("%.2f" % v).ljust(w - 3) for v in vals)))


if __name__ == "__main__":
    main()
