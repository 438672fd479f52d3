"""Turn `ncu -i X.ncu-rep --page raw --csv` output into a compact metric,unit,value CSV for profiles/.

usage: ncu -i rep.ncu-rep --page raw --csv | python scripts/ncu_summary.py > profiles/NAME.csv
Keeps the metrics the roofline discussion uses (time, DRAM/L2 traffic, pipe utilisation, occupancy, stall reasons).
"""
import csv
import sys

KEEP = ("gpu__time_duration", "dram__bytes", "dram__throughput", "lts__t_bytes.sum", "lts__t_sector_hit_rate",
        "l1tex__t_sector_hit_rate", "l1tex__throughput", "sm__throughput", "sm__warps_active", "launch__",
        "sm__inst_executed_pipe_fp64", "sm__inst_executed_pipe_tensor_subpipe_dmma", "sm__pipe_tensor_subpipe_dmma",
        "sm__ops_path_tensor_src_fp64.sum", "smsp__inst_executed.sum", "sm__issue_active",
        "smsp__average_warps_issue_stalled", "smsp__warps_eligible", "smsp__issue_active.avg",
        "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_lsu",
        "smsp__pcsamp_warps_issue_stalled", "gpu__dram_throughput", "sm__cycles_elapsed.max")

rows = list(csv.reader(sys.stdin))
head, units = rows[0], rows[1]
out = csv.writer(sys.stdout)
for rec in rows[2:]:
    for i, name in enumerate(head):
        if name == "Kernel Name":
            out.writerow(["Kernel Name", "", rec[i]])
    for i, name in enumerate(head):
        if any(k in name for k in KEEP) and rec[i] not in ("", "0", "n/a"):
            if ".max." in name or ".min." in name or "peak_sustained" in name and "pct" not in name:
                continue
            out.writerow([name, units[i], rec[i]])
