#!/usr/bin/env python
"""Selected metrics of an .ncu-rep as a small CSV (one column per captured launch) for profiles/.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.csv"""
import csv, subprocess, sys
KEYS = ("Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__occupancy_limit",
        "launch__shared_mem_per_block", "sm__throughput.avg.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "smsp__warp_issue_stalled", "lts__throughput.avg.pct", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_write.sum",
        "smsp__average_warp", "smsp__pcsamp_warps_issue_stalled")
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = [i for i, h in enumerate(hdr) if any(k in h for k in KEYS) and ".peak_sustained" not in h and "per_second" not in h.replace("dram__bytes", "")]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{k}" for k in range(len(data))])
    for i in keep:
        w.writerow([hdr[i], units[i]] + [r[i] for r in data])
print(out, len(keep), "metrics x", len(data), "launches")
