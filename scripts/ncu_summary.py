#!/usr/bin/env python
"""Extract the judged columns of an `ncu --set full` capture into a small CSV (profiles/*_ncu_summary.csv).
Usage: scripts/ncu_summary.py capture.ncu-rep TAG > profiles/TAG_ncu_summary.csv"""
import csv
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
stalls = ["barrier", "branch_resolving", "dispatch_stall", "drain", "lg_throttle", "long_scoreboard", "math_pipe_throttle",
          "membar", "mio_throttle", "misc", "no_instruction", "not_selected", "selected", "short_scoreboard", "sleeping",
          "tex_throttle", "wait"]
names = cols + [f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" for s in stalls]
idx = [hdr.index(n) if n in hdr else None for n in names]
w = csv.writer(sys.stdout)
w.writerow(["capture"] + cols + [f"stall_{s}" for s in stalls])
w.writerow(["unit"] + [units[i] if i is not None else "" for i in idx])
for r in data:
    w.writerow([tag] + [r[i] if i is not None else "" for i in idx])
