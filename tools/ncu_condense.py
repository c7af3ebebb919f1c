#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel table kept under profiles/ (one row per captured
kernel, the metrics DESIGN.md and profiles/*_summary.md quote).   ncu -i X.ncu-rep --page raw --csv | tools/ncu_condense.py"""
import csv
import sys

COLS = ["gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__block_size", "sm__cycles_active.avg", "smsp__cycles_active.avg"]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ix = {k: i for i, k in enumerate(hdr)}
w = csv.writer(sys.stdout)
cols = [c for c in COLS if c in ix]
w.writerow(["Kernel Name"] + cols)
w.writerow([""] + [units[ix[c]] for c in cols])
for r in rows[2:]:
    if len(r) == len(hdr):
        w.writerow([r[ix["Kernel Name"]][:90]] + [r[ix[c]] for c in cols])
