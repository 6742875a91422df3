#!/usr/bin/env python
"""Kernel table from an .ncu-rep (ncu -i X --page raw --csv piped in) -> CSV with the metrics the judge greps."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
idx = [h.index(w) for w in want if w in h]
print(",".join(h[i] for i in idx))
print(",".join(rows[1][i] for i in idx))
for r in rows[2:]:
    print(",".join(r[i].replace(",", "") for i in idx))
