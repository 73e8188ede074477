#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters the roofline notes quote.
usage: python scripts/ncu_summary.py file.ncu-rep [row]"""
import csv, subprocess, sys
rep = sys.argv[1]
row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__cycles_active.avg", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
r = rows[2 + row]
for h, u, v in zip(hdr, units, r):
    if h in keys or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        try:
            if "issue_stalled" in h and float(v) < 0.2:
                continue
        except ValueError:
            pass
        print(f"{h} [{u}] = {v}")
