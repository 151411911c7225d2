#!/usr/bin/env python
"""Summarise an .ncu-rep (first kernel): duration, DRAM bytes, utilisations, stall reasons per issue."""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u = rows[0], rows[1]
for v in rows[2:]:
    d = dict(zip(h, v))
    print("==", d.get("Kernel Name"), "grid", d.get("launch__grid_size"), "regs", d.get("launch__registers_per_thread"))
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_issued.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
    for k in keys:
        if k in d: print("  %-90s %s %s" % (k, d[k], u[h.index(k)]))
    st = [(float(d[k]), k) for k in h if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", k)]
    for val, k in sorted(st, reverse=True):
        if val > 0.02: print("  stall %-40s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), val))
