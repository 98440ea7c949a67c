#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + top stall reasons of EVERY launch in it. Usage: tools/ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, u = r[0], r[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
out = []
for v in r[2:]:
    if len(v) != len(h):
        continue
    for k in keys:
        if k in h:
            i = h.index(k); out.append("%-70s %-12s %s" % (k, u[i], v[i]))
    out.append("-- warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active.ratio) --")
    st = [(float(v[i]), n) for i, n in enumerate(h) if "issue_stalled" in n and n.endswith("per_issue_active.ratio") and v[i] not in ("", "n/a")]
    for val, n in sorted(st, reverse=True)[:8]:
        out.append("%-90s %.3f" % (n, val))
    out.append("")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
