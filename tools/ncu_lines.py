#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + hottest source lines (needs -lineinfo).  Usage: ncu_lines.py rep [N]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg"]
for r in rows[2:]:
    print("== kernel", r[hdr.index("Kernel Name")][:90] if "Kernel Name" in hdr else "")
    for k in keys:
        if k in hdr:
            print("  %-72s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None; lines = []
for r in rows:
    if r and r[0] == "Line No":
        h = r; continue
    if h and r and r[0].isdigit() and len(r) == len(h):
        lines.append(r)
if h:
    iS, iI = h.index("# Samples"), h.index("Instructions Executed")
    stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot_s = sum(float(r[iS]) for r in lines); tot_i = sum(float(r[iI]) for r in lines)
    agg = collections.Counter()
    for r in lines:
        for i in stall:
            agg[h[i]] += float(r[i])
    print("total samples %d, instructions %.3g; stalls:" % (tot_s, tot_i), ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(1, tot_s)) for k, v in agg.most_common(9)))
    print("top lines by samples:  line  samples%  inst%  top-stall  source")
    for r in sorted(lines, key=lambda r: -float(r[iS]))[:N]:
        st = max(stall, key=lambda i: float(r[i]))
        print("  %5s %6.2f%% %6.2f%%  %-14s %s" % (r[0], 100 * float(r[iS]) / tot_s, 100 * float(r[iI]) / tot_i, h[st][6:], r[1].strip()[:110]))
