#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + per-source-line share of instructions / stall samples.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio")
print("kernel:", v[h.index("Kernel Name")] if "Kernel Name" in h else "?")
for a, b, c in zip(h, u, v):
    if a in want:
        print("  %-90s %-12s %s" % (a, b, c))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, agg = None, []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 7 and r[0].isdigit():
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:105], int(r[6]), int(r[7])))
        except ValueError:
            pass
ts, ti = sum(a[3] for a in agg) or 1, sum(a[4] for a in agg) or 1
print("per source line: share of warp instructions executed / of warp-stall samples (total %d samples, %d instructions)" % (ts, ti))
for a in sorted(agg, key=lambda x: -x[3])[:top]:
    print("  %-16s %4d %5.1f%% inst %5.1f%% smp  %s" % (a[0][:16], a[1], 100.0 * a[4] / ti, 100.0 * a[3] / ts, a[2]))
