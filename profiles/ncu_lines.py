#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + per-source-line / per-opcode instruction shares.
usage: python profiles/ncu_lines.py gpurun_out/prof.ncu-rep [frames_in_launch]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
PX = 524288.0 * frames  # warp-pixels per 4096x4096 frame

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, val = rows[0], rows[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-80s %s %s" % (w, val[i][:90], rows[1][i]))
if "smsp__inst_executed.sum" in hdr:
    print("warp instructions per input pixel: %.1f" % (float(val[hdr.index("smsp__inst_executed.sum")]) / PX))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
agg = defaultdict(lambda: [0, 0])
srcs, ops, cur, tot = {}, defaultdict(int), None, 0
for r in rows[hi + 1:]:
    if r[0] != "":
        cur = r[0]
        srcs.setdefault(cur, ",".join(r[1:-1])[:90])
    elif len(r) > 7 and r[2].startswith("0x"):
        try:
            n, s = int(r[7]), int(r[4])
        except ValueError:
            continue
        agg[cur][0] += n
        agg[cur][1] += s
        tot += n
        t = r[3].split()
        ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += n
print("\nper source line (share of executed warp instructions, stall samples):")
for ln, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print("%5s %6.1f%% %7d  %s" % (ln, 100.0 * n / tot, s, srcs[ln].strip()))
print("\nper opcode:")
for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:18]:
    print("%-12s %5.1f%%" % (op, 100.0 * n / tot))
