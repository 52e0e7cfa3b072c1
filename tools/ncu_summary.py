#!/usr/bin/env python
"""Reduce `ncu -i X.ncu-rep --page raw --csv` output to the per-kernel summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/r2c_colour.raw.csv [kernel-substring] > profiles/r2/colour.txt
"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    only = sys.argv[2] if len(sys.argv) > 2 else None
    hdr, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]
        if (only and only not in name) or name in seen:
            continue
        seen.add(name)
        print("Kernel Name  %s" % name[:160])
        for k in KEYS:
            if k in d:
                print("%-80s %s %s" % (k, d[k], units[hdr.index(k)]))
        try:
            t = float(d["gpu__time_duration.sum"].replace(",", ""))
            tu = units[hdr.index("gpu__time_duration.sum")]
            t *= {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(tu, 1e-9)
            rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
            mul = lambda k: {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[hdr.index(k)]]
            b = rd * mul("dram__bytes_read.sum") + wr * mul("dram__bytes_write.sum")
            print("%-80s %.1f GB/s (dram read + write bytes / duration)" % ("derived dram bandwidth", b / t / 1e9))
        except (KeyError, ValueError):
            pass
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(d[h])
                except ValueError:
                    continue
                if v >= 0.3:
                    print("%-80s %s" % ("stall " + h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], d[h]))
        print()


if __name__ == "__main__":
    main()
