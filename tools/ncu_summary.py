#!/usr/bin/env python
"""One line of key metrics per kernel from an .ncu-rep (ncu --set full).  usage: ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

KEYS = [("dur_us", "gpu__time_duration.sum"), ("dramR_MB", "dram__bytes_read.sum"), ("dramW_MB", "dram__bytes_write.sum"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"),
        ("lanes", "smsp__thread_inst_executed_per_inst_executed.ratio"),
        ("fp64%", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        ("inst", "smsp__inst_executed.sum"),
        ("l2%", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed"),
        ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print(d["Kernel Name"][:60], "grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
    print("   ", "  ".join("%s=%s%s" % (n, d.get(k, "?"), ("" if n not in ("dur_us", "dramR_MB", "dramW_MB") else "[" + u.get(k, "") + "]")) for n, k in KEYS))
    if len(sys.argv) > 2:
        for k in hdr:
            if sys.argv[2] in k:
                print("      ", k, d[k], u[k])
