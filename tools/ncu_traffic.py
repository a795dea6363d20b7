#!/usr/bin/env python
"""Summarise .ncu-rep captures into profiles/: a text table of the key metrics per kernel
(profiles/<name>.summary.txt) and the per-launch DRAM traffic bench.py reports as `roofline.traffic`
(profiles/r2_traffic.json; BN_TRAFFIC_JSON names another file).

usage: ncu_traffic.py <config>:<phase>:<kernel-regex>:<report.ncu-rep> ...
"""
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEYS = [("duration_us", "gpu__time_duration.sum"), ("dram_read_bytes", "dram__bytes_read.sum"),
        ("dram_write_bytes", "dram__bytes_write.sum"),
        ("dram_pct_of_peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("registers", "launch__registers_per_thread"),
        ("threads_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
        ("fp64_pipe_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        ("warp_inst", "smsp__inst_executed.sum"),
        ("l2_pct", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed"),
        ("l1_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed")]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        rec = {"kernel": d["Kernel Name"], "grid": d.get("launch__grid_size"), "block": d.get("launch__block_size")}
        for name, key in KEYS:
            try:
                rec[name] = float(d[key]) * SCALE.get(u.get(key, ""), 1)
            except (KeyError, ValueError):
                rec[name] = None
        yield rec


def main():
    import os
    traffic_path = ROOT / "profiles" / os.environ.get("BN_TRAFFIC_JSON", "r2_traffic.json")
    traffic = json.loads(traffic_path.read_text()) if traffic_path.exists() else {}
    seen = set()
    for spec in sys.argv[1:]:
        config, phase, rx, rep = spec.split(":", 3)
        recs = list(rows_of(rep))
        name = Path(rep).stem
        if rep not in seen:
            seen.add(rep)
            with open(ROOT / "profiles" / (name + ".summary.txt"), "w") as f:
                f.write("# ncu --set full --clock-control none, one row per captured launch (%s)\n" % Path(rep).name)
                for r in recs:
                    f.write("%s  grid=%s block=%s\n" % (r["kernel"][:110], r["grid"], r["block"]))
                    f.write("    " + "  ".join("%s=%s" % (k, ("%.6g" % r[k]) if r[k] is not None else "n/a") for k, _ in KEYS) + "\n")
        hit = [r for r in recs if re.search(rx, r["kernel"])]
        if hit:
            b = sum(r["dram_read_bytes"] + r["dram_write_bytes"] for r in hit) / len(hit)
            traffic.setdefault(config, {})[phase] = {"kernel": hit[0]["kernel"].split("(")[0], "bytes_per_launch": b,
                                                     "launches_captured": len(hit), "source": "profiles/%s.summary.txt" % name}
    traffic_path.write_text(json.dumps(traffic, indent=1, sort_keys=True) + "\n")
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
