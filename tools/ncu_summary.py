#!/usr/bin/env python
"""Headline counters of the first kernel in an .ncu-rep (raw page): duration, warp-instructions, issue utilisation, lanes,
registers, DRAM bytes, top stall reasons.  usage: python tools/ncu_summary.py REPORT.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines())); h = r[0]
for v in r[2:]:
    print("kernel:", v[h.index("Kernel Name")])
    for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
              "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum",
              "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]:
        if k in h: print(f"  {k:60s} {v[h.index(k)]} {r[1][h.index(k)]}")
    st = [(h[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v[i]))
          for i in range(len(h)) if "issue_stalled" in h[i] and h[i].endswith("_per_issue_active.ratio")]
    st.sort(key=lambda x: -x[1])
    print("  stalls (cycles per issued instruction):", ", ".join(f"{a}={b:.2f}" for a, b in st[:8]))
