#!/usr/bin/env python3
"""The handful of `ncu --page details --csv` rows the docs quote, per profiled launch.   python profiles/tools/ncu_details_summary.py X.csv [header ...]"""
import csv
import sys

WANT = ["Duration", "DRAM Throughput", "Memory Throughput", "Compute (SM) Throughput", "Executed Ipc Active", "Issue Slots Busy", "No Eligible",
        "Active Warps Per Scheduler", "Eligible Warps Per Scheduler", "Warp Cycles Per Issued Instruction", "Avg. Active Threads Per Warp", "L1/TEX Hit Rate", "L2 Hit Rate",
        "Registers Per Thread", "Dynamic Shared Memory Per Block", "Static Shared Memory Per Block", "Theoretical Occupancy", "Achieved Occupancy",
        "Block Limit Registers", "Block Limit Shared Mem", "Block Limit Warps"]
rows = list(csv.DictReader(open(sys.argv[1])))
for h in sys.argv[2:]:
    print("# " + h)
seen = set()
for r in rows:
    key = r["ID"]
    if key not in seen:
        seen.add(key)
        print(f"  {r['Kernel Name'][:110]} {r['Grid Size']}x{r['Block Size']}")
    if r["Metric Name"] in WANT:
        print(f"    {r['Metric Name']:42s} {r['Metric Unit']:16s} {r['Metric Value']}")
