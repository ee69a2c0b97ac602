#!/usr/bin/env python3
"""Summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`):
total / count / mean / share per kernel.   python profiles/tools/summarize_launches.py X.csv [header line ...]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
        k = row["Kernel Name"][:90]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    for h in sys.argv[2:]:
        print("# " + h)
    print("# total_us launches avg_us share kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:10.1f} {n:5d} {t / n:9.2f} {100 * t / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main()
