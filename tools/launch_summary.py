#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (cold-cache, serialised times:
compare SHARES).  usage: launch_summary.py launches.csv [steps_in_capture]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v * 1e6 if unit == "s" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'ms/step':>9} {'launches':>9} {'us/launch':>10} {'share':>7}  kernel   (total {tot / 1e3 / steps:.2f} ms/step over {steps:g} steps)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{v[1] / 1e3 / steps:9.3f} {v[0] / steps:9.1f} {v[1] / v[0]:10.1f} {100 * v[1] / tot:6.1f}%  {k[:120]}")


if __name__ == "__main__":
    main()
