#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel / grid.
usage: python scripts/launch_summary.py launches.csv  -> markdown table on stdout"""
import collections
import csv
import sys


def main():
    rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].replace("void ", "").split("(")[0]
        agg.setdefault((name[:70], r["Grid Size"], r["Block Size"]), []).append(float(r["Metric Value"]) / 1e3)
    tot = sum(sum(v) for v in agg.values())
    print("| launches | avg us | sum us | share | kernel | grid | block |")
    print("|---:|---:|---:|---:|---|---|---|")
    for (name, grid, block), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {len(v)} | {sum(v) / len(v):.1f} | {sum(v):.1f} | {100 * sum(v) / tot:.1f}% | `{name}` | {grid} | {block} |")
    print(f"\ntotal {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")


if __name__ == "__main__":
    main()
