#!/usr/bin/env python3
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: count, total, share.

    python profiles/launch_summary.py gpurun_out/launches.csv [--only-ours] > profiles/rNN_launch_summary.txt
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    only = "--only-ours" in sys.argv
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        name = row["Kernel Name"]
        if only and "b200id" not in name:
            continue
        agg[name.split("(")[0][-70:]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {sum(len(v) for v in agg.values())} launches, {tot / 1e3:.3f} ms total (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:72s} n={len(v):5d} total={sum(v) / 1e3:9.3f} ms mean={sum(v) / len(v):9.1f} us share={sum(v) / tot * 100:5.1f}%")


if __name__ == "__main__":
    main()
