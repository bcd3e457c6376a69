#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a
markdown table (kernel, launches, total ms, share).  Usage:
    python tools/summarise_launches.py gpurun_out/launches.csv"""
import collections
import csv
import sys


def summarise(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        a = agg.setdefault(row["Kernel Name"].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values()) or 1.0
    out = ["| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, a in agg.items():
        out.append(f"| `{k}` | {a[0]} | {a[1]:.3f} | {a[1] / total:.3f} |")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarise(sys.argv[1]))
