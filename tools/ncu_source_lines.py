#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export per CUDA source line.

    ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > src.csv
    python tools/ncu_source_lines.py src.csv [kernel-index] [top]

Prints, for one captured launch, the source lines ranked by warp instructions executed
with their share of stall samples -- the evidence behind the "where do the issue slots
go" tables in profiles/."""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(path)))
    # the export concatenates one table per launch, each introduced by "File Path"/"Function Name" rows
    # (one table per launch AND source file: pick the tables of the file given by JB_SRC, default kernels.cuh)
    import os
    want = os.environ.get("JB_SRC", "kernels.cuh")
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
    mine = [i for i in starts if rows[i - 1][1].endswith(want)]
    s = mine[which]
    e = min([i for i in starts if i > s] + [len(rows) + 1]) - 1
    hdr = rows[s + 1]
    print("launch", which, rows[s][1], rows[s - 1][1])
    ci = {h: i for i, h in enumerate(hdr)}
    line_c, inst_c, samp_c = ci["Line No"], ci["Instructions Executed"], ci["# Samples"]
    src_c = 1
    inst, samp, text = defaultdict(float), defaultdict(float), {}
    for r in rows[s + 2:e]:
        if len(r) <= inst_c or not r[line_c].strip():
            continue
        try:
            ln = int(r[line_c])
        except ValueError:
            continue
        try:
            inst[ln] += float(r[inst_c] or 0)
            samp[ln] += float(r[samp_c] or 0)
        except ValueError:
            pass
        text.setdefault(ln, r[src_c])
    ti, ts = sum(inst.values()), sum(samp.values())
    print(f"total warp instructions {ti:.3e}, samples {ts:.0f}")
    for ln in sorted(inst, key=lambda k: -inst[k])[:top]:
        print(f"{ln:5d} {100 * inst[ln] / ti:5.1f}% inst {100 * samp[ln] / max(ts, 1):5.1f}% samp  {text[ln].strip()[:110]}")


if __name__ == "__main__":
    main()
