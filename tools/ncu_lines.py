#!/usr/bin/env python3
"""Instruction share per CUDA source line, in line order, of one kernel of an
`ncu --page source --csv --print-source cuda,sass` export.
    python tools/ncu_lines.py src.csv <kernel-index> [file-suffix] [min-pct]"""
import csv
import sys
from collections import defaultdict


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    which = int(sys.argv[2])
    want = sys.argv[3] if len(sys.argv) > 3 else "expand_grid.cuh"
    minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
    mine = [i for i in starts if rows[i - 1][1].endswith(want)]
    s = mine[which]
    e = min([i for i in starts if i > s] + [len(rows) + 1]) - 1
    hdr = rows[s + 1]
    ci = {h: i for i, h in enumerate(hdr)}
    print(rows[s][1])
    cols = ["Instructions Executed", "Thread Instructions Executed", "# Samples"]
    agg = defaultdict(lambda: defaultdict(float))
    text, cur = {}, None
    for r in rows[s + 2:e]:
        if r[0].strip():
            try:
                cur = int(r[0]); text[cur] = r[1].strip()
                continue
            except ValueError:
                pass
        if cur is None:
            continue
        for c in cols:
            try:
                agg[cur][c] += float(r[ci[c]] or 0)
            except (ValueError, IndexError, KeyError):
                pass
    tot = {c: sum(a[c] for a in agg.values()) or 1.0 for c in cols}
    print(f"{'line':>5} {'inst%':>6} {'lanes':>5} {'smpl%':>6}")
    for ln in sorted(agg):
        a = agg[ln]
        if 100 * a[cols[0]] / tot[cols[0]] >= minpct:
            print(f"{ln:5d} {100 * a[cols[0]] / tot[cols[0]]:6.2f} {a[cols[1]] / max(a[cols[0]], 1):5.1f} {100 * a[cols[2]] / tot[cols[2]]:6.2f}  {text.get(ln, '')[:120]}")


if __name__ == "__main__":
    main()
