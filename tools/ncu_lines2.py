#!/usr/bin/env python3
"""Per-CUDA-source-line aggregation of an `ncu --page source --csv --print-source cuda,sass`
export: instructions, stall samples, global L1 tag requests / L2 sectors, shared wavefronts.

    python tools/ncu_lines2.py src.csv <table-index> [file-suffix] [top]"""
import csv
import sys
from collections import defaultdict

COLS = ["Instructions Executed", "# Samples", "L1 Tag Requests Global", "L2 Theoretical Sectors Global",
        "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive", "stall_long_sb", "stall_short_sb"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    which = int(sys.argv[2])
    want = sys.argv[3] if len(sys.argv) > 3 else "expand_lv.cuh"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
    mine = [i for i in starts if rows[i - 1][1].endswith(want)]
    s = mine[which]
    e = min([i for i in starts if i > s] + [len(rows) + 1]) - 1
    hdr = rows[s + 1]
    print(rows[s][1])
    ci = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: defaultdict(float))
    text = {}
    cur = None
    for r in rows[s + 2:e]:
        if r[0].strip():
            try:
                cur = int(r[0]); text[cur] = r[1].strip()
                continue
            except ValueError:
                pass
        if cur is None:
            continue
        for c in COLS:
            try:
                agg[cur][c] += float(r[ci[c]] or 0)
            except (ValueError, IndexError):
                pass
    tot = {c: sum(a[c] for a in agg.values()) or 1.0 for c in COLS}
    print("totals", {c: f"{tot[c]:.3e}" for c in COLS})
    key = sys.argv[5] if len(sys.argv) > 5 else "L1 Tag Requests Global"
    print(f"{'line':>5} " + " ".join(f"{c[:12]:>12}" for c in COLS))
    for ln in sorted(agg, key=lambda k: -agg[k][key])[:top]:
        print(f"{ln:5d} " + " ".join(f"{100 * agg[ln][c] / tot[c]:11.1f}%" for c in COLS) + "  " + text.get(ln, "")[:90])


if __name__ == "__main__":
    main()
