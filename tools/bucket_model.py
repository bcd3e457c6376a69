"""CPU model for DESIGN 7 "next" item 1: how many candidate records a partial match has to box-test
under three layouts of ONE candidate list, on the synthetic structures of configs 3 and 4.

  x32      -- today (k_xindex): records sorted along x, 32 equal-width x buckets, scan the buckets
              that meet [cx - r, cx + r]
  x32+y    -- the same buckets, records sorted along y INSIDE a bucket, scan only the y range
              (cost: one binary search per touched bucket, counted as log2(bucket size) probes)
  g8x4x... -- a coarse gx x gy grid in (x, y), cells stored contiguously, scan the touched cells

A list is "atoms named N of residue type T" (what one distinct atom spec of a template selects);
centres are atoms of the same structure; r is the outer shell radius d + t.
No GPU, no engine: numpy over `pyjess_b200.synth` structures.
"""
import sys

import numpy as np

from pyjess_b200 import synth


def lists_of(mol):
    a = mol._atoms
    key = np.char.add(a["name"].astype("S4"), a["resName"].astype("S4"))
    xyz = a["x"].astype(np.float64)
    for k in np.unique(key):
        sel = xyz[key == k]
        if len(sel) >= 8:
            yield sel


def x32(pts, c, r, nb=32):
    lo, hi = pts[:, 0].min(), pts[:, 0].max()
    w = max(hi - lo, 1e-9) / nb
    b = np.minimum(((pts[:, 0] - lo) / w).astype(int), nb - 1)
    b0 = np.clip(np.floor((c[:, 0] - r - lo) / w), 0, nb - 1).astype(int)
    b1 = np.clip(np.floor((c[:, 0] + r - lo) / w), 0, nb - 1).astype(int)
    miss = (c[:, 0] + r < lo) | (c[:, 0] - r > hi)
    cum = np.concatenate([[0], np.cumsum(np.bincount(b, minlength=nb))])
    tests = np.where(miss, 0, cum[b1 + 1] - cum[b0])
    # second key: y range inside every touched bucket
    ty = np.zeros(len(c))
    probes = np.zeros(len(c))
    for k in range(nb):
        yk = np.sort(pts[b == k, 1])
        if not len(yk):
            continue
        touch = (~miss) & (b0 <= k) & (k <= b1)
        n = np.searchsorted(yk, c[:, 1] + r, "right") - np.searchsorted(yk, c[:, 1] - r, "left")
        ty += np.where(touch, n, 0)
        probes += np.where(touch, np.ceil(np.log2(len(yk) + 1)), 0)
    return tests, ty, probes


def grid(pts, c, r, gx, gy):
    lo, hi = pts[:, :2].min(0), pts[:, :2].max(0)
    w = np.maximum(hi - lo, 1e-9) / (gx, gy)
    cell = np.minimum(((pts[:, :2] - lo) / w).astype(int), (gx - 1, gy - 1))
    cnt = np.zeros((gx + 1, gy + 1))
    np.add.at(cnt, (cell[:, 0] + 1, cell[:, 1] + 1), 1)
    cum = cnt.cumsum(0).cumsum(1)
    a0 = np.clip(np.floor((c[:, :2] - r - lo) / w), 0, (gx - 1, gy - 1)).astype(int)
    a1 = np.clip(np.floor((c[:, :2] + r - lo) / w), 0, (gx - 1, gy - 1)).astype(int)
    miss = ((c[:, :2] + r < lo) | (c[:, :2] - r > hi)).any(1)
    t = cum[a1[:, 0] + 1, a1[:, 1] + 1] - cum[a0[:, 0], a1[:, 1] + 1] - cum[a1[:, 0] + 1, a0[:, 1]] + cum[a0[:, 0], a0[:, 1]]
    return np.where(miss, 0, t), np.where(miss, 0, (a1[:, 0] - a0[:, 0] + 1) * (a1[:, 1] - a0[:, 1] + 1))


def main():
    rng = np.random.default_rng(0)
    sets = {"cfg3 (300 residues)": synth.proteins(6, 300, seed=1000),
            "cfg4 (4 x 500 residues)": synth.multimers(3, 500, 4, seed=3000)}
    print("| structures | r (A) | mean list | x32 | x32+y tests (+probes) | grid 8x4 tests (cells) | grid 8x8 tests (cells) | in sphere |")
    print("|---|---|---|---|---|---|---|---|")
    for label, mols in sets.items():
        for r in (6.0, 10.0):
            acc = np.zeros(9)
            n = 0
            for mol in mols:
                xyz = mol._atoms["x"].astype(np.float64)
                c = xyz[rng.choice(len(xyz), 256, replace=False)]
                for pts in lists_of(mol):
                    t, ty, pr = x32(pts, c, r)
                    g84, c84 = grid(pts, c, r, 8, 4)
                    g88, c88 = grid(pts, c, r, 8, 8)
                    d2 = ((pts[None] - c[:, None]) ** 2).sum(-1)
                    acc += [len(pts), t.mean(), ty.mean(), pr.mean(), g84.mean(), c84.mean(), g88.mean(),
                            c88.mean(), (d2 <= r * r).sum(1).mean()]
                    n += 1
            m = acc / n
            print(f"| {label} | {r:.0f} | {m[0]:.0f} | {m[1]:.1f} | {m[2]:.1f} (+{m[3]:.1f}) | {m[4]:.1f} ({m[5]:.1f}) | "
                  f"{m[6]:.1f} ({m[7]:.1f}) | {m[8]:.2f} |")


if __name__ == "__main__":
    sys.exit(main())
