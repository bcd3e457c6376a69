#!/usr/bin/env python3
"""How much of the join box (Join_computeBox: intersection of the cubes around the placed atoms) do tighter bounds
keep?  Monte-Carlo over template-shaped shell sets (placed atoms and the new atom inside a 7 A ball, shells d +- 1.5 A
as BASELINE config 3's thresholds give): volume fractions, relative to the box, of the box cut by the six face-diagonal
slabs, by three of them, of the intersection of the outer balls and of the shells themselves.  Behind DESIGN.md "Next" 1.

    python tools/box_polytope_model.py
"""
import numpy as np


def trial(rng, L, n_trials=300, npts=200000, t=1.5):
    tot = np.zeros(5)
    dirs = np.array([[1, 1, 0], [1, -1, 0], [0, 1, 1], [0, 1, -1], [1, 0, 1], [1, 0, -1]]) / np.sqrt(2)
    for _ in range(n_trials):
        P = rng.normal(size=(L + 1, 3))
        P /= np.linalg.norm(P, axis=1)[:, None]
        P *= rng.uniform(0, 1, (L + 1, 1)) ** (1 / 3) * 7
        c, new = P[:L], P[L]
        d = np.linalg.norm(c - new, axis=1)
        lo, hi = np.maximum(d - t, 0.5), d + t
        b0, b1 = (c - hi[:, None]).max(0), (c + hi[:, None]).min(0)
        if (b0 > b1).any():
            continue
        pts = rng.uniform(b0, b1, (npts, 3))
        vol = np.prod(b1 - b0)
        dist = np.linalg.norm(pts[:, None, :] - c[None], axis=2)
        true = ((dist >= lo) & (dist <= hi)).all(1)
        outer = (dist <= hi).all(1)
        pc = c @ dirs.T
        s0, s1 = (pc - hi[:, None]).max(0), (pc + hi[:, None]).min(0)
        pp = pts @ dirs.T
        diag6 = ((pp >= s0) & (pp <= s1)).all(1)
        k = [0, 2, 4]
        diag3 = ((pp[:, k] >= s0[k]) & (pp[:, k] <= s1[k])).all(1)
        tot += np.array([vol, vol * diag6.mean(), vol * diag3.mean(), vol * outer.mean(), vol * true.mean()])
    return tot


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    print("placed atoms | box + 6 diagonals | box + 3 diagonals | outer balls | shells")
    for L in (1, 2, 3):
        v = trial(rng, L)
        print(f"{L:12d} | {v[1] / v[0]:17.3f} | {v[2] / v[0]:17.3f} | {v[3] / v[0]:11.3f} | {v[4] / v[0]:6.3f}")
