"""`k_kd_order_team` (pyjess_b200/csrc/order.cuh) replays the reference's k-d tree construction with two
re-arrangements of the thread-per-list kernel: the tree is built LEVEL BY LEVEL (lane t takes segment t of the level)
instead of through a stack, and the partition loop reads elements i + 1 and i + 2 BEFORE the swap of step i and writes
`idx[i] = idx[store]; idx[store] = v` instead of swapping.  This is a line-by-line Python model of that kernel; it must
produce the permutation of `_order_model.kd_leaf_order` -- the model pinned to the reference's golden hit order and to
the unmodified C library (tests/test_order_model.py) -- on random lists, with ties on every axis."""
import numpy as np
import pytest

from _order_model import kd_leaf_order


def _cmp(u, a, b, axis):
    c, d = u[a][axis], u[b][axis]
    return -1 if c < d else (1 if c > d else 0)


def _qselect_team(idx, s, n, k, u, axis):
    """kd_qselect_sm on idx[s:s+n]."""
    left, right = 0, n - 1
    while left != right:
        mid = (left + right + 1) // 2
        if (_cmp(u, idx[s + left], idx[s + mid], axis) > 0) != (_cmp(u, idx[s + left], idx[s + right], axis) > 0):
            pivot = left
        elif (_cmp(u, idx[s + mid], idx[s + left], axis) < 0) != (_cmp(u, idx[s + mid], idx[s + right], axis) < 0):
            pivot = mid
        else:
            pivot = right
        idx[s + right], idx[s + pivot] = idx[s + pivot], idx[s + right]
        store = left
        pv = u[idx[s + right]][axis]
        v, v1 = idx[s + left], idx[s + min(left + 1, right)]
        c, c1 = u[v][axis], u[v1][axis]
        for i in range(left, right):
            v2 = idx[s + min(i + 2, right)]                 # fetched BEFORE the writes of step i
            c2 = u[v2][axis]
            if not c > pv:
                idx[s + i] = idx[s + store]
                idx[s + store] = v
                store += 1
            v, c, v1, c1 = v1, c1, v2, c2
        idx[s + store], idx[s + right] = idx[s + right], idx[s + store]
        if store == k:
            return
        if k < store:
            right = store - 1
        else:
            left = store + 1


def kd_leaf_order_team(coords):
    n = len(coords)
    idx = list(range(n))
    u = [tuple(map(float, c)) for c in coords]
    cur, depth = ([(0, n)] if n >= 2 else []), 0
    while cur:
        nxt = []
        for s, m in cur:                                    # (one lane per segment on the device: the segments are disjoint)
            split = m // 2
            _qselect_team(idx, s, m, split, u, depth % 3)
            if split >= 2:
                nxt.append((s, split))
            if m - split >= 2:
                nxt.append((s + split, m - split))
        cur, depth = nxt, depth + 1
    return idx


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 31, 32, 33, 64, 100, 257, 700])
def test_level_by_level_replay_equals_the_recursive_model(n):
    rng = np.random.default_rng(n)
    for trial in range(6):
        if trial % 3 == 0:
            coords = rng.normal(scale=20.0, size=(n, 3))                     # all distinct
        elif trial % 3 == 1:
            coords = np.round(rng.normal(scale=3.0, size=(n, 3)))            # many ties on every axis
        else:
            coords = np.round(rng.normal(scale=20.0, size=(n, 3)), 3)        # PDB-like: three decimals, a few ties
            coords[rng.integers(0, n, max(1, n // 4))] = coords[0]           # exact duplicates of one atom
        assert kd_leaf_order_team(coords) == kd_leaf_order(coords)
