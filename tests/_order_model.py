"""Pure-Python model of the reference's ITERATION ORDER (test infrastructure).

Restates, for small cases, the three reference pieces that fix the order in which
`Scanner_next` yields the tuples of one (structure, template) pair:

* `candidate_set_order`  TessTemplate_candidates (eff:TessTemplate.c:141-201) over the
  residue index (ResIndex_create, eff:ResIndex.c:18-27,58-61);
* `kd_leaf_order`        KdTreeNode_create (eff:KdTree.c:343-423) with qselect_r
  (patches/qselect.h.patch:60-119); KdTreeQuery_next (eff:KdTree.c:238-296) pops the
  RIGHT child first, so leaves are visited in reverse;
* `scan_order`           Scanner_reuse's argsort by (|C_k|, k) (eff:Scanner.c:88-99,160-174).

`tests/test_order_model.py` pins this model to the order of the reference's own hits in
tests/golden/reference_hits.json; the GPU tests then use it (and oracle/_ref) to check
csrc/order.cuh on inputs that have no golden file.
"""
import numpy as np

import oracle
from pyjess_b200._tables import MoleculeBatch, TemplateTable


def _upper(tok: bytes) -> bytes:
    return bytes(tok).upper()


def candidate_set_order(molecule, tatom, cand):
    """`cand` (atom indices, any order) in CandidateSet order."""
    atoms = molecule._atoms
    cand = np.sort(np.asarray(cand, dtype=np.int64))
    if not (-1 <= tatom._code <= 8):
        return [int(a) for a in cand]
    res = [_upper(atoms["resName"][a]).ljust(4, b"\0")[:4].rstrip(b"\0") for a in cand]
    out, done = [], set()
    for name in tatom._resnames:
        tok = _upper(name).ljust(4, b"\0")[:4].rstrip(b"\0")
        if tok in done:
            continue
        done.add(tok)
        sel = [int(a) for a, r in zip(cand, res) if r == tok]
        sel.sort(key=lambda a: (int(atoms["serial"][a]), a))
        out.extend(sel)
    return out


def _qselect(idx, lo, n, k, u, axis):
    """qselect_r on idx[lo:lo+n] (in place)."""
    def cmp(a, b):
        c, d = u[idx[lo + a]][axis], u[idx[lo + b]][axis]
        return -1 if c < d else (1 if c > d else 0)

    def swap(a, b):
        idx[lo + a], idx[lo + b] = idx[lo + b], idx[lo + a]

    left, right = 0, n - 1
    while left != right:
        mid = (left + right + 1) // 2
        if (cmp(left, mid) > 0) != (cmp(left, right) > 0):
            pivot = left
        elif (cmp(mid, left) < 0) != (cmp(mid, right) < 0):
            pivot = mid
        else:
            pivot = right
        swap(right, pivot)
        store = left
        for i in range(left, right):
            if cmp(i, right) <= 0:
                swap(i, store)
                store += 1
        swap(store, right)
        if store == k:
            return
        if k < store:
            right = store - 1
        else:
            left = store + 1


def kd_leaf_order(coords):
    """Left-to-right leaf order (indices into `coords`) of the reference's k-d tree."""
    n = len(coords)
    idx = list(range(n))
    u = [tuple(map(float, c)) for c in coords]
    stack = [(0, n, 0)]
    while stack:
        s, m, d = stack.pop()
        if m <= 1:
            continue
        split = m // 2
        _qselect(idx, s, m, split, u, d % 3)
        stack.append((s + split, m - split, d + 1))
        stack.append((s, split, d + 1))
    return idx


def scan_order(counts, reorder):
    n = len(counts)
    return sorted(range(n), key=lambda k: (counts[k], k)) if reorder else list(range(n))


def rank_tables(template, molecule):
    """Per template atom k: (position in the CandidateSet, reverse k-d leaf rank) per atom index."""
    tt, mb = TemplateTable([template]), MoleculeBatch([molecule])
    set_pos, kd_rank, counts = [], [], []
    xyz = molecule._atoms["x"]
    for k in range(len(template)):
        cand = oracle.candidates(tt, mb, 0, 0, k)
        order = candidate_set_order(molecule, template[k], cand)
        counts.append(len(order))
        set_pos.append({a: i for i, a in enumerate(order)})
        leaves = kd_leaf_order([xyz[a] for a in order])
        kd_rank.append({order[j]: len(order) - 1 - q for q, j in enumerate(leaves)})
    return set_pos, kd_rank, counts


def reference_sort_key(template, molecule, reorder):
    """key(atom tuple in template order) -> rank vector in the reference's scan order."""
    set_pos, kd_rank, counts = rank_tables(template, molecule)
    order = scan_order(counts, reorder)

    def key(atoms):
        return tuple((set_pos if l == 0 else kd_rank)[k][int(atoms[k])] for l, k in enumerate(order))
    return key
