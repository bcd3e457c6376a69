"""Sharding of a query batch over the GPUs of one box.

Every (template, structure) pair is independent, so the shard unit is the structure:
templates are replicated on every rank, each rank scans its own structures with its own
engine context, and the only exchange is gathering the compacted hit tables -- there is
no data-path collective.  `partition` balances by atom count (longest-processing-time
first: one 58,870-atom structure costs as much as 25 median ones), `merge` re-numbers the
per-shard molecule indices back to the caller's input order.  `query_sharded` runs under
`torch.distributed` (one process per GPU, nccl or gloo for the object gather).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from ._tables import HitTable


def partition(atom_counts: Sequence[int], n_shards: int) -> List[np.ndarray]:
    """LPT: structures by descending atom count, each to the currently lightest shard.
    Returns, per shard, the ASCENDING input indices it owns (deterministic)."""
    if n_shards < 1:
        raise ValueError("n_shards must be >= 1")
    counts = np.asarray(atom_counts, dtype=np.int64)
    order = np.argsort(-counts, kind="stable")
    load = np.zeros(n_shards, dtype=np.int64)
    owner = np.zeros(len(counts), dtype=np.int64)
    for i in order:
        s = int(np.argmin(load))          # first lightest shard
        owner[i] = s
        load[s] += counts[i] + 1          # +1: empty structures still spread out
    return [np.nonzero(owner == s)[0] for s in range(n_shards)]


def merge(tables: Sequence[Optional[HitTable]], shards: Sequence[np.ndarray], stride: Optional[int] = None) -> HitTable:
    """Concatenate per-shard hit tables into ONE table in canonical order of the
    caller's numbering: (input molecule index, template, atom tuple)."""
    present = [t for t in tables if t is not None]
    if stride is None:
        stride = max((t.stride for t in present), default=1)
    cols = [[] for _ in range(7)]
    for table, idx in zip(tables, shards):
        if table is None or table.count == 0:
            continue
        mol, tpl, atoms, rmsd, rot, c0, c1 = table.arrays()
        if atoms.shape[1] < stride:       # pad to the common stride
            pad = np.full((len(atoms), stride - atoms.shape[1]), -1, dtype=np.int32)
            atoms = np.concatenate([atoms, pad], axis=1)
        parts = (np.asarray(idx, dtype=np.int64)[mol].astype(np.int32), tpl, atoms, rmsd, rot, c0, c1)
        for col, part in zip(cols, parts):
            col.append(part)
    if not cols[0]:
        return HitTable.from_arrays(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, stride), np.int32),
                                    np.zeros(0), np.zeros((0, 3, 3)), np.zeros((0, 3)), np.zeros((0, 3)))
    mol, tpl, atoms, rmsd, rot, c0, c1 = (np.concatenate(c) for c in cols)
    # every shard is already canonical; a stable sort by molecule keeps (template, tuple) order
    perm = np.argsort(mol, kind="stable")
    return HitTable.from_arrays(mol[perm], tpl[perm], atoms[perm], rmsd[perm], rot[perm], c0[perm], c1[perm])


def query_sharded(run_shard, molecules: Sequence, rank: int, world: int, group=None, dst: int = 0):
    """Scan `molecules` with `world` ranks.  `run_shard(list_of_molecules) -> HitTable` is
    the per-rank engine call (e.g. a closure over `_capi.query_host`).  Every rank must
    call this with the same `molecules`; rank `dst` gets the merged table, the others None."""
    shards = partition([len(m) for m in molecules], world)
    mine = shards[rank]
    table = run_shard([molecules[int(i)] for i in mine]) if len(mine) else None
    if world == 1:
        return merge([table], shards)
    import torch.distributed as dist
    payload = None if table is None else tuple(np.ascontiguousarray(a) for a in table.arrays())
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    tables = [None if p is None else HitTable.from_arrays(*p) for p in gathered]
    return merge(tables, shards)
