"""Sharding of a query batch over the GPUs of one box.

Every (template, structure) pair is independent, so the shard unit is the structure:
templates are replicated on every rank, each rank scans its own structures with its own
engine context, and the only exchange is gathering the compacted hit tables -- there is
no data-path collective.  `partition` balances by atom count (longest-processing-time
first: one 58,870-atom structure costs as much as 25 median ones), `merge` re-numbers the
per-shard molecule indices back to the caller's input order.  `query_sharded` runs under
`torch.distributed` (one process per GPU): the per-rank hit tables meet on rank 0 in ONE
shared-memory segment (`HitExchange`), `ShardedHits` indexes them per structure without copying.
Inside ONE process, `Jess.query_many(..., devices=[...])` does the same with a thread per GPU.
"""
from __future__ import annotations

import os
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


class ShardedHits:
    """The hit tables of the shards plus, per structure of the caller's batch, WHERE its records
    are (table, row range): the merged view of a sharded run without copying a record.  Every
    table is sorted by (structure, template) in its shard's own numbering, shards own ascending
    input indices, so the index is two binary searches per shard."""

    def __init__(self, tables: Sequence[Optional[HitTable]], shards: Sequence[np.ndarray]):
        self.tables, self.shards = list(tables), [np.asarray(s, dtype=np.int64) for s in shards]
        n = int(sum(len(s) for s in self.shards))
        self.where = np.zeros(n, dtype=np.int32)
        self.lo = np.zeros(n, dtype=np.int64)
        self.hi = np.zeros(n, dtype=np.int64)
        for k, (t, idx) in enumerate(zip(self.tables, self.shards)):
            if len(idx) == 0:
                continue
            self.where[idx] = k
            if t is not None and t.count:
                local = np.arange(len(idx))
                self.lo[idx] = np.searchsorted(t.molecule, local, side="left")
                self.hi[idx] = np.searchsorted(t.molecule, local, side="right")

    @property
    def count(self) -> int:
        return int(sum(t.count for t in self.tables if t is not None))

    def of(self, m: int):
        """(table, first row, end row) of structure `m` (input numbering)."""
        return self.tables[self.where[m]], int(self.lo[m]), int(self.hi[m])

    def merged(self) -> HitTable:
        """One table in canonical order of the caller's numbering (copies every record)."""
        return merge(self.tables, self.shards)

    def checksum(self) -> int:
        """Order-independent 64-bit digest of the hit SET {(structure, template, atom tuple)} in
        the caller's numbering: equal digests across 1/2/4/8-GPU runs = identical merged sets."""
        total = np.uint64(0)
        P = np.uint64(0x9E3779B97F4A7C15)
        with np.errstate(over="ignore"):
            for t, idx in zip(self.tables, self.shards):
                if t is None or t.count == 0:
                    continue
                acc = idx[t.molecule].astype(np.uint64) + np.uint64(1)
                acc = acc * P + t.template.astype(np.uint64)
                for k in range(t.atoms.shape[1]):
                    acc = acc * P + (t.atoms[:, k].astype(np.int64) + 2).astype(np.uint64)
                acc ^= acc >> np.uint64(29)
                acc *= np.uint64(0xBF58476D1CE4E5B9)
                acc ^= acc >> np.uint64(32)
                total = total + acc.sum(dtype=np.uint64)
        return int(total)


class HitExchange:
    """Gathers the per-rank hit tables of one box on rank `dst` through ONE shared-memory segment
    (a file under /dev/shm mapped by every rank): each rank copies its seven arrays into its own
    region, rank `dst` reads all regions in place as numpy views.  No pickling, no socket, no
    collective on the data path; `torch.distributed` only carries the record counts (one small
    all_gather) and the barrier.  The views stay valid until the next `gather` on this object."""

    def __init__(self, rank: int, world: int, group=None, dst: int = 0):
        import torch.distributed as dist
        self.rank, self.world, self.group, self.dst = rank, world, group, dst
        token = [None]
        if rank == dst:
            import secrets
            token[0] = f"jb_hits_{os.getpid()}_{secrets.token_hex(4)}"
        dist.broadcast_object_list(token, src=dst, group=group)
        base = "/dev/shm" if os.path.isdir("/dev/shm") else __import__("tempfile").gettempdir()
        self.stem = os.path.join(base, token[0])
        self.capacity, self.generation, self.map, self.pool = 0, 0, None, None
        self.device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"

    @staticmethod
    def _layout(count: int, stride: int):
        sizes = (4 * count, 4 * count, 4 * count * stride, 8 * count, 72 * count, 24 * count, 24 * count)
        offs, pos = [], 0
        for b in sizes:
            offs.append(pos)
            pos += (b + 63) & ~63
        return offs, pos

    def _path(self) -> str:
        return f"{self.stem}_{self.generation}"

    def _ensure(self, need: int):
        import torch.distributed as dist
        if need <= self.capacity and self.map is not None:
            return
        old = self._path() if self.map is not None else None
        self.map = None
        self.generation += 1
        self.capacity = max(need + need // 4, 1 << 20)
        if self.rank == self.dst:
            if old and os.path.exists(old):
                os.unlink(old)
            with open(self._path(), "wb") as f:
                f.truncate(self.capacity)
        dist.barrier(group=self.group)                 # the segment exists before anyone maps it
        self.map = np.memmap(self._path(), dtype=np.uint8, mode="r+", shape=(self.capacity,))

    def gather(self, table: Optional[HitTable], stride: int):
        """Every rank calls this with its table (None = no structures); returns, on `dst`, the list
        of per-rank tables (views of the shared segment; rank dst's own table is passed through),
        elsewhere None."""
        import torch
        import torch.distributed as dist
        count = 0 if table is None else int(table.count)
        mine = torch.tensor([count], dtype=torch.int64, device=self.device)
        counts = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(counts, mine, group=self.group)
        counts = counts.cpu().numpy()
        bases, pos = [], 0
        for r in range(self.world):
            bases.append(pos)
            if r != self.dst:
                pos += self._layout(int(counts[r]), stride)[1]
        self._ensure(max(pos, 1))
        if self.rank != self.dst and count:
            offs, _ = self._layout(count, stride)
            atoms = table.atoms
            if atoms.shape[1] != stride:
                pad = np.full((count, stride), -1, dtype=np.int32)
                pad[:, :atoms.shape[1]] = atoms
                atoms = pad
            jobs = []
            for off, arr in zip(offs, (table.molecule, table.template, atoms, table.rmsd, table.rotation,
                                       table.centre0, table.centre1)):
                raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
                dst = self.map[bases[self.rank] + off: bases[self.rank] + off + raw.size]
                # large arrays in a few pieces: numpy copies without the GIL, one thread saturates ~10 GB/s only
                pieces = max(1, min(8, raw.size >> 24))
                cuts = [raw.size * i // pieces for i in range(pieces + 1)]
                jobs += [(dst[a:b], raw[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
            if len(jobs) > 7:
                if self.pool is None:
                    from concurrent.futures import ThreadPoolExecutor
                    self.pool = ThreadPoolExecutor(4)
                list(self.pool.map(lambda j: np.copyto(j[0], j[1]), jobs))
            else:
                for d, r in jobs:
                    np.copyto(d, r)
        dist.barrier(group=self.group)                 # all regions written
        if self.rank != self.dst:
            return None
        out = []
        for r in range(self.world):
            n = int(counts[r])
            if r == self.dst:
                out.append(table)
                continue
            if n == 0:
                out.append(None)
                continue
            offs, _ = self._layout(n, stride)
            b = bases[r]
            view = lambda k, dt, shape: self.map[b + offs[k]: b + offs[k] + int(np.prod(shape)) * np.dtype(dt).itemsize].view(dt).reshape(shape)  # noqa: E731
            out.append(HitTable.from_arrays(view(0, np.int32, (n,)), view(1, np.int32, (n,)), view(2, np.int32, (n, stride)),
                                            view(3, np.float64, (n,)), view(4, np.float64, (n, 3, 3)),
                                            view(5, np.float64, (n, 3)), view(6, np.float64, (n, 3))))
        return out

    def close(self):
        self.map = None
        if self.pool is not None:
            self.pool.shutdown()
            self.pool = None
        if self.rank == self.dst:
            try:
                os.unlink(self._path())
            except OSError:
                pass


def query_sharded(run_shard, molecules: Sequence, rank: int, world: int, group=None, dst: int = 0, exchange=None,
                  stride: Optional[int] = None):
    """Scan `molecules` with `world` ranks.  `run_shard(list_of_molecules) -> HitTable` is
    the per-rank engine call (e.g. a closure over `_capi.query_host`).  Every rank must
    call this with the same `molecules`; rank `dst` gets a `ShardedHits` (per-rank tables in
    shared memory + the per-structure index; `.merged()` builds the single canonical table),
    the others None.  `stride`: the template set's largest size (needed when a rank may have no
    hit table at all); `exchange`: a `HitExchange` to re-use across calls."""
    shards = partition([len(m) for m in molecules], world)
    mine = shards[rank]
    table = run_shard([molecules[int(i)] for i in mine]) if len(mine) else None
    if world == 1:
        return ShardedHits([table], shards)
    import torch
    import torch.distributed as dist
    own = exchange is None
    if own:
        exchange = HitExchange(rank, world, group, dst)
    try:
        if stride is None:                               # agree on the widest table
            dev = exchange.device
            s = torch.tensor([0 if table is None else int(table.stride)], dtype=torch.int64, device=dev)
            dist.all_reduce(s, op=dist.ReduceOp.MAX, group=group)
            stride = max(int(s.item()), 1)
        tables = exchange.gather(table, stride)
        if rank != dst:
            return None
        result = ShardedHits(tables, shards)
        if own:                                          # the views die with the segment: detach them
            result.tables = [None if t is None else HitTable.from_arrays(*(np.array(a) for a in t.arrays()))
                             for t in result.tables]
        return result
    finally:
        if own:
            exchange.close()
