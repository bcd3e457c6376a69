"""`Jess`, `Query`, `Hit`: the drop-in query API over the CUDA engine.

Mirrors ``src/pyjess/_jess.pyx`` (Jess :2624-2869, Query :2090-2283, Hit :2286-2621).
Where the reference pulls one tuple at a time out of a depth-first scanner
(`Query.__next__` -> ``JessQuery_next``), this module runs the whole
(templates x structures) batch on the GPU in ONE engine call and then iterates the
compacted hit records lazily:

* `Jess.query(molecule, ...)` is the single-structure special case;
* `Jess.query_many(molecules, ...)` is the batch entry point: one
  launch sequence for all structures, one `Query` per structure.

Iteration order: templates in REVERSE insertion order like the reference
(eff:Jess.c:109-117), hits of one template in the order the reference's depth-first
scanner yields them: the engine reconstructs it on the device from the candidate-set
positions and k-d tree leaf ranks (csrc/order.cuh, `hit_order="reference"`).  That
makes the two order-dependent features exact: the tie rule of ``best_match`` and
``max_candidates`` (`Query.__next__` _jess.pyx:2197-2283, replayed by
`_replay_query` over the ordered stream of ALL tuples).  There is no CPU fallback.
"""
from __future__ import annotations

import io
import math
import threading
import warnings
from typing import Iterable, List, Optional, Sequence

import numpy as np

from ._model import Atom, Mat4, Molecule, Template, _format_atom, _transform_points
from ._tables import JB_MAX_TEMPLATE_ATOMS, HitTable, TemplateTable, query_params

__all__ = ["Jess", "Query", "Hit"]


class Hit:
    """One (template, atom tuple) match with its superposition (_jess.pyx:2286-2621).

    A hit produced by a query is a lazy VIEW of one record of the engine's hit table: the
    atom records, the rotation and the centroids are only cut out of the table when they are
    first read, so iterating a query costs ~1 us per hit instead of a handful of numpy
    copies."""

    __slots__ = ("rmsd", "_rot", "_cen", "_at", "_template", "_molecule", "_table", "_row")

    @property
    def _rotation(self):
        if self._rot is None:
            self._rot = self._table.rotation[self._row].copy()
        return self._rot

    @property
    def _centre(self):
        if self._cen is None:
            self._cen = np.stack([self._table.centre0[self._row], self._table.centre1[self._row]])
        return self._cen

    @property
    def _atoms(self):
        if self._at is None:
            n = len(self._template)
            self._at = self._molecule._atoms[self._table.atoms[self._row, :n]]      # fancy index: a copy
        return self._at

    def __getstate__(self):
        return {
            "rotation": [float(v) for v in self._rotation.reshape(-1)],
            "centre": [[float(v) for v in row] for row in self._centre],
            "atoms": self.atoms(transform=False),
            "rmsd": self.rmsd,
            "template": self.template(transform=False),
            "molecule": self.molecule(transform=False),
        }

    def __setstate__(self, state):
        self.rmsd = state["rmsd"]
        self._template = state["template"]
        self._molecule = state["molecule"]
        self._table, self._row = None, -1
        self._rot = np.asarray(state["rotation"], dtype=np.float64).reshape(3, 3)
        self._cen = np.asarray(state["centre"], dtype=np.float64).reshape(2, 3)
        atoms = state["atoms"]
        if len(atoms) != len(self._template):
            raise ValueError(f"unexpected number of atoms: {len(atoms)!r} (expected {len(self._template)!r})")
        self._at = np.concatenate([a._rec for a in atoms]) if atoms else np.zeros(0, dtype=self._molecule._atoms.dtype)

    def __reduce__(self):
        return _rebuild_hit, (self.__getstate__(),)

    # T(c1) . R . T(-c0)  (_jess.pyx:2343-2377)
    @property
    def transformation(self) -> Mat4:
        mc, mr, mv = np.eye(4), np.eye(4), np.eye(4)
        mc[:3, 3] = -self._centre[0]
        mr[:3, :3] = self._rotation
        mv[:3, 3] = self._centre[1]
        return Mat4(mv) @ Mat4(mr) @ Mat4(mc)

    @property
    def inverse_transformation(self) -> Mat4:
        mc, mr, mv = np.eye(4), np.eye(4), np.eye(4)
        mv[:3, 3] = -self._centre[1]
        mr[:3, :3] = self._rotation.T
        mc[:3, 3] = self._centre[0]
        return Mat4(mc) @ Mat4(mr) @ Mat4(mv)

    @property
    def determinant(self) -> float:
        p = self._rotation.reshape(-1)
        det = 0.0
        det += p[0] * (p[4] * p[8] - p[5] * p[7])
        det -= p[1] * (p[3] * p[8] - p[5] * p[6])
        det += p[2] * (p[3] * p[7] - p[4] * p[6])
        return float(det)

    @property
    def log_evalue(self) -> float:
        return self._template._log_evalue(self.rmsd, len(self._molecule))

    @property
    def evalue(self) -> float:
        return math.exp(self.log_evalue)

    def _transformed(self) -> np.ndarray:
        """Matched atoms in the template frame: read from the device-computed table when the query
        asked for it (`coordinates=True`, JB_QUERY_COORDS), else transformed here."""
        t = self._table
        if t is not None and getattr(t, "coords", None) is not None:
            return t.coords[self._row, :len(self._template)]
        return _transform_points(self._atoms["x"].copy(), np.asarray(self.transformation))

    def atoms(self, transform: bool = True) -> List[Atom]:
        recs = self._atoms
        if transform:
            recs = recs.copy()
            recs["x"] = self._transformed()
        return [Atom._view(recs[k:k + 1], owner=self) for k in range(len(recs))]

    def molecule(self, transform: bool = False) -> Molecule:
        if not transform:
            return self._molecule
        return self._molecule.transform(self.transformation)

    def template(self, transform: bool = False) -> Template:
        if not transform:
            return self._template
        return self._template.transform(self.inverse_transformation)

    def dumps(self, format: str = "pdb", transform: bool = True) -> str:
        buf = io.StringIO()
        self.dump(buf, format=format, transform=transform)
        return buf.getvalue()

    def dump(self, file, format: str = "pdb", transform: bool = True) -> None:
        """Same text as the reference / original Jess (_jess.pyx:2541-2621)."""
        if self._template.id is None:
            raise RuntimeError("cannot dump `Hit` where `self.template.id` is `None`")
        if self._molecule.id is None:
            raise RuntimeError("cannot dump `Hit` where `self.molecule.id` is `None`")
        file.write("REMARK ")
        file.write(self._molecule.id)
        file.write(f" {self.rmsd:5.3f} ")
        file.write(self._template.id)
        file.write(f" Det={self.determinant:4,.1f} log(E)~ {self.log_evalue:4.2f}\n")
        from . import _capi
        xyz = np.ascontiguousarray(self._transformed()) if transform else None
        file.write(_capi.format_atoms_pdb(self._atoms, xyz))          # the per-atom sprintf loop, natively
        file.write("ENDMDL\n")


def _rebuild_hit(state):
    hit = Hit.__new__(Hit)
    hit.__setstate__(state)
    return hit


def _iteration_rows(molecule: np.ndarray, template: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """`rows` (records of a table sorted by (structure, template)) in iteration order: structures
    ascending, the (structure, template) RUNS of one structure in DESCENDING template order
    (reverse insertion, eff:Jess.c:109-117), records inside a run untouched.  No sort: the runs of
    a structure are contiguous and ascending, so they are simply mirrored inside their group."""
    if rows is None:                                         # every record of the table
        n = len(molecule)
        mol, tpl = molecule, template
    else:
        n = len(rows)
        mol, tpl = molecule[rows], template[rows]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    head = np.ones(n, dtype=bool)
    head[1:] = (mol[1:] != mol[:-1]) | (tpl[1:] != tpl[:-1])
    start = np.nonzero(head)[0]
    length = np.diff(np.append(start, n))
    run_mol = mol[start]
    group_head = np.ones(len(start), dtype=bool)
    group_head[1:] = run_mol[1:] != run_mol[:-1]
    first = np.nonzero(group_head)[0]                         # first run of every structure
    last = np.append(first[1:], len(start)) - 1
    group = np.cumsum(group_head) - 1
    perm = first[group] + last[group] - np.arange(len(start))  # position i takes the mirrored run
    start, length = start[perm], length[perm]
    offs = np.cumsum(length) - length
    order = np.repeat(start - offs, length) + np.arange(n)
    return order if rows is None else rows[order]


def _best_of_runs(molecule: np.ndarray, template: np.ndarray, rmsd: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """best_match on the host (`Query.__next__`, _jess.pyx:2238,2275-2283): of every (structure,
    template) run of `rows` the record with the smallest RMSD, the FIRST of equal ones."""
    if len(rows) == 0:
        return rows
    mol, tpl, r = molecule[rows], template[rows], rmsd[rows]
    head = np.ones(len(rows), dtype=bool)
    head[1:] = (mol[1:] != mol[:-1]) | (tpl[1:] != tpl[:-1])
    start = np.nonzero(head)[0]
    run = np.cumsum(head) - 1
    best = np.minimum.reduceat(r, start)
    pos = np.where(r == best[run], np.arange(len(rows)), len(rows))
    return rows[np.minimum.reduceat(pos, start)]


def _resolve_devices(devices) -> List[int]:
    """`devices` of `Jess.query_many`: None -> $JB_DEVICES ("all" or "0,1,..") or the process's
    default device ($JB_DEVICE / $LOCAL_RANK / 0); "all"; an int; a sequence of ints."""
    import os
    from . import _capi
    if devices is None:
        devices = os.environ.get("JB_DEVICES") or None
        if devices is None:
            return [int(os.environ.get("JB_DEVICE", os.environ.get("LOCAL_RANK", "0")))]
    if isinstance(devices, str):
        if devices.strip().lower() == "all":
            return list(range(max(1, _capi.device_count())))
        devices = [int(d) for d in devices.split(",") if d.strip()]
    elif isinstance(devices, int):
        devices = [devices]
    devices = [int(d) for d in devices]
    if not devices or len(set(devices)) != len(devices):
        raise ValueError(f"invalid value for `devices`: {devices!r}")
    return devices


class _BatchRun:
    """Shared, lazily executed engine work of one `query` / `query_many`.

    The batch is cut into one shard of structures per device (`sharding.partition`: LPT by atom
    count).  Every device is driven by its own host thread and engine context
    (`jb_query_records`: the engine packs the atom records slice by slice while it uploads and
    searches the previous slice).  Each shard yields its own hit table (page-locked memory of the
    engine, zero-copy): nothing is merged or copied, a `Query` only remembers which table and
    which rows belong to its structure -- and puts them into iteration order when it is first read."""

    __slots__ = ("jess", "molecules", "args", "all_tuples", "devices", "hit_order", "coords", "tables", "stats",
                 "item_stats", "_slices", "_where", "_lo", "_hi", "_keep", "done", "_lock", "_error", "_served")

    def __init__(self, jess, molecules, args, all_tuples, devices=None, hit_order="reference", coords=False):
        self.jess = jess
        self.molecules = molecules
        self.args = args                  # rmsd, distance, max_dynamic, ignore_chain, reorder, best_match
        self.all_tuples = all_tuples      # stream of every tuple (max_candidates replay) instead of hits
        self.devices = devices
        self.hit_order = hit_order        # "reference" | "canonical"
        self.coords = coords              # JB_QUERY_COORDS: matched atoms in the template frame, from the device
        self.tables: List[HitTable] = []  # one per work item
        self.stats = None
        self.item_stats = []
        self._slices = None               # per structure: rows of ITS table in iteration order
        self._where = None                # per structure: index into `tables`
        self._lo = self._hi = self._keep = None   # per structure: its block of the table (of tables[k][keep[k]])
        self.done = False                 # set once tables and slices are in place (or the run has failed)
        self._lock = threading.Lock()     # the Query objects of one batch may be consumed from several threads
        self._error = None
        self._served = False              # combined single-structure runs: results (or the error) are in place

    # single-table view (one work item): kept for callers and tests that predate sharded runs
    @property
    def table(self) -> Optional[HitTable]:
        return self.tables[0] if self.tables else None

    @table.setter
    def table(self, value):
        self.tables = [value]
        if self._where is None and self._slices is not None:
            self._where = [0] * len(self._slices)

    def table_of(self, m: int) -> HitTable:
        return self.tables[self._where[m] if self._where is not None else 0]

    def run(self):
        if not self.done:
            with self._lock:
                if not self.done:
                    try:
                        if len(self.molecules) == 1 and self.devices is None:
                            self._run_combined()
                        else:
                            self._run_locked()
                    except BaseException as exc:      # every later next() re-raises the real error
                        self._error = exc
                        raise
                    finally:
                        self.done = True
                    return
        if self._error is not None:
            raise self._error

    # -- single-structure queries arriving from several threads -----------------------------------
    # The reference's recipe for many structures is a pool of threads, each calling `jess.query(molecule, ...)`
    # (README.md:147-162): its C scanner releases the GIL and the calls run side by side.  Here they all
    # meet at ONE engine context, and a GPU query of one structure costs ~0.3 ms however small it is: the
    # calls would queue and the pool would run at the speed of a serial loop.  Instead the calls COMBINE:
    # a thread that finds the engine busy leaves its run in `jess._pending` and waits; when the engine is
    # free, ONE of the waiting threads (the leader of the round) takes everything that has piled up, runs
    # the runs with equal parameters as one batch -- one upload, one set of launches -- and hands every
    # run its block of the shared hit table.  A lone call takes the same path with a batch of one.
    def _combine_key(self):
        return (self.args, self.all_tuples, self.hit_order, self.coords)

    def _run_combined(self):
        jess = self.jess
        cv = jess._combine
        with cv:
            jess._pending.append(self)
            while True:
                if self._served:
                    if self._error is not None:
                        raise self._error
                    return
                if not jess._leader:
                    jess._leader = True
                    batch, jess._pending = jess._pending, []
                    break
                cv.wait()
        try:
            groups: dict = {}
            for r in batch:
                groups.setdefault(r._combine_key(), []).append(r)
            for group in groups.values():
                _serve_group(group)
        finally:
            with cv:
                for r in batch:
                    if not r._served:             # the leader failed outside an engine call
                        r._error = r._error or RuntimeError("combined query failed")
                        r._served = True
                jess._leader = False
                cv.notify_all()
        if self._error is not None:
            raise self._error

    def _plan(self):
        """[(device, ascending input indices of its shard)]."""
        mols = self.molecules
        devices = _resolve_devices(self.devices)
        if len(devices) > 1 and len(mols) >= 2 * len(devices):
            from .sharding import partition
            shards = partition([len(m) for m in mols], len(devices))
        else:
            devices, shards = devices[:1], [np.arange(len(mols), dtype=np.int64)]
        return [(dev, idx) for dev, idx in zip(devices, shards) if len(idx)]

    def _run_device(self, dev, idx, params, out):
        """Worker of one device: its shard through `jb_query_records` (the engine packs the atom
        records slice by slice while it uploads and searches the previous slice)."""
        from . import _capi  # raises if the CUDA engine is missing: no fallback
        ctx = _capi.default_context(dev)
        dev_tpl = self.jess._device_templates(ctx)
        mols = self.molecules
        sub = mols if len(idx) == len(mols) else [mols[int(i)] for i in idx]
        out.append((idx,) + _capi.query_records(ctx, dev_tpl, sub, params))

    def _run_locked(self, warn: bool = True):
        jess, mols = self.jess, self.molecules
        n_mol = len(mols)
        if len(jess) == 0 or not mols or all(len(m) == 0 for m in mols) or jess._table.max_atoms == 0:
            self._slices = [np.zeros(0, dtype=np.int64) for _ in mols]
            self._where = [0] * n_mol
            return
        rmsd, dist, maxdyn, ignore_chain, reorder, best_match = self.args
        single = np.asarray([len(tpl) == 1 for tpl in jess._templates], dtype=bool)
        # single-atom templates: the engine reports every candidate, the reference never the
        # FIRST of its candidate list (Scanner_reuse pre-selects index 0 and Scanner_next
        # pre-increments, eff:Scanner.c:176-181,212-227) -- in reference order that is the first
        # record of the (structure, template) run.  best_match must choose AFTER that record is
        # gone (all single-atom tuples have RMSD 0: the engine would keep exactly the one the
        # reference never reports), so with such templates the selection moves to the host.
        host_best = bool(best_match) and not self.all_tuples and bool(single.any())
        hit_order = self.hit_order
        if single.any() or self.all_tuples or best_match:
            hit_order = "reference"          # first-candidate drop, max_candidates and the best_match tie rule need it
        # the engine searches in its own scan order; the ORDER of the records is the
        # reference's for the `reorder` flag the caller passed
        params = query_params(rmsd, dist, maxdyn, ignore_chain, 2,
                              best_match and not self.all_tuples and not host_best,
                              hit_order="canonical" if hit_order == "canonical"
                              else ("reference" if reorder else "reference-noreorder"),
                              all_tuples=self.all_tuples, coords=self.coords)
        plan = self._plan()
        results = [[] for _ in plan]
        if len(plan) == 1:
            self._run_device(plan[0][0], plan[0][1], params, results[0])
        else:
            errors = []

            def work(i):
                try:
                    self._run_device(plan[i][0], plan[i][1], params, results[i])
                except BaseException as exc:
                    errors.append(exc)

            threads = [threading.Thread(target=work, args=(i,)) for i in range(len(plan))]
            for th in threads:
                th.start()
            for th in threads:
                th.join()
            if errors:
                raise errors[0]
        self._slices = [None] * n_mol
        where = np.zeros(n_mol, dtype=np.int64)
        self._lo = np.zeros(n_mol, dtype=np.int64)
        self._hi = np.zeros(n_mol, dtype=np.int64)
        self._keep = []
        total: dict = {}
        for items in results:
            for idx, t, st in items:
                k = len(self.tables)
                self.tables.append(t)
                self.item_stats.append(st)
                for key, v in st.items():
                    total[key] = total.get(key, 0) + v
                if t.count and len(idx) != n_mol:
                    t.molecule[:] = idx[t.molecule]            # the table speaks the caller's numbering
                keep = None
                if single.any() and t.count:
                    first = np.ones(t.count, dtype=bool)
                    first[1:] = (t.molecule[1:] != t.molecule[:-1]) | (t.template[1:] != t.template[:-1])
                    keep = np.nonzero(~(first & single[t.template]))[0]
                    if host_best:
                        keep = _best_of_runs(t.molecule, t.template, t.rmsd, keep)
                # The table is sorted by (structure, template): the records of a structure are one block of it
                # (of `keep`, when records were dropped).  Only the block boundaries are computed here -- two
                # binary searches per structure; a block is put into iteration order when its Query first asks.
                tm = t.molecule[:t.count] if keep is None else t.molecule[keep]
                where[idx] = k
                if len(tm):
                    self._lo[idx] = np.searchsorted(tm, idx, side="left")
                    self._hi[idx] = np.searchsorted(tm, idx, side="right")
                self._keep.append(keep)
        self._where = where.tolist()
        self.stats = total
        if warn and total.get("nan_dropped"):
            warnings.warn("Jess returned a superposition matrix with NaN values", UserWarning, stacklevel=5)

    def indices(self, m: int) -> np.ndarray:
        """Rows (of `table_of(m)`) holding the hits of structure `m`, in iteration order: templates DESCENDING
        (reverse insertion, eff:Jess.c:109-117), engine order within a template."""
        self.run()
        rows = self._slices[m]
        if rows is None:
            k = self._where[m]
            t, keep = self.tables[k], self._keep[k]
            block = np.arange(self._lo[m], self._hi[m], dtype=np.int64)
            if keep is not None:
                block = keep[block]
            rows = self._slices[m] = _iteration_rows(t.molecule, t.template, block)
        return rows


def _serve_group(group: List[_BatchRun]) -> None:
    """Run single-structure runs with equal parameters as ONE engine batch and give each its block of the shared
    hit table.  A batch whose engine run dropped NaN superpositions is re-run structure by structure, so that the
    reference's warning reaches the caller whose structure it concerns."""
    first = group[0]
    if len(group) > 1:
        big = _BatchRun(first.jess, [r.molecules[0] for r in group], first.args, first.all_tuples, None,
                        first.hit_order, first.coords)
        try:
            big._run_locked(warn=False)
        except BaseException:
            big = None                            # e.g. out of memory: every run gets its own attempt (and its own error)
        if big is not None and not (big.stats or {}).get("nan_dropped"):
            for j, r in enumerate(group):
                r.tables, r.item_stats, r.stats = big.tables, big.item_stats, big.stats
                r._keep = big._keep if big._keep is not None else []
                r._where = [big._where[j]] if big._where else [0]
                r._slices = [big._slices[j]] if big._lo is None else [None]
                if big._lo is not None:
                    r._lo, r._hi = big._lo[j:j + 1], big._hi[j:j + 1]
                r._served = True
            return
    for r in group:
        try:
            r._run_locked()
        except BaseException as exc:
            r._error = exc
        r._served = True


class Query:
    """Iterator over the hits of one structure (_jess.pyx:2090-2283)."""

    # (a `query_many` over 1,000 structures creates 1,000 of these: no per-object dict, a third of the allocations)
    __slots__ = ("_run", "_index", "_pos", "_ignore_chain", "_max_candidates", "_jess", "_molecule", "_rmsd_threshold",
                 "_best_match", "_partial", "_candidates", "_skip_template", "__weakref__")

    def __init__(self, jess, molecule, rmsd_threshold, best_match, ignore_chain, max_candidates, run: _BatchRun,
                 index: int):
        self._run = run
        self._index = index
        self._pos = 0
        # (validated once by `query_many`, which also told the run about `max_candidates`: the property setters
        #  below serve later assignments -- a batch of 1,000 structures creates 1,000 of these objects)
        self._ignore_chain = ignore_chain
        self._max_candidates = max_candidates
        self._jess = jess
        self._molecule = molecule
        self._rmsd_threshold = float(rmsd_threshold)
        self._best_match = bool(best_match)
        # replay state of Query.__next__ (max_candidates mode)
        self._partial = False
        self._candidates = 0
        self._skip_template = None

    jess = property(lambda self: self._jess)
    molecule = property(lambda self: self._molecule)
    rmsd_threshold = property(lambda self: self._rmsd_threshold)
    best_match = property(lambda self: self._best_match)

    @property
    def ignore_chain(self):
        return self._ignore_chain

    @ignore_chain.setter
    def ignore_chain(self, value):
        if value not in (None, "residues", "atoms"):
            raise ValueError(f"invalid value for `ignore_chain`: {value!r}")
        # The reference reads the property at every `next()` (_jess.pyx:2215-2221); here the whole
        # batch is searched by ONE engine run on the first `next()`: a new value reaches the search
        # until then (single-structure queries), afterwards it cannot be honoured -- say so loudly.
        run = self._run
        if value != run.args[3]:
            with run._lock:
                if run.done or len(run.molecules) != 1:
                    raise RuntimeError("ignore_chain cannot be changed once the query has started "
                                       "(or for one query of a `query_many` batch)")
                run.args = run.args[:3] + (value,) + run.args[4:]
        self._ignore_chain = value

    @property
    def max_candidates(self):
        return self._max_candidates

    @max_candidates.setter
    def max_candidates(self, value):
        if value is not None and value < 0:
            raise ValueError(f"invalid value for `max_candidates` argument: {value!r}")
        if value is not None and not self._run.all_tuples:
            with self._run._lock:            # a run in flight on another thread finishes first
                if self._run.done:
                    raise RuntimeError("max_candidates cannot be enabled once the query has started")
                self._run.all_tuples = True  # lazily executed: switch the shared run to the tuple stream
        self._max_candidates = value

    def __iter__(self):
        return self

    def __next__(self) -> Hit:
        idx = self._run.indices(self._index)
        if self._run.all_tuples:
            return self._replay_next(idx)
        if self._pos >= len(idx):
            raise StopIteration
        i = int(idx[self._pos])
        self._pos += 1
        return self._make_hit(i)

    def _replay_next(self, idx: np.ndarray) -> Hit:
        """`Query.__next__` of the reference (_jess.pyx:2197-2283) replayed over the stream
        of ALL tuples in the reference's iteration order: RMSD filter, NaN drop,
        best_match, and the max_candidates counter with its quirks (the counter restarts
        on every call because the local `tpl` starts as NULL, :2203,2222; exceeding it
        abandons the rest of the template, JessQuery_nextTemplate eff:Jess.c:197-205)."""
        t = self._run.table_of(self._index)
        limit = -1 if self._max_candidates is None else int(self._max_candidates)
        tpl = None
        hit_i, hit_rmsd, hit_tpl = -1, math.inf, None
        while True:
            # _advance (:2169-2173)
            if self._partial:
                self._partial = False
            else:
                self._pos += 1
                if self._skip_template is not None:
                    while self._pos <= len(idx) and int(t.template[idx[self._pos - 1]]) == self._skip_template:
                        self._pos += 1
                    self._skip_template = None
            if self._pos > len(idx):
                self._pos = len(idx) + 1
                break
            i = int(idx[self._pos - 1])
            prev_tpl = tpl
            tpl = int(t.template[i])
            if hit_i >= 0 and hit_tpl != tpl:
                self._partial = True          # _rewind
                break
            rmsd = float(t.rmsd[i])
            if rmsd <= self._rmsd_threshold and rmsd < hit_rmsd:
                if np.isnan(t.rotation[i]).any():
                    warnings.warn("Jess returned a superposition matrix with NaN values", UserWarning, stacklevel=3)
                else:
                    hit_i, hit_rmsd, hit_tpl = i, rmsd, tpl
            if tpl != prev_tpl:
                self._candidates = 0
            else:
                self._candidates += 1
            if limit != -1 and self._candidates > limit:
                self._candidates = 0
                self._skip_template = tpl      # JessQuery_nextTemplate
            if hit_i >= 0 and not self._best_match:
                break
        if hit_i < 0:
            raise StopIteration
        return self._make_hit(hit_i)

    def _make_hit(self, i: int) -> Hit:
        t = self._run.table_of(self._index)
        hit = Hit.__new__(Hit)
        hit.rmsd = float(t.rmsd[i])
        hit._rot = hit._cen = hit._at = None
        hit._table, hit._row = t, i
        hit._template = self._jess._templates[t.template[i]]
        hit._molecule = self._molecule
        return hit

    # vectorised access for heavy consumers (no Hit objects)
    def hit_table(self):
        """(HitTable, row indices of this structure) of the underlying batch."""
        rows = self._run.indices(self._index)        # runs the batch on first use: fetch it before the table
        return self._run.table_of(self._index), rows

    @property
    def stats(self):
        self._run.run()
        return self._run.stats


class Jess:
    """A set of templates to query structures with (_jess.pyx:2624-2869)."""

    def __init__(self, templates: Iterable[Template] = ()):
        tpls = []
        for t in templates:
            if not isinstance(t, Template):
                raise TypeError(f"expected Template, found {type(t).__name__}")
            tpls.append(t)
        self._templates = tuple(tpls)
        # the engine works on its own packed copy: later edits of the Template
        # objects cannot affect it (the reference deep-copies, _jess.pyx:2703-2707)
        self._table = TemplateTable(self._templates)
        # The one documented limit of the GPU path (include/jess_b200.h: JB_MAX_TEMPLATE_ATOMS): the scan
        # order and spec list of a (structure, template) pair live in one 128-byte record.  The reference
        # has no limit (Template.__init__, _jess.pyx:1825-1912); M-CSA / EnzyMM catalytic-site templates
        # hold 3 atoms x 3-8 residues.  Refused here, loudly, rather than at the first next().
        if self._table.max_atoms > JB_MAX_TEMPLATE_ATOMS:
            big = next(i for i, t in enumerate(self._templates) if len(t) > JB_MAX_TEMPLATE_ATOMS)
            raise ValueError(f"template {big} has {len(self._templates[big])} atoms; the B200 engine supports at most "
                             f"{JB_MAX_TEMPLATE_ATOMS} atoms per template")
        self._device = {}
        self._device_lock = threading.Lock()
        # single-structure queries from several threads combine into one engine batch (_BatchRun._run_combined)
        self._combine = threading.Condition()
        self._pending: List[_BatchRun] = []
        self._leader = False

    def _device_templates(self, ctx):
        from . import _capi
        dev = self._device.get(id(ctx))
        if dev is None:
            with self._device_lock:      # several device workers / user threads may arrive at once
                dev = self._device.get(id(ctx))
                if dev is None:
                    dev = self._device[id(ctx)] = _capi.DeviceTemplates(ctx, self._table)
        return dev

    def __copy__(self):
        return self.copy()

    def __reduce__(self):
        return type(self), (self._templates,)

    def __eq__(self, other):
        if not isinstance(other, Jess):
            return NotImplemented
        return self._templates == other._templates

    def __hash__(self):
        return hash((Jess, self._templates))

    def __len__(self):
        return len(self._templates)

    def __bool__(self):
        return len(self._templates) > 0

    def __getitem__(self, index):
        if isinstance(index, slice):
            return type(self)(self._templates[index])
        n = len(self._templates)
        i = int(index)
        if i < 0:
            i += n
        if i < 0 or i >= n:
            raise IndexError(index)
        return self._templates[i]

    def __iter__(self):
        return iter(self._templates)

    def copy(self):
        return type(self)(self._templates)

    @staticmethod
    def _normalise_ignore_chain(ignore_chain):
        if ignore_chain is True:
            warnings.warn(
                "`ignore_chain` parameter expects string parameters to specificy the mode since PyJess v0.7.0. "
                "Use `ignore_chain='atoms'` instead of `ignore_chain=True`", DeprecationWarning, stacklevel=3)
            return "atoms"
        if ignore_chain is False:
            warnings.warn(
                "`ignore_chain` parameter expects string parameters to specificy the mode since PyJess v0.7.0. "
                "Use `ignore_chain=None` instead of `ignore_chain=False`", DeprecationWarning, stacklevel=3)
            return None
        return ignore_chain

    def query(self, molecule: Molecule, rmsd_threshold: float, distance_cutoff: float, max_dynamic_distance: float, *,
              max_candidates: Optional[int] = None, ignore_chain=None, best_match: bool = False,
              reorder: bool = True, hit_order: str = "reference", coordinates: bool = False) -> Query:
        """Scan one structure (_jess.pyx:2758-2869).  Lazy: the engine runs on the
        first ``next()``.  `hit_order`, `coordinates`: see `query_many`."""
        return self.query_many([molecule], rmsd_threshold, distance_cutoff, max_dynamic_distance,
                               max_candidates=max_candidates, ignore_chain=ignore_chain, best_match=best_match,
                               reorder=reorder, hit_order=hit_order, coordinates=coordinates)[0]

    def query_many(self, molecules: Sequence[Molecule], rmsd_threshold: float, distance_cutoff: float,
                   max_dynamic_distance: float, *, max_candidates: Optional[int] = None, ignore_chain=None,
                   best_match: bool = False, reorder: bool = True, hit_order: str = "reference",
                   coordinates: bool = False, devices=None) -> List[Query]:
        """Scan a batch of structures; returns one `Query` per structure, in input order.

        Beyond the reference's arguments (keyword only, defaults = reference behaviour):

        `hit_order`: ``"reference"`` -- the hits of one template come in the order the reference's
        depth-first scanner yields them (rebuilt on the device); ``"canonical"`` -- ascending atom
        tuples, which skips that ranking (the reference documents the order under ``reorder=True`` as
        an implementation detail).  `best_match`, `max_candidates` and single-atom templates always
        use the reference order, their results depend on it.
        `coordinates`: also fetch, per hit, the matched atoms transformed into the template frame,
        computed on the device (`Hit.atoms(transform=True)`, `Hit.dump` and `Query.hit_table()` read
        them instead of transforming atom by atom).
        `devices`: GPUs to spread the batch over -- ``None`` ($JB_DEVICES, else the process's default
        device), ``"all"``, or a list of device indices.  Structures are dealt to the devices by atom
        count (`sharding.partition`), one host thread and engine context per device; large shards
        run in slices so that packing overlaps the engine."""
        molecules = list(molecules)
        for m in molecules:
            if not isinstance(m, Molecule):
                raise TypeError(f"expected Molecule, found {type(m).__name__}")
        ignore_chain = self._normalise_ignore_chain(ignore_chain)
        if ignore_chain not in (None, "residues", "atoms"):
            raise ValueError(f"invalid value for `ignore_chain`: {ignore_chain!r}")
        if max_candidates is not None and max_candidates < 0:
            raise ValueError(f"invalid value for `max_candidates` argument: {max_candidates!r}")
        if hit_order not in ("reference", "canonical"):
            raise ValueError(f"invalid value for `hit_order`: {hit_order!r}")
        run = _BatchRun(self, molecules,
                        (rmsd_threshold, distance_cutoff, max_dynamic_distance, ignore_chain, bool(reorder),
                         bool(best_match)), all_tuples=max_candidates is not None, devices=devices,
                        hit_order=hit_order, coords=bool(coordinates))
        return [Query(self, m, rmsd_threshold, best_match, ignore_chain, max_candidates, run, i)
                for i, m in enumerate(molecules)]
