"""Flat host tables of the C ABI (``include/jess_b200.h``) as ctypes structs.

`TemplateTable` is what `Jess.__init__` hands to ``jb_templates_create``
(the stand-in for Jess_create/Jess_addTemplate, reference include/jess/jess.pxd:18-21);
`MoleculeBatch` is the structure-of-arrays image of a list of molecules handed to
``jb_molecules_upload`` (stand-in for Molecule_* / ResIndex_create).  Both are cut
out of the numpy storage of `_model` without per-atom Python work.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from ._model import Molecule, Template

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u32p = C.POINTER(C.c_uint32)
c_u8p = C.POINTER(C.c_uint8)
c_f64p = C.POINTER(C.c_double)

JB_MAX_TEMPLATE_ATOMS = 32
IGNORE_CHAIN = {None: 0, "residues": 1, "atoms": 2}


class jb_template_table(C.Structure):
    _fields_ = [
        ("n_templates", C.c_int32),
        ("atom_offset", c_i32p),
        ("match_mode", c_i32p),
        ("residue_number", c_i32p),
        ("chain_id2", c_u8p),
        ("position", c_f64p),
        ("distance_weight", c_f64p),
        ("name_offset", c_i32p),
        ("atom_names", c_u32p),
        ("resname_offset", c_i32p),
        ("residue_names", c_u32p),
        ("distance_offset", c_i64p),
        ("distances", c_f64p),
    ]


class jb_molecule_batch(C.Structure):
    _fields_ = [
        ("n_molecules", C.c_int32),
        ("atom_offset", c_i64p),
        ("x", c_f64p),
        ("y", c_f64p),
        ("z", c_f64p),
        ("name", c_u32p),
        ("residue_name", c_u32p),
        ("residue_number", c_i32p),
        ("chain_id2", c_u8p),
        ("serial", c_i32p),
    ]


class jb_query_params(C.Structure):
    _fields_ = [
        ("rmsd_threshold", C.c_double),
        ("distance_cutoff", C.c_double),
        ("max_dynamic_distance", C.c_double),
        ("ignore_chain", C.c_int32),
        ("reorder", C.c_int32),
        ("best_match", C.c_int32),
        ("hit_order", C.c_int32),
        ("all_tuples", C.c_int32),
        ("flags", C.c_int32),
    ]


class jb_pdb_info(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int64),
        ("header_offset", C.c_int64),
        ("header_length", C.c_int64),
        ("stop_offset", C.c_int64),
        ("mmcif_tags", C.c_int32),
        ("n_headers", C.c_int32),
    ]


class jb_cif_info(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int64),
        ("loop_offset", C.c_int64),
        ("loop_end", C.c_int64),
        ("n_blocks", C.c_int32),
        ("hetatm_unnumbered", C.c_int32),
        ("fallback", C.c_int32),
        ("n_meta", C.c_int32),
        ("name_offset", C.c_int64),
        ("name_length", C.c_int64),
        ("meta", (C.c_int64 * 2) * 16),
    ]


class jb_hits(C.Structure):
    _fields_ = [
        ("count", C.c_int64),
        ("stride", C.c_int32),
        ("reserved", C.c_int32),
        ("molecule", c_i32p),
        ("template_index", c_i32p),
        ("atoms", c_i32p),
        ("rmsd", c_f64p),
        ("rotation", c_f64p),
        ("centre0", c_f64p),
        ("centre1", c_f64p),
        ("coords", c_f64p),
    ]


class jb_stats(C.Structure):
    _fields_ = [
        ("pairs", C.c_int64),
        ("live_pairs", C.c_int64),
        ("candidate_atoms", C.c_int64),
        ("candidate_records", C.c_int64),
        ("frontier_items", C.c_int64),
        ("frontier_ints", C.c_int64),
        ("candidate_tests", C.c_int64),
        ("scan_tests", C.c_int64),
        ("box_survivors", C.c_int64),
        ("tuples", C.c_int64),
        ("hits", C.c_int64),
        ("nan_dropped", C.c_int64),
        ("chunks", C.c_int64),
        ("retries", C.c_int64),
        ("launches", C.c_int64),
        ("kernel_ms", C.c_double),
        ("classify_ms", C.c_double),
        ("pairs_ms", C.c_double),
        ("expand_ms", C.c_double),
        ("superpose_ms", C.c_double),
    ]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


def _ptr(arr: np.ndarray, typ):
    return arr.ctypes.data_as(typ)


class TemplateTable:
    """Concatenated tables of a template list (insertion order)."""

    def __init__(self, templates: Sequence[Template]):
        templates = list(templates)
        self.n_templates = len(templates)
        sizes = [len(t) for t in templates]
        self.sizes = np.asarray(sizes, dtype=np.int32)
        self.max_atoms = max(sizes, default=0)
        self.atom_offset = np.zeros(self.n_templates + 1, dtype=np.int32)
        np.cumsum(sizes, out=self.atom_offset[1:])
        ta = int(self.atom_offset[-1])
        self.match_mode = np.zeros(ta, dtype=np.int32)
        self.residue_number = np.zeros(ta, dtype=np.int32)
        self.chain_id2 = np.zeros(ta, dtype=np.uint8)
        self.position = np.zeros((ta, 3), dtype=np.float64)
        self.distance_weight = np.zeros(ta, dtype=np.float64)
        self.name_offset = np.zeros(ta + 1, dtype=np.int32)
        self.resname_offset = np.zeros(ta + 1, dtype=np.int32)
        names, resnames = [], []
        k = 0
        for t in templates:
            for a in t:
                self.match_mode[k] = a._code
                self.residue_number[k] = a._resseq
                self.chain_id2[k] = a._chain2
                self.position[k] = a._pos
                self.distance_weight[k] = a._weight
                nt, rt = a._name_tokens(), a._resname_tokens()
                names.append(nt)
                resnames.append(rt)
                self.name_offset[k + 1] = self.name_offset[k] + len(nt)
                self.resname_offset[k + 1] = self.resname_offset[k] + len(rt)
                k += 1
        self.atom_names = np.concatenate(names).astype("<u4") if names else np.zeros(0, "<u4")
        self.residue_names = np.concatenate(resnames).astype("<u4") if resnames else np.zeros(0, "<u4")
        if len(self.atom_names) == 0:
            self.atom_names = np.zeros(1, "<u4")
        if len(self.residue_names) == 0:
            self.residue_names = np.zeros(1, "<u4")
        self.distance_offset = np.zeros(self.n_templates + 1, dtype=np.int64)
        np.cumsum([n * n for n in sizes], out=self.distance_offset[1:])
        dist = [t._dist.reshape(-1) for t in templates]
        self.distances = np.concatenate(dist) if dist else np.zeros(0, np.float64)
        if len(self.distances) == 0:
            self.distances = np.zeros(1, np.float64)
        self._struct = None

    def as_struct(self) -> jb_template_table:
        if self._struct is None:
            s = jb_template_table()
            s.n_templates = self.n_templates
            s.atom_offset = _ptr(self.atom_offset, c_i32p)
            s.match_mode = _ptr(self.match_mode, c_i32p)
            s.residue_number = _ptr(self.residue_number, c_i32p)
            s.chain_id2 = _ptr(self.chain_id2, c_u8p)
            s.position = _ptr(self.position, c_f64p)
            s.distance_weight = _ptr(self.distance_weight, c_f64p)
            s.name_offset = _ptr(self.name_offset, c_i32p)
            s.atom_names = _ptr(self.atom_names, c_u32p)
            s.resname_offset = _ptr(self.resname_offset, c_i32p)
            s.residue_names = _ptr(self.residue_names, c_u32p)
            s.distance_offset = _ptr(self.distance_offset, c_i64p)
            s.distances = _ptr(self.distances, c_f64p)
            self._struct = s
        return self._struct


class _PlainArray:
    def __init__(self, n, dtype):
        self.array = np.zeros(n, dtype=dtype)


def _staging_array(n, dtype):
    """Long-lived staging memory of a thread's pool: page-locked through the engine library
    (`jb_host_alloc`) so that the upload is one asynchronous DMA; pageable numpy memory where
    no CUDA device exists (host-logic tests over a stubbed engine -- staging, never compute)."""
    from . import _capi
    try:
        return _capi.PinnedArray(n, dtype)
    except _capi.EngineError:
        return _PlainArray(n, dtype)


class MoleculeBatch:
    """Structure-of-arrays image of a list of molecules.

    With ``pinned=True`` the arrays live in page-locked host memory (allocated
    through torch, which is plumbing only) so that the upload is a true async
    DMA; this is the staging buffer of SURVEY 7.2 "Host".
    """

    FIELDS = (("x", np.float64), ("y", np.float64), ("z", np.float64), ("name", np.uint32),
              ("resname", np.uint32), ("resseq", np.int32), ("chain", np.uint8))

    def __init__(self, molecules: Sequence[Molecule], pinned: bool = False, serial: bool = False, pool=None):
        """`serial=True` also ships Atom.serial (only the reference hit order reads it).
        `pool`: a dict that keeps the staging arrays alive between short-lived batches of one
        thread (no first-touch page faults on every query); the batch must not outlive the
        next batch built from the same pool."""
        self._pool = pool
        molecules = list(molecules)
        self.molecules = molecules
        self.n_molecules = len(molecules)
        counts = np.asarray([len(m._atoms) for m in molecules], dtype=np.int64)
        self.atom_offset = np.zeros(self.n_molecules + 1, dtype=np.int64)
        total = int(counts.sum())
        self.n_atoms = total
        self._keep = []
        for field, dtype in self.FIELDS:
            setattr(self, field, self._alloc(max(total, 1), dtype, pinned))
        self.serial = self._alloc(max(total, 1), np.int32, pinned) if serial else None
        # one native pass over the molecules' record arrays (jb_molecules_pack, csrc/ingest.cu)
        from . import _capi
        lib = _capi.load_library()
        recs = [np.ascontiguousarray(m._atoms) for m in molecules]
        ptrs = np.asarray([r.ctypes.data for r in recs], dtype=np.uintp)
        ptr = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
        rc = lib.jb_molecules_pack(ptr(ptrs), ptr(counts), self.n_molecules, ptr(self.atom_offset), ptr(self.x),
                                   ptr(self.y), ptr(self.z), ptr(self.name), ptr(self.resname), ptr(self.resseq),
                                   ptr(self.chain), ptr(self.serial) if serial else None)
        if rc != 0:
            raise RuntimeError(f"jb_molecules_pack failed with status {rc}")
        self._struct = None

    def _alloc(self, n, dtype, pinned):
        if self._pool is not None and not pinned:
            key = (len(self._keep), np.dtype(dtype).str)
            self._keep.append(key)
            buf = self._pool.get(key)
            if buf is None or len(buf.array) < n:
                buf = self._pool[key] = _staging_array(max(n, 1 << 16) + n // 4, dtype)
            return buf.array[:n]
        if pinned:
            import torch
            t = torch.empty(n * np.dtype(dtype).itemsize, dtype=torch.uint8, pin_memory=True)
            self._keep.append(t)
            return t.numpy().view(dtype)
        return np.zeros(n, dtype=dtype)

    @property
    def nbytes(self) -> int:
        return int(self.atom_offset.nbytes + sum(getattr(self, f).nbytes for f, _ in self.FIELDS)
                   + (self.serial.nbytes if self.serial is not None else 0))

    def as_struct(self) -> jb_molecule_batch:
        if self._struct is None:
            s = jb_molecule_batch()
            s.n_molecules = self.n_molecules
            s.atom_offset = _ptr(self.atom_offset, c_i64p)
            s.x = _ptr(self.x, c_f64p)
            s.y = _ptr(self.y, c_f64p)
            s.z = _ptr(self.z, c_f64p)
            s.name = _ptr(self.name, c_u32p)
            s.residue_name = _ptr(self.resname, c_u32p)
            s.residue_number = _ptr(self.resseq, c_i32p)
            s.chain_id2 = _ptr(self.chain, c_u8p)
            s.serial = _ptr(self.serial, c_i32p) if self.serial is not None else None
            self._struct = s
        return self._struct


HIT_ORDER = {"canonical": 0, "reference": 1, "reference-noreorder": 2}


def query_params(rmsd_threshold, distance_cutoff, max_dynamic_distance, ignore_chain=None,
                 reorder=True, best_match=False, hit_order="canonical", all_tuples=False,
                 coords=False) -> jb_query_params:
    """`reorder`: True -> 2 (the engine's expected-fan-out scan order), False -> 0
    (template order); an int selects the mode explicitly (1 = the reference's argsort
    by |C_k|).  The hit set never depends on it.  `hit_order`: "canonical" (atom tuple
    ascending) or the reference's own iteration order for its reorder=True /
    reorder=False runs ("reference" / "reference-noreorder", include/jess_b200.h).  `coords`:
    also return the matched atoms transformed into the template frame (JB_QUERY_COORDS)."""
    if ignore_chain not in IGNORE_CHAIN:
        raise ValueError(f"invalid value for `ignore_chain`: {ignore_chain!r}")
    if isinstance(reorder, bool):
        reorder = 2 if reorder else 0
    if hit_order not in HIT_ORDER:
        raise ValueError(f"invalid value for `hit_order`: {hit_order!r}")
    return jb_query_params(float(rmsd_threshold), float(distance_cutoff), float(max_dynamic_distance),
                           IGNORE_CHAIN[ignore_chain], int(reorder), int(bool(best_match)), HIT_ORDER[hit_order],
                           int(bool(all_tuples)), 1 if coords else 0)


class HitTable:
    """Host copy of a `jb_hits` record set as numpy arrays (canonical order)."""

    __slots__ = ("count", "stride", "molecule", "template", "atoms", "rmsd", "rotation", "centre0", "centre1",
                 "coords", "_owner")

    def __init__(self, hits: jb_hits, owner=None):
        """With `owner` (the object that keeps the C arrays alive, e.g. a `_capi.Result`)
        the arrays are zero-copy views; without, they are copied."""
        n, stride = int(hits.count), int(hits.stride)
        self.count, self.stride = n, stride
        self._owner = owner

        def grab(ptr, shape, dtype):
            size = int(np.prod(shape))
            if n == 0 or size == 0:
                return np.zeros(shape, dtype=dtype)
            arr = np.ctypeslib.as_array(ptr, shape=(size,)).reshape(shape)
            return arr if owner is not None else arr.copy()

        self.molecule = grab(hits.molecule, (n,), np.int32)
        self.template = grab(hits.template_index, (n,), np.int32)
        self.atoms = grab(hits.atoms, (n, stride), np.int32)
        self.rmsd = grab(hits.rmsd, (n,), np.float64)
        self.rotation = grab(hits.rotation, (n, 3, 3), np.float64)
        self.centre0 = grab(hits.centre0, (n, 3), np.float64)
        self.centre1 = grab(hits.centre1, (n, 3), np.float64)
        # JB_QUERY_COORDS: matched atoms in the template frame, (n, stride, 3); None when not requested
        self.coords = grab(hits.coords, (n, stride, 3), np.float64) if hits.coords else None

    @classmethod
    def from_arrays(cls, molecule, template, atoms, rmsd, rotation, centre0, centre1) -> "HitTable":
        self = cls.__new__(cls)
        self.count, self.stride = int(len(molecule)), int(atoms.shape[1]) if atoms.ndim == 2 else 1
        self._owner = None
        self.molecule, self.template, self.atoms, self.rmsd = molecule, template, atoms, rmsd
        self.rotation, self.centre0, self.centre1 = rotation, centre0, centre1
        self.coords = None
        return self

    def arrays(self):
        return (self.molecule, self.template, self.atoms, self.rmsd, self.rotation, self.centre0, self.centre1)

    def keys(self):
        """Canonical hit identities: (molecule, template, atom-index tuple)."""
        return [
            (int(self.molecule[i]), int(self.template[i]), tuple(int(a) for a in self.atoms[i] if a >= 0))
            for i in range(self.count)
        ]
