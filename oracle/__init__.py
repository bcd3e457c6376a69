"""Parity oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  `oracle.query` runs our CPU restatement
(jess_oracle.c); `oracle.ref_query` runs the UNMODIFIED reference C library built by
oracle/Makefile into oracle/_ref/libjess_ref.so (absent => `ref_available()` False).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

from pyjess_b200._tables import (HitTable, MoleculeBatch, TemplateTable, jb_hits, jb_molecule_batch,
                                 jb_query_params, jb_stats, jb_template_table)

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libjess_ref.so")

_libs = {}


def build(force: bool = False) -> None:
    """Compile liboracle.so and, when /root/reference exists, _ref/libjess_ref.so."""
    args = ["make", "-C", HERE, "-s"] + (["-B"] if force else [])
    subprocess.run(args, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def _load(path: str, prefix: str):
    lib = _libs.get(path)
    if lib is None:
        if not os.path.exists(path):
            if prefix == "oracle":
                build()
            if not os.path.exists(path):
                raise FileNotFoundError(path)
        lib = C.CDLL(path)
        q = getattr(lib, f"{prefix}_query")
        q.restype = C.c_int
        q.argtypes = [C.POINTER(jb_template_table), C.POINTER(jb_molecule_batch), C.POINTER(jb_query_params),
                      C.c_int, C.POINTER(C.c_void_p)]
        h = getattr(lib, f"{prefix}_result_hits")
        h.restype = C.c_int
        h.argtypes = [C.c_void_p, C.POINTER(jb_hits)]
        s = getattr(lib, f"{prefix}_result_stats")
        s.restype = C.c_int
        s.argtypes = [C.c_void_p, C.POINTER(jb_stats)]
        f = getattr(lib, f"{prefix}_result_free")
        f.restype = None
        f.argtypes = [C.c_void_p]
        _libs[path] = lib
    return lib


def numpy_batch(molecules, serial: bool = False) -> MoleculeBatch:
    """SoA image of `molecules` packed with numpy alone: the reference arm of bench.py must not
    even map the product's engine library, which `MoleculeBatch(...)` packs through."""
    import numpy as np
    mols = list(molecules)
    b = MoleculeBatch.__new__(MoleculeBatch)
    b._pool, b._keep, b._struct = None, [], None
    b.molecules, b.n_molecules = mols, len(mols)
    counts = np.asarray([len(m._atoms) for m in mols], dtype=np.int64)
    b.atom_offset = np.zeros(len(mols) + 1, dtype=np.int64)
    np.cumsum(counts, out=b.atom_offset[1:])
    b.n_atoms = int(b.atom_offset[-1])
    rec = np.concatenate([m._atoms for m in mols]) if mols and b.n_atoms else np.zeros(1, dtype=mols[0]._atoms.dtype if mols else "V1")
    if b.n_atoms == 0:
        z = lambda dt: np.zeros(1, dtype=dt)  # noqa: E731
        b.x, b.y, b.z, b.name, b.resname, b.resseq, b.chain = (z(np.float64), z(np.float64), z(np.float64), z(np.uint32),
                                                                z(np.uint32), z(np.int32), z(np.uint8))
        b.serial = z(np.int32) if serial else None
        return b
    xyz = rec["x"]
    b.x, b.y, b.z = (np.ascontiguousarray(xyz[:, k]) for k in range(3))
    b.name = np.ascontiguousarray(rec["name"]).view(np.uint32).copy()
    b.resname = np.ascontiguousarray(rec["resName"]).view(np.uint32).copy()
    b.resseq = np.ascontiguousarray(rec["resSeq"], dtype=np.int32)
    b.chain = np.ascontiguousarray(rec["chainID2"], dtype=np.uint8)
    b.serial = np.ascontiguousarray(rec["serial"], dtype=np.int32) if serial else None
    return b


def ref_available() -> bool:
    return os.path.exists(REF_SO)


def _run(path, prefix, table: TemplateTable, batch: MoleculeBatch, params: jb_query_params, threads: int):
    lib = _load(path, prefix)
    handle = C.c_void_p()
    rc = getattr(lib, f"{prefix}_query")(C.byref(table.as_struct()), C.byref(batch.as_struct()),
                                         C.byref(params), int(threads), C.byref(handle))
    if rc != 0:
        raise RuntimeError(f"{prefix}_query failed with status {rc}")
    try:
        hits = jb_hits()
        rc = getattr(lib, f"{prefix}_result_hits")(handle, C.byref(hits))
        if rc != 0:
            raise RuntimeError(f"{prefix}_result_hits failed with status {rc}")
        stats = jb_stats()
        getattr(lib, f"{prefix}_result_stats")(handle, C.byref(stats))
        return HitTable(hits), stats.as_dict()
    finally:
        getattr(lib, f"{prefix}_result_free")(handle)


def query(table, batch, params, threads: int = 1):
    """CPU restatement: returns (HitTable, stats dict)."""
    return _run(ORACLE_SO, "oracle", table, batch, params, threads)


def ref_query(table, batch, params, threads: int = 1):
    """Unmodified reference C library: returns (HitTable, stats dict);
    stats['kernel_ms'] is the wall time of the query loops alone."""
    return _run(REF_SO, "ref", table, batch, params, threads)


def superpose(mol_xyz, tpl_xyz):
    """oracle_superpose on n x 3 arrays in template order -> (rmsd, rot, c0, c1)."""
    import numpy as np
    lib = _load(ORACLE_SO, "oracle")
    a = np.ascontiguousarray(mol_xyz, dtype=np.float64)
    b = np.ascontiguousarray(tpl_xyz, dtype=np.float64)
    n = a.shape[0]
    rmsd = C.c_double()
    rot = np.zeros(9); c0 = np.zeros(3); c1 = np.zeros(3)
    dp = C.POINTER(C.c_double)
    lib.oracle_superpose.restype = C.c_int
    lib.oracle_superpose.argtypes = [C.c_int, dp, dp, dp, dp, dp, dp]
    lib.oracle_superpose(n, a.ctypes.data_as(dp), b.ctypes.data_as(dp), C.byref(rmsd),
                         rot.ctypes.data_as(dp), c0.ctypes.data_as(dp), c1.ctypes.data_as(dp))
    return rmsd.value, rot.reshape(3, 3), c0, c1


def candidates(table: TemplateTable, batch: MoleculeBatch, mol: int, tpl: int, k: int):
    """Candidate set C_k (atom indices within the molecule) of template atom k."""
    import numpy as np
    lib = _load(ORACLE_SO, "oracle")
    cap = int(batch.atom_offset[mol + 1] - batch.atom_offset[mol]) + 1
    out = np.zeros(cap, dtype=np.int32)
    lib.oracle_candidates.restype = C.c_int
    lib.oracle_candidates.argtypes = [C.POINTER(jb_template_table), C.POINTER(jb_molecule_batch), C.c_int, C.c_int,
                                      C.c_int, C.POINTER(C.c_int32), C.c_int]
    n = lib.oracle_candidates(C.byref(table.as_struct()), C.byref(batch.as_struct()), mol, tpl, k,
                              out.ctypes.data_as(C.POINTER(C.c_int32)), cap)
    if n < 0:
        raise RuntimeError("oracle_candidates failed")
    return out[:n].copy()
