"""ctypes binding of the C ABI (include/jess_b200.h) -- the ONLY compute path.

If ``libjess_b200.so`` is missing or no CUDA device is usable this module raises;
there is deliberately no CPU fallback (BASELINE.json north_star).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

from ._tables import (HitTable, MoleculeBatch, TemplateTable, jb_cif_info, jb_hits, jb_molecule_batch, jb_pdb_info,
                      jb_query_params, jb_stats, jb_template_table)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JB_LIB_PATH") or os.path.join(HERE, "libjess_b200.so")   # JB_LIB_PATH: A/B builds

EXPORTS = [
    "jb_abi_version", "jb_last_error", "jb_device_count", "jb_host_alloc", "jb_host_free",
    "jb_context_create", "jb_context_destroy", "jb_context_synchronize", "jb_context_trim",
    "jb_templates_create", "jb_templates_free", "jb_templates_spec_count",
    "jb_molecules_upload", "jb_molecules_free",
    "jb_query_run", "jb_result_fetch", "jb_result_stats", "jb_result_free", "jb_query_host", "jb_query_records",
    "jb_pdb_parse", "jb_cif_parse", "jb_format_atoms_pdb", "jb_molecules_pack",
]

_lib = None
_lock = threading.Lock()


class EngineError(RuntimeError):
    """A C-ABI call failed (status code + jb_last_error text)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"jess_b200 status {status}: {message}")
        self.status = status


def load_library():
    """dlopen the engine; raises if it was not built (never falls back)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} not found: build it with `python -m pyjess_b200._build` "
                    "(there is no CPU fallback)")
            lib = C.CDLL(LIB_PATH)
            vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
            P = C.POINTER
            sig = {
                "jb_abi_version": (i32, []),
                "jb_last_error": (C.c_char_p, []),
                "jb_device_count": (i32, []),
                "jb_host_alloc": (i32, [i64, P(vp)]),
                "jb_host_free": (None, [vp]),
                "jb_context_create": (i32, [i32, vp, i64, P(vp)]),
                "jb_context_destroy": (None, [vp]),
                "jb_context_synchronize": (i32, [vp]),
                "jb_context_trim": (i32, [vp]),
                "jb_templates_create": (i32, [vp, P(jb_template_table), P(vp)]),
                "jb_templates_free": (None, [vp]),
                "jb_templates_spec_count": (i32, [vp]),
                "jb_molecules_upload": (i32, [vp, P(jb_molecule_batch), P(vp)]),
                "jb_molecules_free": (None, [vp]),
                "jb_query_run": (i32, [vp, vp, vp, P(jb_query_params), P(vp)]),
                "jb_result_fetch": (i32, [vp, P(jb_hits)]),
                "jb_result_stats": (i32, [vp, P(jb_stats)]),
                "jb_result_free": (None, [vp]),
                "jb_query_host": (i32, [vp, vp, P(jb_molecule_batch), P(jb_query_params), P(vp), P(jb_hits)]),
                "jb_query_records": (i32, [vp, vp, vp, vp, i32, P(jb_query_params), P(vp), P(jb_hits)]),
                "jb_pdb_parse": (i32, [C.c_char_p, i64, i32, vp, i64, P(jb_pdb_info)]),
                "jb_cif_parse": (i32, [C.c_char_p, i64, i32, vp, i64, P(jb_cif_info)]),
                "jb_format_atoms_pdb": (i64, [vp, i64, vp, vp, i64]),
                "jb_molecules_pack": (i32, [vp, vp, i32] + [vp] * 9),
            }
            for name, (res, args) in sig.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def _check(status: int):
    if status != 0:
        raise EngineError(status, load_library().jb_last_error().decode("utf-8", "replace"))


class Context:
    """One engine context = one GPU + one stream (include/jess_b200.h: jb_context)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None, workspace_bytes: int = 0):
        self.lib = load_library()
        self._h = C.c_void_p()
        _check(self.lib.jb_context_create(int(device), C.c_void_p(stream or 0), int(workspace_bytes), C.byref(self._h)))
        self.device = int(device)
        # one stream + one grow-only workspace per context: engine calls are serialised.  The
        # reference's `Jess.query` is re-entrant (README thread-pool recipe); here concurrent
        # threads simply queue on the GPU (ctypes drops the GIL while a call runs).
        self.lock = threading.RLock()

    def close(self):
        if self._h:
            self.lib.jb_context_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(self.lib.jb_context_synchronize(self._h))

    def trim(self):
        """Give the context's grow-only device buffers back (`jb_context_trim`)."""
        with self.lock:
            _check(self.lib.jb_context_trim(self._h))


class DeviceTemplates:
    def __init__(self, ctx: Context, table: TemplateTable):
        self.ctx, self.table = ctx, table
        self._h = C.c_void_p()
        with ctx.lock:
            _check(ctx.lib.jb_templates_create(ctx._h, C.byref(table.as_struct()), C.byref(self._h)))

    @property
    def spec_count(self) -> int:
        return int(self.ctx.lib.jb_templates_spec_count(self._h))

    def close(self):
        if self._h and self.ctx._h:
            with self.ctx.lock:           # never while another thread runs a query on this context
                self.ctx.lib.jb_templates_free(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceMolecules:
    def __init__(self, ctx: Context, batch: MoleculeBatch):
        self.ctx, self.batch = ctx, batch
        self._h = C.c_void_p()
        with ctx.lock:
            _check(ctx.lib.jb_molecules_upload(ctx._h, C.byref(batch.as_struct()), C.byref(self._h)))

    def close(self):
        if self._h and self.ctx._h:
            with self.ctx.lock:
                self.ctx.lib.jb_molecules_free(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Result:
    def __init__(self, ctx: Context, handle):
        self.ctx, self._h = ctx, handle

    def fetch(self) -> HitTable:
        hits = jb_hits()
        with self.ctx.lock:
            _check(self.ctx.lib.jb_result_fetch(self._h, C.byref(hits)))
        return HitTable(hits, owner=self)      # zero-copy views; the table keeps this result alive

    def stats(self) -> dict:
        st = jb_stats()
        _check(self.ctx.lib.jb_result_stats(self._h, C.byref(st)))
        return st.as_dict()

    def close(self):
        # No context lock here: a hit table is often garbage-collected in one thread while another thread
        # queries (README thread-pool recipe), and jb_result_free only touches the pool of page-locked
        # blocks, which the library guards with its own mutex.
        if self._h and self.ctx._h:       # a result never outlives its context
            self.ctx.lib.jb_result_free(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def query_run(ctx: Context, templates: DeviceTemplates, molecules: DeviceMolecules, params: jb_query_params) -> Result:
    """Kernels only: inputs already resident in HBM, hits stay in HBM."""
    h = C.c_void_p()
    with ctx.lock:
        _check(ctx.lib.jb_query_run(ctx._h, templates._h, molecules._h, C.byref(params), C.byref(h)))
    return Result(ctx, h)


def query_host(ctx: Context, templates: DeviceTemplates, batch: MoleculeBatch, params: jb_query_params):
    """End to end from host buffers: upload + kernels + hit download.
    Returns (HitTable, stats)."""
    h = C.c_void_p()
    hits = jb_hits()
    with ctx.lock:
        _check(ctx.lib.jb_query_host(ctx._h, templates._h, C.byref(batch.as_struct()), C.byref(params), C.byref(h),
                                     C.byref(hits)))
    res = Result(ctx, h)
    return HitTable(hits, owner=res), res.stats()


def query_records(ctx: Context, templates: DeviceTemplates, molecules, params: jb_query_params):
    """End to end from `Molecule` objects (`jb_query_records`): the engine packs their atom records
    into page-locked staging slice by slice, overlapped with upload and kernels.
    Returns (HitTable, stats)."""
    import numpy as np
    n = len(molecules)
    recs = [m._records() for m in molecules]              # (array kept alive, address, count)
    ptrs = np.fromiter((r[1] for r in recs), dtype=np.uintp, count=n)
    counts = np.fromiter((r[2] for r in recs), dtype=np.int64, count=n)
    h = C.c_void_p()
    hits = jb_hits()
    with ctx.lock:
        _check(ctx.lib.jb_query_records(ctx._h, templates._h, C.c_void_p(ptrs.ctypes.data), C.c_void_p(counts.ctypes.data),
                                        n, C.byref(params), C.byref(h), C.byref(hits)))
    res = Result(ctx, h)
    del recs                                              # (kept alive until the call returned)
    return HitTable(hits, owner=res), res.stats()


def pdb_parse(data: bytes, records, skip_hetatm: bool = False, ignore_endmdl: bool = False) -> jb_pdb_info:
    """One-pass native PDB parser (host code of the engine library, include/jess_b200.h:
    jb_pdb_parse): fills the numpy record array `records` (dtype `_model.ATOM_DTYPE`)."""
    lib = load_library()
    info = jb_pdb_info()
    flags = (1 if skip_hetatm else 0) | (2 if ignore_endmdl else 0)
    rc = lib.jb_pdb_parse(data, len(data), flags, C.c_void_p(records.ctypes.data), len(records), C.byref(info))
    if rc not in (0, -4):
        _check(rc)
    return info


def cif_parse(data: bytes, dtype, skip_hetatm: bool = False, ignore_endmdl: bool = False, use_author: bool = False):
    """One-pass native reader of the `_atom_site` loop of an mmCIF file (`jb_cif_parse`).
    Returns (info, records); `info.fallback != 0`: the generic Python loader must decide."""
    import numpy as np
    lib = load_library()
    info = jb_cif_info()
    flags = (1 if skip_hetatm else 0) | (2 if ignore_endmdl else 0) | (4 if use_author else 0)
    cap = len(data) // 40 + 16                        # a row holds >= 16 values; the reader reports what it needs
    while True:
        records = np.zeros(cap, dtype=dtype)
        rc = lib.jb_cif_parse(data, len(data), flags, C.c_void_p(records.ctypes.data), len(records), C.byref(info))
        if rc == -4 and not info.fallback:            # JB_ERR_CAPACITY: rows not at line starts
            cap = int(info.n_atoms)
            continue
        if rc != 0:
            _check(rc)
        return info, records[:min(int(info.n_atoms), cap)]


_fmt_tls = threading.local()


def format_atoms_pdb(records, xyz=None) -> str:
    """ATOM lines of a block of atom records (`jb_format_atoms_pdb`): the sprintf loop of
    Hit.dump.  `records`: C-contiguous `ATOM_DTYPE` array; `xyz`: C-contiguous float64 (n, 3)."""
    lib = _lib or load_library()
    n = len(records)
    buf = getattr(_fmt_tls, "buf", None)  # per thread: ctypes drops the GIL during the call
    if buf is None or len(buf) < 96 * n + 1:
        buf = _fmt_tls.buf = C.create_string_buffer(max(96 * n + 1, 1 << 14))
    if not records.flags.c_contiguous:
        records = records.copy()
    xp = None
    if xyz is not None:
        if not (xyz.flags.c_contiguous and xyz.dtype.str == "<f8"):
            xyz = xyz.astype("<f8", order="C")
        xp = xyz.ctypes.data
    got = lib.jb_format_atoms_pdb(records.ctypes.data, n, xp, buf, len(buf))
    if got < 0:
        _check(int(got))
    return C.string_at(buf, got).decode("ascii", "replace")


class PinnedArray:
    """numpy view of page-locked host memory from `jb_host_alloc` (freed with the object)."""

    def __init__(self, n: int, dtype):
        import numpy as np
        self.lib = load_library()
        self._p = C.c_void_p()
        dt = np.dtype(dtype)
        _check(self.lib.jb_host_alloc(int(n) * dt.itemsize, C.byref(self._p)))
        raw = (C.c_uint8 * (max(int(n), 1) * dt.itemsize)).from_address(self._p.value)
        self.array = np.frombuffer(raw, dtype=dt, count=int(n))

    def __del__(self):
        try:
            if self._p:
                self.lib.jb_host_free(self._p)
                self._p = C.c_void_p()
        except Exception:
            pass


_default = {}
_default_lock = threading.Lock()


def default_context(device: Optional[int] = None) -> Context:
    """Process-wide context per device (LOCAL_RANK picks the GPU under torchrun)."""
    if device is None:
        device = int(os.environ.get("JB_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    ctx = _default.get(device)
    if ctx is None:
        with _default_lock:
            ctx = _default.get(device)
            if ctx is None:
                ctx = _default[device] = Context(device)
    return ctx


def device_count() -> int:
    return int(load_library().jb_device_count())
