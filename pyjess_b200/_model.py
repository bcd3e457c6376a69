"""Host-side object model: `Mat4`, `Atom`, `Molecule`, `TemplateAtom`, `Template`.

Mirrors the public surface of the reference's Cython classes
(``src/pyjess/_jess.pyx``: Mat4 :177-240, Molecule :446-900, Atom :903-1330,
TemplateAtom :1333-1750, Template :1753-2087) so that the parity tests read like
the reference's own tests.  The storage is different by design: a `Molecule` is
ONE numpy structured array (a struct-of-records image of the reference's 64-byte
``struct _Atom``, eff:Atom.h:29-47) from which the packed structure-of-arrays
upload image for the GPU is cut without touching Python objects
(`Molecule._packed`); a `Template` keeps flat numpy tables ready to be
concatenated into a `jb_template_table` (include/jess_b200.h).

Token conventions are reproduced exactly because they decide
candidate membership: blanks/NULs become ``_`` (``encode_token``
_jess.pyx:133-140, ``Atom_copyToken`` vendor/jess/src/Atom.c:17-30); names shorter
than 4 given through the Python API are right-shifted by one ``_``
(_jess.pyx:1052-1056, :1483-1488).
"""
from __future__ import annotations

import datetime
import functools
import io
import os
import re
from typing import Iterable, List, Optional, Sequence

import numpy as np

__all__ = ["Mat4", "Atom", "Molecule", "TemplateAtom", "Template", "MATCH_MODES"]

MATCH_MODES = frozenset(range(-1, 9)) | frozenset(range(100, 108))

#: record layout of one atom (field order of eff:Atom.h:29-47)
ATOM_DTYPE = np.dtype(
    [
        ("x", "<f8", (3,)),
        ("occupancy", "<f4"),
        ("tempFactor", "<f4"),
        ("serial", "<i4"),
        ("name", "S4"),
        ("altLoc", "u1"),
        ("resName", "S4"),  # 3 chars + NUL, compared over 4 bytes
        ("chainID1", "u1"),
        ("chainID2", "u1"),
        ("resSeq", "<i4"),
        ("iCode", "u1"),
        ("segID", "S4"),
        ("element", "S2"),
        ("charge", "i1"),
    ]
)


# --- token helpers -----------------------------------------------------------

def _encode_token(src: bytes, n: int) -> bytes:
    """blank / NUL -> ``_`` over exactly n bytes (_jess.pyx:133-140)."""
    src = src[:n].ljust(n, b"\0")
    return bytes(0x5F if c in (0x20, 0) else c for c in src)


def _copy_token(src: bytes, n: int) -> bytes:
    """``Atom_copyToken`` (vendor/jess/src/Atom.c:17-30): strncpy + isspace -> ``_``."""
    src = src[:n]
    nul = src.find(b"\0")
    if nul >= 0:
        src = src[:nul]
    out = bytes(0x5F if c in b" \t\n\v\f\r" else c for c in src)
    return out.ljust(n, b"_")          # the NUL padding of strncpy is turned into '_' as well


def _api_atom_name(name) -> bytes:
    raw = bytearray(name, "ascii") if isinstance(name, str) else bytearray(name)
    if len(raw) > 4:
        raise ValueError(f"Invalid atom name: {name!r}")
    return raw


_HEADER_EOL = re.compile(rb"[\r\n]")
_NEXT_HEADER = re.compile(rb"[\r\n]HEADER")
_INT_RE = re.compile(rb"[ \t\n\v\f\r]*([+-]?\d+)")
_FLT_RE = re.compile(
    rb"[ \t\n\v\f\r]*([+-]?(?:(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?|inf(?:inity)?|nan))", re.I
)


def _atoi(b: bytes) -> int:
    m = _INT_RE.match(b)
    if not m:
        return 0
    v = max(-(2**63), min(2**63 - 1, int(m.group(1))))     # atoi = (int) strtol: saturates in long ...
    return ((v + 2**31) % 2**32) - 2**31                      # ... and truncates to int


_HEX_RE = re.compile(rb"[ \t\n\v\f\r]*([+-]?0[xX](?:[0-9a-fA-F]+\.?[0-9a-fA-F]*|\.[0-9a-fA-F]+)(?:[pP][+-]?\d+)?)")


def _atof(b: bytes) -> float:
    m = _HEX_RE.match(b)           # strtod also reads C99 hexadecimal floats
    if m:
        return float.fromhex(m.group(1).decode("ascii"))
    m = _FLT_RE.match(b)
    return float(m.group(1)) if m else 0.0


def _parse_atom_line(line: bytes, rec) -> bool:
    """Fixed-column PDB ATOM/HETATM parser with the semantics of ``Atom_parse``
    (vendor/jess/src/Atom.c:32-72): C ``atoi``/``atof`` prefix parsing from the
    field start, fields beyond the end of the line keep their zero value."""
    if not (line.startswith(b"ATOM") or line.startswith(b"HETATM")):
        return False
    nul = line.find(b"\0")
    if nul >= 0:
        line = line[:nul]
    n = len(line)
    rec["x"] = 0.0
    rec["occupancy"] = 0.0
    rec["tempFactor"] = 0.0
    rec["serial"] = 0
    rec["name"] = b""
    rec["altLoc"] = 0
    rec["resName"] = b""
    rec["chainID1"] = 0
    rec["chainID2"] = 0
    rec["resSeq"] = 0
    rec["iCode"] = 0
    rec["segID"] = b""
    rec["element"] = b""
    rec["charge"] = 0
    if n > 6:
        rec["serial"] = _atoi(line[6:])
    if n > 12:
        rec["name"] = _copy_token(line[12:], 4)
    if n > 16:
        rec["altLoc"] = line[16]
    if n > 17:
        rec["resName"] = _copy_token(line[17:], 3)
    if n > 20:
        rec["chainID1"] = line[20]
    if n > 21:
        rec["chainID2"] = 0x30 if line[21:22] in (b" ", b"\t", b"\n", b"\v", b"\f", b"\r") else line[21]
    if n > 22:
        rec["resSeq"] = _atoi(line[22:])
    if n > 26:
        rec["iCode"] = line[26]
    x = [0.0, 0.0, 0.0]
    if n > 30:
        x[0] = _atof(line[30:])
    if n > 38:
        x[1] = _atof(line[38:])
    if n > 46:
        x[2] = _atof(line[46:])
    rec["x"] = x
    if n > 54:
        rec["occupancy"] = np.float32(_atof(line[54:]))
    if n > 60:
        rec["tempFactor"] = np.float32(_atof(line[60:]))
    if n > 72:
        rec["segID"] = _copy_token(line[72:], 4)
    if n > 76:
        rec["element"] = _copy_token(line[76:], 2)
    if n > 78:
        c = _atoi(line[78:])
        rec["charge"] = ((c + 128) % 256) - 128
    return True


def _as_line(text) -> bytes:
    b = bytearray(text, "utf-8") if isinstance(text, str) else bytearray(text)
    if not b.endswith(b"\n"):
        b.append(0x0A)
    return bytes(b)


def _format_atom(rec, x) -> str:
    """PDB line exactly as the reference's sprintf (_jess.pyx:1273-1291)."""
    name = rec["name"].tobytes()[:4].replace(b"_", b" ").replace(b"\0", b" ").decode("ascii")
    resname = rec["resName"].tobytes()[:3].replace(b"_", b" ").replace(b"\0", b" ").decode("ascii")
    elem = rec["element"].tobytes()[:2].replace(b"_", b" ").replace(b"\0", b" ").decode("ascii")
    return "ATOM  %5i%5s%c%-3s%c%c%4i%-4c%8.3f%8.3f%8.3f%6.2f%6.2f          %2s\n" % (
        int(rec["serial"]),
        name.rstrip("\0"),
        chr(int(rec["altLoc"]) or 32),
        resname,
        chr(int(rec["chainID1"]) or 32),
        chr(int(rec["chainID2"]) or 32),
        int(rec["resSeq"]),
        chr(int(rec["iCode"]) or 32),
        x[0],
        x[1],
        x[2],
        float(rec["occupancy"]),
        float(rec["tempFactor"]),
        elem,
    )


# --- Mat4 --------------------------------------------------------------------

class Mat4:
    """A 4x4 matrix in homogeneous coordinates (_jess.pyx:177-240)."""

    __slots__ = ("_data",)

    def __init__(self, data):
        arr = np.array(data, dtype=np.float64)
        if arr.shape != (4, 4):
            raise ValueError(f"dimension error: expected shape (4, 4), got {arr.shape!r}")
        self._data = arr

    @classmethod
    def identity(cls) -> "Mat4":
        return cls(np.eye(4))

    def __matmul__(self, other):
        if not isinstance(other, Mat4):
            return NotImplemented
        out = Mat4.__new__(Mat4)
        out._data = self._data @ other._data
        return out

    def __array__(self, dtype=None, copy=None):
        return self._data if dtype is None else self._data.astype(dtype)

    def __getitem__(self, index):
        return self._data[index]

    def __len__(self):
        return 4

    def __eq__(self, other):
        if not isinstance(other, Mat4):
            return NotImplemented
        return bool(np.array_equal(self._data, other._data))

    def __hash__(self):
        return hash(self._data.tobytes())

    def __repr__(self):
        return f"Mat4({self._data.tolist()!r})"

    def __reduce__(self):
        return Mat4, (self._data.tolist(),)

    @property
    def shape(self):
        return (4, 4)


def _as_matrix(matrix) -> np.ndarray:
    m = np.asarray(matrix, dtype=np.float64)
    if m.ndim != 2 or (m.shape[0] != 4 and m.shape[1] != 4):
        raise ValueError(f"dimension error: expected shape (4, 4), got {tuple(m.shape)!r}")
    return m


def _transform_points(xyz: np.ndarray, m: np.ndarray) -> np.ndarray:
    """x' = M[:3,:4] @ (x,1), summed left to right like ``Atom._transform``
    (_jess.pyx:1293-1307)."""
    out = np.zeros_like(xyz)
    for j in range(3):
        out += m[:3, j][None, :] * xyz[:, j][:, None]
    out += m[:3, 3][None, :] * 1.0
    return out


# --- Atom --------------------------------------------------------------------

class Atom:
    """A single atom of a molecule (_jess.pyx:903-1330)."""

    __slots__ = ("_rec", "_owner")

    @classmethod
    def load(cls, file):
        return cls.loads(file.read())

    @classmethod
    def loads(cls, text):
        atom = cls.__new__(cls)
        arr = np.zeros(1, dtype=ATOM_DTYPE)
        if not _parse_atom_line(_as_line(text), arr[0]):
            raise ValueError(f"Failed to parse atom: {text!r}")
        atom._rec = arr
        atom._owner = None
        return atom

    @classmethod
    def _view(cls, arr: np.ndarray, owner=None) -> "Atom":
        atom = cls.__new__(cls)
        atom._rec = arr
        atom._owner = owner
        return atom

    def __init__(
        self,
        *,
        serial: int,
        name: str,
        residue_name: str,
        chain_id: str,
        residue_number: int,
        x: float,
        y: float,
        z: float,
        occupancy: float = 0.0,
        temperature_factor: float = 0.0,
        altloc: str = " ",
        insertion_code: str = " ",
        segment: str = "",
        element: str = "",
        charge: int = 0,
    ):
        if len(name) > 4:
            raise ValueError(f"Invalid atom name: {name!r}")
        if len(residue_name) > 3:
            raise ValueError(f"Invalid residue name: {residue_name!r}")
        if len(segment) > 4:
            raise ValueError(f"Invalid segment: {segment!r}")
        if len(element) > 2:
            raise ValueError(f"Invalid element: {element!r}")
        if len(chain_id) > 2:
            raise ValueError(f"Invalid chain ID: {chain_id!r}")
        arr = np.zeros(1, dtype=ATOM_DTYPE)
        r = arr[0]
        r["serial"] = serial
        r["altLoc"] = ord(altloc)
        r["resSeq"] = residue_number
        r["iCode"] = ord(insertion_code)
        r["x"] = (x, y, z)
        r["occupancy"] = occupancy
        r["tempFactor"] = temperature_factor
        r["charge"] = charge
        r["resName"] = _encode_token(residue_name.encode("ascii"), 3)
        r["segID"] = _encode_token(segment.encode("ascii"), 4)
        r["element"] = _encode_token(element.encode("ascii").rjust(2, b"\0"), 2)
        if len(chain_id) == 2:
            r["chainID1"], r["chainID2"] = ord(chain_id[0]), ord(chain_id[1])
        elif len(chain_id) == 1:
            r["chainID1"], r["chainID2"] = 0x20, ord(chain_id[0])
        else:
            r["chainID1"], r["chainID2"] = 0x20, 0x30
        raw = _api_atom_name(name)
        if len(raw) < 4:
            raw.insert(0, 0x5F)
        r["name"] = _encode_token(bytes(raw), 4)
        self._rec = arr
        self._owner = None

    def _state(self) -> dict:
        return {
            "serial": self.serial,
            "name": self.name,
            "altloc": self.altloc,
            "residue_name": self.residue_name,
            "chain_id": self.chain_id,
            "residue_number": self.residue_number,
            "insertion_code": self.insertion_code,
            "x": self.x,
            "y": self.y,
            "z": self.z,
            "temperature_factor": self.temperature_factor,
            "occupancy": self.occupancy,
            "segment": self.segment,
            "element": self.element,
            "charge": self.charge,
        }

    def __reduce__(self):
        return functools.partial(type(self), **self._state()), ()

    def __repr__(self):
        args = [f"{k}={v!r}" for k, v in self._state().items() if v is not None]
        return f"{type(self).__name__}({', '.join(args)})"

    def __eq__(self, other):
        if not isinstance(other, Atom):
            return NotImplemented
        return self._state() == other._state()

    def __hash__(self):
        return hash(tuple(self._state().values()))

    def __copy__(self):
        return self.copy()

    # fields
    @property
    def serial(self) -> int:
        return int(self._rec["serial"][0])

    @property
    def altloc(self) -> str:
        return chr(int(self._rec["altLoc"][0]))

    @property
    def name(self) -> str:
        return self._rec["name"].tobytes()[:4].decode("ascii").strip("_")       # raw bytes, like the reference

    @property
    def residue_name(self) -> str:
        return self._rec["resName"].tobytes()[:3].decode("ascii").strip("_")

    @property
    def residue_number(self) -> int:
        return int(self._rec["resSeq"][0])

    @property
    def segment(self) -> str:
        return self._rec["segID"].tobytes()[:4].decode("ascii").strip("_")      # raw bytes: NULs of a short line stay

    @property
    def element(self) -> str:
        return self._rec["element"].tobytes()[:2].decode("ascii").strip("_")

    @property
    def insertion_code(self) -> str:
        return chr(int(self._rec["iCode"][0]))

    @property
    def chain_id(self) -> str:
        return (chr(int(self._rec["chainID1"][0])) + chr(int(self._rec["chainID2"][0]))).strip()

    @property
    def occupancy(self) -> float:
        return float(self._rec["occupancy"][0])

    @property
    def temperature_factor(self) -> float:
        return float(self._rec["tempFactor"][0])

    @property
    def charge(self) -> int:
        return int(self._rec["charge"][0])

    @property
    def x(self) -> float:
        return float(self._rec["x"][0, 0])

    @property
    def y(self) -> float:
        return float(self._rec["x"][0, 1])

    @property
    def z(self) -> float:
        return float(self._rec["x"][0, 2])

    def copy(self) -> "Atom":
        return Atom._view(self._rec.copy())

    def dumps(self, format: str = "pdb") -> str:
        return _format_atom(self._rec[0], self._rec["x"][0])

    def dump(self, file, format: str = "pdb") -> None:
        file.write(self.dumps(format))

    def transform(self, matrix) -> "Atom":
        m = _as_matrix(matrix)
        out = self.copy()
        out._rec["x"] = _transform_points(self._rec["x"].copy(), m)
        return out


# --- Molecule ----------------------------------------------------------------

class Molecule:
    """A sequence of atoms (_jess.pyx:446-900), stored as one record array."""

    __slots__ = ("_atoms", "_id", "_date", "_name", "_pack", "_rec", "__weakref__")

    @classmethod
    def load(cls, file, format: str = "detect", *, id=None, date=None, name=None,
             ignore_endmdl: bool = False, use_author: bool = False, skip_hetatm: bool = False):
        """Read a PDB file (``PDBMoleculeParser.load``, src/pyjess/_io.py:68-103: the native
        one-pass parser) or an mmCIF file (``CIFMoleculeParser``, _io.py:105-245: `_cif`, no
        gemmi needed); ``format="detect"`` as src/pyjess/_jess.pyx:407-441."""
        try:
            handle = open(file, "rb")
            close = True
        except TypeError:
            handle, close = file, False
        try:
            data = handle.read()
        finally:
            if close:
                handle.close()
        if isinstance(data, str):
            data = data.encode("utf-8")
        if format not in ("pdb", "cif", "detect"):
            raise ValueError(f"invalid value for `format` argument: {format!r}")
        from . import _cif
        if format == "cif" or (format == "detect" and _cif.looks_like_cif(data)):
            return _cif.load_molecule(data, cls, id=id, date=date, name=name, use_author=use_author,
                                      skip_hetatm=skip_hetatm, ignore_endmdl=ignore_endmdl)
        return cls._parse_pdb(data, id=id, date=date, name=name,
                              ignore_endmdl=ignore_endmdl, skip_hetatm=skip_hetatm, format=format)

    @classmethod
    def load_many(cls, files, format: str = "detect", *, jobs: Optional[int] = None,
                  ids: Optional[Sequence[Optional[str]]] = None, **kwargs) -> List["Molecule"]:
        """`load` for a list of files on `jobs` threads (default: the host cores available to the
        process), results in input order.  The native parsers run without the GIL and the
        Python part of a PDB load is a few tens of microseconds, so ingest scales with the cores:
        what a batch for `Jess.query_many` needs (parsing, not querying, bounds a
        file-based run once the queries take milliseconds).  `ids`: one identifier per file
        (instead of the single `id` of `load`)."""
        files = list(files)
        if ids is not None:
            ids = list(ids)
            if len(ids) != len(files):
                raise ValueError("load_many: `ids` and `files` differ in length")
            if "id" in kwargs:
                raise TypeError("load_many: pass either `id` or `ids`")
            files = list(zip(files, ids))
            one = lambda item: cls.load(item[0], format, id=item[1], **kwargs)      # noqa: E731
        else:
            one = lambda f: cls.load(f, format, **kwargs)                          # noqa: E731
        if jobs is None:
            jobs = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        jobs = max(1, min(int(jobs), len(files)))
        if jobs == 1:
            return [one(f) for f in files]
        from concurrent.futures import ThreadPoolExecutor
        # a few contiguous runs of files per worker: one task per file costs more in queue traffic and
        # GIL hand-offs than a millisecond-sized parse gains
        size = max(1, -(-len(files) // (4 * jobs)))
        runs = [files[i:i + size] for i in range(0, len(files), size)]
        with ThreadPoolExecutor(jobs) as pool:
            parts = pool.map(lambda run: [one(f) for f in run], runs)
            return [mol for part in parts for mol in part]

    @classmethod
    def loads(cls, text, format: str = "detect", **kwargs):
        return cls.load(io.BytesIO(text.encode("utf-8") if isinstance(text, str) else bytes(text)),
                        format=format, **kwargs)

    @classmethod
    def _parse_pdb(cls, data: bytes, *, id, date, name, ignore_endmdl, skip_hetatm, format):
        """One pass of the engine library's native parser (`jb_pdb_parse`, csrc/ingest.cu)
        over the file image; only the optional HEADER line is decoded in Python."""
        from . import _capi
        data = bytes(data)
        # capacity: an ATOM line with coordinates is at least 54 characters; shorter ones make the parser
        # report the count it needs.  (No bytes.count / find over the file here: they hold the GIL, the
        # parser does not -- `load_many` reads files on several threads.)
        cap = len(data) // 54 + 16
        while True:
            arr = np.zeros(cap, dtype=ATOM_DTYPE)
            info = _capi.pdb_parse(data, arr, skip_hetatm=skip_hetatm, ignore_endmdl=ignore_endmdl)
            if info.mmcif_tags:
                raise ValueError("mmCIF data tags found, file is not in PDB format")
            if info.n_atoms <= cap:
                break
            cap = int(info.n_atoms)
        arr = arr[:info.n_atoms].copy() if info.n_atoms * 2 < cap else arr[:info.n_atoms]
        pos = info.header_offset
        while pos >= 0 and pos < info.stop_offset and (id is None or date is None or name is None):
            end = _HEADER_EOL.search(data, pos)           # a lone "\r" ends a line too (text mode of the reference)
            text = data[pos:end.start() if end else len(data)].decode("utf-8", "replace")
            if id is None:
                id = text[62:66].strip() or None
            if date is None:
                date_text = text[50:59].strip()
                if date_text:
                    date = datetime.datetime.strptime(date_text, "%d-%b-%y").date()
            if name is None:
                name = text[10:50].strip() or None
            if info.n_headers <= 1:
                break
            nxt = _NEXT_HEADER.search(data, pos + 1, info.stop_offset)
            pos = nxt.start() + 1 if nxt else -1
        return cls._from_records(arr, id=id, date=date, name=name)

    @classmethod
    def _parse_pdb_python(cls, data: bytes, *, id=None, date=None, name=None, ignore_endmdl=False, skip_hetatm=False):
        """Line-by-line Python restatement of the same loader (test reference for the
        native parser; `Atom.loads` shares `_parse_atom_line`)."""
        lines = []
        # the reference reads in text mode: universal newlines ("\r\n" and a lone "\r" end a line as "\n")
        data = data.replace(b"\r\n", b"\n").replace(b"\r", b"\n")
        for line in data.split(b"\n"):
            line += b"\n"
            if line.startswith(b"HEADER"):
                text = line.decode("utf-8", "replace")
                if id is None:
                    id = text[62:66].strip() or None
                if date is None:
                    date_text = text[50:59].strip()
                    if date_text:
                        date = datetime.datetime.strptime(date_text, "%d-%b-%y").date()
                if name is None:
                    name = text[10:50].strip() or None
            elif line.startswith(b"ATOM"):
                lines.append(line)
            elif line.startswith(b"HETATM"):
                if not skip_hetatm:
                    lines.append(line)
            elif line.startswith(b"ENDMDL"):
                if not ignore_endmdl:
                    break
            elif line[:5].lower() in (b"data_", b"loop_"):
                raise ValueError("mmCIF data tags found, file is not in PDB format")
        arr = np.zeros(len(lines), dtype=ATOM_DTYPE)
        for i, line in enumerate(lines):
            _parse_atom_line(line if line.endswith(b"\n") else line + b"\n", arr[i])
        return cls._from_records(arr, id=id, date=date, name=name)

    @classmethod
    def _from_records(cls, arr: np.ndarray, id=None, date=None, name=None) -> "Molecule":
        mol = cls.__new__(cls)
        mol._atoms = arr
        mol._id, mol._date, mol._name = id, date, name
        mol._pack = None
        return mol

    # -- adapters to third-party structure objects (src/pyjess/_jess.pyx:444-631) --------------
    # Duck-typed like the reference's: none of the three libraries is imported, only the
    # attributes the reference reads are touched, so the keyword each one feeds `Atom` is the same.

    @classmethod
    def from_biopython(cls, structure, id=None, date=None, name=None) -> "Molecule":
        """A `Bio.PDB` entity (structure, model or chain list) -> `Molecule`
        (src/pyjess/_jess.pyx:444-513); ``id`` defaults to ``structure.id``."""
        atoms = []
        for chain in structure.get_chains():
            for residue in chain.get_residues():
                _, residue_number, insertion_code = residue.id
                for a in residue.get_atoms():
                    xyz = a.get_coord()
                    atoms.append(Atom(
                        serial=a.serial_number, name=a.fullname, altloc=a.altloc,
                        residue_name=residue.resname, chain_id=chain.id,
                        residue_number=residue_number, insertion_code=insertion_code,
                        x=xyz[0], y=xyz[1], z=xyz[2],
                        occupancy=a.occupancy, temperature_factor=a.bfactor,
                        segment=residue.segid, element=a.element, charge=a.pqr_charge or 0,
                    ))
        return cls(atoms, id=structure.id if id is None else id, date=date, name=name)

    @classmethod
    def from_gemmi(cls, model, id=None, date=None, name=None) -> "Molecule":
        """A `gemmi.Model` -> `Molecule` (src/pyjess/_jess.pyx:515-564): the CRA walk of
        ``model.all()``; gemmi's NUL alternate location becomes a blank."""
        atoms = []
        for cra in model.all():
            a, r, c = cra.atom, cra.residue, cra.chain
            atoms.append(Atom(
                serial=a.serial, name=a.padded_name(),
                altloc=" " if a.altloc == "\0" else a.altloc,
                residue_name=r.name, chain_id=c.name,
                residue_number=r.seqid.num, insertion_code=r.seqid.icode,
                x=a.pos[0], y=a.pos[1], z=a.pos[2],
                occupancy=a.occ, temperature_factor=a.b_iso,
                segment=r.segment, element=a.element.name.upper(), charge=a.charge,
            ))
        return cls(atoms, id=id, date=date, name=name)

    @classmethod
    def from_biotite(cls, atom_array, id=None, date=None, name=None) -> "Molecule":
        """A `biotite.structure.AtomArray` -> `Molecule` (src/pyjess/_jess.pyx:566-631);
        the optional annotations (altloc, charge, occupancy, segment) take the reference's
        defaults when the array was read without them."""
        atoms = []
        for a in atom_array:
            atoms.append(Atom(
                serial=a.atom_id, name=str(a.atom_name),
                altloc=str(getattr(a, "altloc", " ")),
                residue_name=str(a.res_name), chain_id=str(a.chain_id),
                residue_number=a.res_id, insertion_code=str(a.ins_code).ljust(1),
                x=a.coord[0], y=a.coord[1], z=a.coord[2],
                occupancy=getattr(a, "occupancy", 1.0), temperature_factor=a.b_factor,
                segment=str(getattr(a, "segment", "")), element=str(a.element),
                charge=getattr(a, "charge", 0),
            ))
        return cls(atoms, id=id, date=date, name=name)

    @classmethod
    def from_arrays(cls, *, name, residue_name, residue_number, chain_id, x, y, z,
                    serial=None, occupancy=None, temperature_factor=None, id=None) -> "Molecule":
        """Vectorised constructor (no per-atom Python objects): names are given
        as PDB-style strings and encoded like `Atom.__init__`."""
        n = len(x)
        arr = np.zeros(n, dtype=ATOM_DTYPE)
        arr["x"][:, 0], arr["x"][:, 1], arr["x"][:, 2] = x, y, z
        arr["serial"] = np.arange(1, n + 1) if serial is None else serial
        arr["occupancy"] = 1.0 if occupancy is None else occupancy
        arr["tempFactor"] = 0.0 if temperature_factor is None else temperature_factor
        cache = {}

        def enc_name(s):
            v = cache.get(s)
            if v is None:
                raw = _api_atom_name(s)
                if len(raw) < 4:
                    raw.insert(0, 0x5F)
                v = cache[s] = _encode_token(bytes(raw), 4)
            return v

        arr["name"] = [enc_name(s) for s in name]
        rcache = {}
        arr["resName"] = [rcache.setdefault(s, _encode_token(s.encode("ascii"), 3)) for s in residue_name]
        arr["resSeq"] = residue_number
        if isinstance(chain_id, str):
            chain_id = [chain_id] * n
        arr["chainID1"] = [ord(c[0]) if len(c) == 2 else 0x20 for c in chain_id]
        arr["chainID2"] = [ord(c[-1]) if len(c) >= 1 else 0x30 for c in chain_id]
        arr["altLoc"] = 0x20
        arr["iCode"] = 0x20
        arr["segID"] = b"____"
        arr["element"] = b"__"
        return cls._from_records(arr, id=id)

    def __init__(self, atoms: Sequence[Atom] = (), id: Optional[str] = None,
                 date: Optional[datetime.date] = None, name: Optional[str] = None):
        if date is not None and not isinstance(date, datetime.date):
            raise TypeError(f"expected datetime.date or None, found {type(date).__name__} ")
        atoms = list(atoms)
        arr = np.zeros(len(atoms), dtype=ATOM_DTYPE)
        for i, atom in enumerate(atoms):
            if not isinstance(atom, Atom):
                raise TypeError(f"expected Atom, found {type(atom).__name__}")
            arr[i] = atom._rec[0]
        self._atoms = arr
        self._id, self._date, self._name = id, date, name
        self._pack = None

    def __len__(self) -> int:
        return len(self._atoms)

    def __getitem__(self, index):
        if isinstance(index, slice):
            return type(self)._from_records(self._atoms[index].copy(), id=self._id, date=self._date, name=self._name)
        n = len(self._atoms)
        i = int(index)
        if i < 0:
            i += n
        if i < 0 or i >= n:
            raise IndexError(index)
        return Atom._view(self._atoms[i:i + 1], owner=self)

    def __iter__(self):
        for i in range(len(self._atoms)):
            yield Atom._view(self._atoms[i:i + 1], owner=self)

    def __copy__(self):
        return self.copy()

    def __eq__(self, other):
        if not isinstance(other, Molecule):
            return NotImplemented
        if self._id != other._id or len(self) != len(other):
            return False
        return all(a == b for a, b in zip(self, other))

    def __hash__(self):
        return hash((self._id, *(hash(a) for a in self)))

    def __reduce__(self):
        return functools.partial(type(self), id=self._id, date=self._date, name=self._name), (list(self),)

    @property
    def id(self):
        return self._id

    @property
    def date(self):
        return self._date

    @property
    def name(self):
        return self._name

    def conserved(self, cutoff: float = 0.0) -> "Molecule":
        if cutoff <= 0.0:
            return self.copy()
        keep = self._atoms["tempFactor"] >= np.float32(cutoff)
        return type(self)._from_records(self._atoms[keep].copy(), id=self._id, date=self._date, name=self._name)

    def copy(self) -> "Molecule":
        return type(self)._from_records(self._atoms.copy(), id=self._id, date=self._date, name=self._name)

    def transform(self, matrix) -> "Molecule":
        m = _as_matrix(matrix)
        out = self.copy()
        out._atoms["x"] = _transform_points(self._atoms["x"].copy(), m)
        return out

    def dumps(self, format: str = "pdb", write_header: bool = True) -> str:
        buf = io.StringIO()
        self.dump(buf, format=format, write_header=write_header)
        return buf.getvalue()

    def dump(self, file, format: str = "pdb", write_header: bool = True) -> None:
        if format != "pdb":
            raise ValueError(f"invalid format: {format!r}")
        try:
            handle = open(file, "w")
            close = True
        except TypeError:
            handle, close = file, False
        try:
            if write_header:
                date = self._date.strftime("%d-%b-%y").upper() if self._date else ""
                handle.write("HEADER    %-40s%-9s   %-4s\n" % ((self._name or "")[:40], date, (self._id or "")[:4]))
            for i in range(len(self._atoms)):
                handle.write(_format_atom(self._atoms[i], self._atoms["x"][i]))
        finally:
            if close:
                handle.close()

    # --- device image ---------------------------------------------------------
    def _records(self):
        """(array, address, count) of the atom records as `jb_query_records` reads them (the
        reference hands Jess_query its `_Atom` array the same way).  The address is cached with
        the array it belongs to: taking it costs ~2 us, a batch holds thousands of structures."""
        a = self._atoms
        try:
            c = self._rec
        except AttributeError:
            c = None
        if c is None or c[0] is not a:
            if not a.flags.c_contiguous:
                b = np.ascontiguousarray(a)
                return b, b.ctypes.data, len(b)             # a copy: not cached, later edits must be seen
            c = self._rec = (a, a.ctypes.data, len(a))
        return c

    def _packed(self) -> dict:
        """Structure-of-arrays image of the fields the query path reads
        (jb_molecule_batch, include/jess_b200.h); built once, cached."""
        p = self._pack
        if p is None or p["n"] != len(self._atoms):
            a = self._atoms
            p = {
                "n": len(a),
                "x": np.ascontiguousarray(a["x"][:, 0]),
                "y": np.ascontiguousarray(a["x"][:, 1]),
                "z": np.ascontiguousarray(a["x"][:, 2]),
                "name": np.ascontiguousarray(a["name"]).view("<u4").reshape(-1),
                "resname": np.ascontiguousarray(a["resName"]).view("<u4").reshape(-1),
                "resseq": np.ascontiguousarray(a["resSeq"]),
                "chain": np.ascontiguousarray(a["chainID2"]),
            }
            self._pack = p
        return p


# --- TemplateAtom --------------------------------------------------------------

_AA1 = "ACDEFGHIKLMNPQRSTVWYX"
_AA3 = ["ALA", "CYS", "ASP", "GLU", "PHE", "GLY", "HIS", "ILE", "LYS", "LEU", "MET",
        "ASN", "PRO", "GLN", "ARG", "SER", "THR", "VAL", "TRP", "TYR", "XXX"]
# TessAtom_convertResidue (eff:TessAtom.c:21-56) additionally knows O (PYL) and U (SEC)
_ONE_TO_THREE = dict(zip(_AA1, _AA3))
_ONE_TO_THREE.update({"O": "PYL", "U": "SEC"})


class TemplateAtom:
    """One atom of a template (_jess.pyx:1333-1750, eff:TessAtom.h:36-49)."""

    __slots__ = ("_code", "_resseq", "_chain1", "_chain2", "_pos", "_weight", "_names", "_resnames")

    @classmethod
    def load(cls, file):
        try:
            handle = open(file)
            close = True
        except TypeError:
            handle, close = file, False
        try:
            return cls.loads(handle.read())
        finally:
            if close:
                handle.close()

    @classmethod
    def loads(cls, text):
        """Parse a TESS/Jess template line with the rules of ``TessAtom_create``
        (eff:TessAtom.c:58-234): the record is a PDB atom whose serial column is
        the match mode and whose B-factor column is the distance weight (through a
        ``float``); after column 54 come one-letter alternate residues and
        ``(NAME)`` alternate atom names."""
        line = _as_line(text)
        nul = line.find(b"\0")
        if nul >= 0:
            line = line[:nul]
        rec = np.zeros(1, dtype=ATOM_DTYPE)
        if len(line) < 54 or not _parse_atom_line(line, rec[0]):
            raise ValueError(f"Failed to parse template atom: {text!r}")
        q = line[54:].decode("latin-1")
        names = [rec["name"].tobytes()[:4].split(b"\0")[0]]
        resnames = [rec["resName"].tobytes()[:3].split(b"\0")[0]]
        # pass 1: count and validate nesting
        nest, ac, rc = 0, 1, 1
        for ch in q:
            if ch == "(":
                if nest != 0:
                    raise ValueError(f"Failed to parse template atom: {text!r}")
                nest, ac = 1, ac + 1
            elif ch == ")":
                if nest != 1:
                    raise ValueError(f"Failed to parse template atom: {text!r}")
                nest = 0
            elif nest == 0 and ch in _ONE_TO_THREE:
                rc += 1
        if nest != 0:
            raise ValueError(f"Failed to parse template atom: {text!r}")
        # pass 2: collect
        i = 0
        while (len(resnames) < rc or len(names) < ac) and i < len(q):
            ch = q[i]
            if ch.isspace():
                i += 1
            elif ch == "(":
                i += 1
                m = 0
                while m < 4 and i + m < len(q) and q[i + m] != ")":
                    m += 1
                if i + m >= len(q) or q[i + m] != ")":
                    raise ValueError(f"Failed to parse template atom: {text!r}")
                raw = q[i:i + m].encode("latin-1").ljust(4, b"\0")
                names.append(bytes(0x5F if (c in b" \t\n\v\f\r" or c == 0) else c for c in raw))
                i += m + 1
            else:
                m = 0
                while m <= 4 and i + m < len(q) and not q[i + m].isspace():
                    three = _ONE_TO_THREE.get(q[i + m])
                    if three is None:
                        raise ValueError(f"Failed to parse template atom: {text!r}")
                    resnames.append(three.encode("ascii"))
                    m += 1
                # The reference does NOT step over the codes it has just read (eff:TessAtom.c:236-252 leaves `q`
                # where it is): while names are still owed, the same run of codes is read again, so " E   C"
                # yields GLU, GLU and "W     L0.00" is accepted.  With an atom name still owed it would re-read
                # them for ever, writing past its allocation: that input is refused here.
                if len(names) < ac:
                    raise ValueError(f"Failed to parse template atom: {text!r}")
        atom = cls.__new__(cls)
        atom._code = int(rec["serial"][0])
        atom._resseq = int(rec["resSeq"][0])
        atom._chain1 = int(rec["chainID1"][0])
        atom._chain2 = int(rec["chainID2"][0])
        atom._pos = tuple(float(v) for v in rec["x"][0])
        atom._weight = float(rec["tempFactor"][0])
        atom._names = tuple(n.ljust(4, b"\0")[:4] for n in names)
        atom._resnames = tuple(r.ljust(3, b"\0")[:3] for r in resnames[:rc])
        if atom._code not in MATCH_MODES:
            raise ValueError(f"Invalid match mode: {atom._code!r}")
        return atom

    def __init__(self, *, chain_id: str, residue_number: int, x: float, y: float, z: float,
                 residue_names: Sequence[str], atom_names: Sequence[str],
                 distance_weight: float = 0.0, match_mode: int = 0):
        if match_mode not in MATCH_MODES:
            raise ValueError(f"Invalid match mode: {match_mode!r}")
        if len(chain_id) > 2:
            raise ValueError(f"Invalid chain ID: {chain_id!r}")
        self._code = int(match_mode)
        self._resseq = int(residue_number)
        self._pos = (float(x), float(y), float(z))
        self._weight = float(distance_weight)
        if len(chain_id) == 2:
            self._chain1, self._chain2 = ord(chain_id[0]), ord(chain_id[1])
        elif len(chain_id) == 1:
            self._chain1, self._chain2 = 0x20, ord(chain_id[0])
        else:
            self._chain1, self._chain2 = 0x20, 0x30
        names = []
        for name in atom_names:
            raw = _api_atom_name(name)
            if len(raw) <= 3:
                raw.insert(0, 0x5F)
            names.append(_encode_token(bytes(raw), 4))
        resnames = []
        for name in residue_names:
            raw = name.encode("ascii") if isinstance(name, str) else bytes(name)
            if len(raw) > 3:
                raise ValueError(f"Invalid residue name: {name!r}")
            resnames.append(_encode_token(raw, 3))
        self._names = tuple(names)
        self._resnames = tuple(resnames)

    def _state(self) -> dict:
        return {
            "chain_id": self.chain_id,
            "residue_number": self.residue_number,
            "x": self.x,
            "y": self.y,
            "z": self.z,
            "residue_names": self.residue_names,
            "atom_names": self.atom_names,
            "distance_weight": self.distance_weight,
            "match_mode": self.match_mode,
        }

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(f'{k}={v!r}' for k, v in self._state().items())})"

    def __copy__(self):
        return self.copy()

    def __eq__(self, other):
        if not isinstance(other, TemplateAtom):
            return NotImplemented
        return self._state() == other._state()

    def __hash__(self):
        return hash(tuple(self._state().values()))

    def __reduce__(self):
        return functools.partial(type(self), **self._state()), ()

    @property
    def match_mode(self) -> int:
        return self._code

    @property
    def residue_number(self) -> int:
        return self._resseq

    @property
    def chain_id(self) -> str:
        if self._chain1 == 0x20:
            return chr(self._chain2)
        return (chr(self._chain1) + chr(self._chain2)).strip()

    @property
    def x(self) -> float:
        return self._pos[0]

    @property
    def y(self) -> float:
        return self._pos[1]

    @property
    def z(self) -> float:
        return self._pos[2]

    @property
    def atom_names(self):
        return tuple(n.split(b"\0")[0].replace(b"_", b"").decode() for n in self._names)

    @property
    def residue_names(self):
        return tuple(r.split(b"\0")[0].replace(b"_", b"").decode() for r in self._resnames)

    @property
    def distance_weight(self) -> float:
        return self._weight

    def copy(self):
        atom = type(self).__new__(type(self))
        for slot in TemplateAtom.__slots__:
            setattr(atom, slot, getattr(self, slot))
        return atom

    def dumps(self) -> str:
        buf = io.StringIO()
        self.dump(buf)
        return buf.getvalue()

    def dump(self, file) -> None:
        """Same text as the reference's sprintf (_jess.pyx:1663-1705); alternate atom
        names are not written there either."""
        name = self._names[0].replace(b"_", b" ").replace(b"\0", b" ").decode("ascii")
        res = self._resnames[0].replace(b"_", b" ").replace(b"\0", b" ").decode("ascii")
        head = "ATOM  %5i %4s %3s%c%c%4i    %8.3f%8.3f%8.3f" % (
            self._code, name, res, chr(self._chain1), chr(self._chain2), self._resseq, *self._pos)
        three_to_one = dict(zip(_AA3, _AA1))
        alts = [" "] * 7
        for k, r in enumerate(self._resnames[1:7], start=1):
            alts[k] = three_to_one.get(r.decode("ascii"), "X")
        file.write(head + "".join(alts) + "%4.2f" % self._weight)

    def transform(self, matrix):
        m = _as_matrix(matrix)
        out = self.copy()
        p = _transform_points(np.array([self._pos], dtype=np.float64), m)[0]
        out._pos = (float(p[0]), float(p[1]), float(p[2]))
        return out

    # token views used by the spec tables
    def _name_tokens(self) -> np.ndarray:
        return np.frombuffer(b"".join(n.ljust(4, b"\0") for n in self._names), dtype="<u4")

    def _resname_tokens(self) -> np.ndarray:
        return np.frombuffer(b"".join(r.ljust(4, b"\0") for r in self._resnames), dtype="<u4")


# --- Template ------------------------------------------------------------------

class Template:
    """A template: n `TemplateAtom` + the n x n FP64 distance matrix
    (_jess.pyx:1753-2087, eff:TessTemplate.h:34-41)."""

    __slots__ = ("_atoms", "_id", "_dist", "_dim", "__weakref__")

    @classmethod
    def loads(cls, text, id: Optional[str] = None):
        return cls.load(io.StringIO(text), id=id)

    @classmethod
    def load(cls, file, id: Optional[str] = None):
        atoms = []
        try:
            handle = open(file)
            close = True
        except TypeError:
            handle, close = file, False
        try:
            for line in handle:
                if line.startswith("ATOM"):
                    atoms.append(TemplateAtom.loads(line))
                elif id is None and line.startswith("REMARK PDB_ID"):
                    id = line.split(" ", maxsplit=2)[2].strip()
        finally:
            if close:
                handle.close()
        return cls(atoms, id=id)

    def __init__(self, atoms: Iterable[TemplateAtom] = (), id: Optional[str] = None):
        atoms = [a.copy() for a in atoms]
        for a in atoms:
            if not isinstance(a, TemplateAtom):
                raise TypeError(f"expected TemplateAtom, found {type(a).__name__}")
        self._atoms = tuple(atoms)
        self._id = id
        n = len(atoms)
        # TessAtom_distance (eff:TessAtom.c:494-509): sqrt of the FP64 sum over
        # x, y, z in that order; computed for i<j and mirrored (_jess.pyx:1895-1901)
        pos = np.array([a._pos for a in atoms], dtype=np.float64).reshape(n, 3)
        dist = np.zeros((n, n), dtype=np.float64)
        for i in range(n):
            for j in range(i + 1, n):
                s = 0.0
                for k in range(3):
                    t = float(pos[i, k]) - float(pos[j, k])
                    s += t * t
                d = float(np.sqrt(np.float64(s)))
                dist[i, j] = dist[j, i] = d
        self._dist = dist
        # Template.dimension: distinct (resSeq, chainID1, chainID2), _jess.pyx:1903-1912
        self._dim = len({(a._resseq, a._chain1, a._chain2) for a in atoms})

    def __copy__(self):
        return self.copy()

    def __len__(self) -> int:
        return len(self._atoms)

    def __getitem__(self, index):
        if isinstance(index, slice):
            return type(self)(self._atoms[index], id=self._id)
        n = len(self._atoms)
        i = int(index)
        if i < 0:
            i += n
        if i < 0 or i >= n:
            raise IndexError(index)
        return self._atoms[i]

    def __iter__(self):
        return iter(self._atoms)

    def __eq__(self, other):
        if not isinstance(other, Template):
            return NotImplemented
        return self._id == other._id and len(self) == len(other) and all(a == b for a, b in zip(self, other))

    def __hash__(self):
        return hash((self._id, *(hash(a) for a in self._atoms)))

    def __reduce__(self):
        return functools.partial(type(self), id=self._id), (list(self._atoms),)

    @property
    def id(self):
        return self._id

    @property
    def dimension(self) -> int:
        return self._dim

    def copy(self):
        return type(self)(self._atoms, id=self._id)

    def dumps(self) -> str:
        buf = io.StringIO()
        self.dump(buf)
        return buf.getvalue()

    def dump(self, file, write_id: bool = True) -> None:
        if write_id and self._id is not None:
            file.write(f"REMARK PDB_ID {self._id}\n")
        for a in self._atoms:
            a.dump(file)
            file.write("\n")

    def transform(self, matrix):
        return type(self)([a.transform(matrix) for a in self._atoms], id=self._id)

    def _log_evalue(self, rmsd: float, n_atoms: int) -> float:
        """``TessTemplate_logE`` (eff:TessTemplate.c:127-139)."""
        import math
        return -8.5 + 2.5 * rmsd + -0.7 * float(self._dim) + math.log(float(n_atoms))
