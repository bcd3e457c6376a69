"""Structure ingest: the native one-pass PDB parser `jb_pdb_parse`
(csrc/ingest.cu, host code of the engine library -- needs no GPU) against the line-by-line
Python restatement of PDBMoleculeParser.load + Atom_parse, and against the atoms the
REFERENCE parsed (tests/golden/cfg2_structures.npz was written by the reference's loader)."""
import io
import os

import numpy as np
import pytest

from _util import DATA, cfg2_molecules
from pyjess_b200 import Molecule
from pyjess_b200._model import ATOM_DTYPE

FILES = ["1AMY.pdb", "pdb1lnb.ent", "pdb1a0p.ent", "5ayx.EF.pdb"]


def both(data: bytes, **kw):
    native = Molecule.loads(data, **kw)
    python = Molecule._parse_pdb_python(data, **kw)
    return native, python


@pytest.mark.parametrize("name", FILES)
def test_native_parser_is_byte_identical_on_fixtures(name):
    with open(os.path.join(DATA, name), "rb") as f:
        data = f.read()
    for kw in ({}, {"skip_hetatm": True}, {"ignore_endmdl": True}):
        a, b = both(data, **kw)
        assert len(a) == len(b) > 0
        assert a._atoms.tobytes() == b._atoms.tobytes()
        assert (a.id, a.name, a.date) == (b.id, b.name, b.date)


def test_native_parser_matches_reference_parsed_atoms():
    ref = {m.id.upper(): m for m in cfg2_molecules()}
    for name, key in (("pdb1a0p.ent", "1A0P"), ("pdb1lnb.ent", "1LNB")):
        ours, theirs = Molecule.load(os.path.join(DATA, name)), ref[key]
        assert len(ours) == len(theirs)
        for field in ("x", "serial", "name", "resName", "resSeq", "chainID1", "chainID2"):
            assert np.array_equal(ours._atoms[field], theirs._atoms[field]), field


ODD_LINES = [
    b"ATOM      1  N   MET A   1      27.340  24.430   2.614  1.00  9.67           N  \n",
    b"ATOM      2  CA  MET A   1      26.266  25.413   2.842  1.00 10.38           C1+\r\n",     # CRLF + charge
    b"HETATM    3 ZN    ZN A 317      1.5e1   -.5     +3.     0.50 20.00      SEG1ZN\n",       # exponent, bare dot
    b"ATOM      4  CB  MET\n",                                                                    # short line
    b"ATOM  \n",
    b"ATOM",                                                                                      # no newline at EOF
]


def test_native_parser_semantics_on_odd_lines():
    data = b"HEADER    HYDROLASE                               01-JAN-99   1ABC\n" + b"".join(ODD_LINES[:-1]) + \
           b"REMARK nothing\nATOM      5  C   GLY B  22A   -100.123-200.456 300.789\n" \
           b"ATOM  12345678901234567890\n" \
           b"ATOM      6  O\tGLY\tB\t 23      1.0     2.0     3.0\n" \
           b"ATOM      7  X\x00 GLY B  24      1.0     2.0     3.0\n" \
           b"ATOM      8  C   GLY B  25  1234567.12345678901234567 0x1p3  nan  inf\n" \
           b"ENDMDL\nATOM      9  N   ALA A   2       0.0     0.0     0.0\n" + ODD_LINES[-1]
    for kw in ({}, {"skip_hetatm": True}, {"ignore_endmdl": True}):
        a, b = both(data, **kw)
        assert len(a) == len(b)
        for field in ATOM_DTYPE.names:
            x, y = a._atoms[field], b._atoms[field]
            same = np.array_equal(x, y, equal_nan=True) if x.dtype.kind == "f" else np.array_equal(x, y)
            assert same, (kw, field, x, y)
        assert (a.id, a.name, a.date) == (b.id, b.name, b.date) == ("1ABC", "HYDROLASE", a.date)
    full = Molecule.loads(data, ignore_endmdl=True)
    assert len(full) == len(Molecule.loads(data)) + 2


def test_native_parser_rejects_mmcif_and_handles_empty():
    with pytest.raises(ValueError):
        Molecule.loads(b"data_1ABC\nloop_\n_atom_site.id\n")
    assert len(Molecule.loads(b"")) == 0
    assert len(Molecule.load(io.BytesIO(b"REMARK only\n"))) == 0
    assert len(Molecule.load(io.StringIO("ATOM      1  N   MET A   1      27.340  24.430   2.614\n"))) == 1


def test_native_packing_matches_numpy_packing():
    """jb_molecules_pack (csrc/ingest.cu) fills the SoA staging image exactly like slicing the
    molecules' record arrays field by field (`Molecule._packed`)."""
    from pyjess_b200._tables import MoleculeBatch
    mols = [Molecule.load(os.path.join(DATA, f)) for f in FILES] + [Molecule([])]
    for kw in ({}, {"serial": True}, {"serial": True, "pool": {}}):
        b = MoleculeBatch(mols, **kw)
        assert b.atom_offset.tolist() == np.cumsum([0] + [len(m) for m in mols]).tolist()
        for field, key in (("x", "x"), ("y", "y"), ("z", "z"), ("name", "name"), ("resname", "resname"),
                           ("resseq", "resseq"), ("chain", "chain")):
            ref = np.concatenate([m._packed()[key] for m in mols])
            assert np.array_equal(getattr(b, field)[:b.n_atoms], ref), field
        if kw.get("serial"):
            assert np.array_equal(b.serial[:b.n_atoms], np.concatenate([m._atoms["serial"] for m in mols]))
        else:
            assert b.serial is None


def test_native_atom_formatter_matches_the_reference_sprintf():
    """jb_format_atoms_pdb against the Python restatement of the reference's sprintf
    (`_format_atom`, pinned to the reference's own Hit.dump text by tests/test_api_*.py)."""
    from pyjess_b200 import _capi
    from pyjess_b200._model import _format_atom
    recs = Molecule.load(os.path.join(DATA, "1AMY.pdb"))._atoms[:300].copy()
    recs["altLoc"][5] = 0
    recs["occupancy"][7] = 0.005
    recs["tempFactor"][9] = 99.995
    recs["serial"][11] = -42
    xyz = recs["x"] * 1.000001 + 0.0005
    assert _capi.format_atoms_pdb(recs, xyz) == "".join(_format_atom(recs[k], xyz[k]) for k in range(len(recs)))
    assert _capi.format_atoms_pdb(recs) == "".join(_format_atom(recs[k], recs["x"][k]) for k in range(len(recs)))
    assert _capi.format_atoms_pdb(recs[:0]) == ""


def test_native_atom_formatter_rounding_and_overflow_cases():
    """The field-by-field fixed-point writer of jb_format_atoms_pdb against printf's exact
    half-to-even rounding (Python's % operator, as `_format_atom`): exact decimal ties (k/16
    and k/8), products next to a tie, negative zero, values overflowing their fields,
    non-finite numbers, and a large random block."""
    from pyjess_b200 import _capi
    from pyjess_b200._model import ATOM_DTYPE, _format_atom
    rng = np.random.default_rng(7)
    special = np.array([0.0625, 0.1875, -0.0625, 2.5e-4, 0.0005, 0.0015, 0.0025, -0.0005, 1.0005, 2.0015, 1234.5675,
                        -1234.5675, 0.125, 0.375, 8.3125, -0.0, -1e-9, 1e-9, 0.9995, 9999.9995, -999.9995, 99999.9999,
                        1e7, -1e7, 1e9, 1.5e9, 1e15, 1e22, 1e300, -1e300, np.nan, np.inf, -np.inf, 5e-324,
                        0.4999999999, 0.0004999999999, 0.00050000000001, 123.4565, 123.45650000000001, 123.45649999999999])
    special = np.concatenate([special, rng.integers(-40000, 40000, 400) / 16.0, rng.integers(-40000, 40000, 400) / 8.0,
                              np.nextafter(rng.integers(1, 90000, 300) / 2000.0, 1e9),
                              np.nextafter(rng.integers(1, 90000, 300) / 2000.0, -1e9)])
    n = len(special) + 20000
    recs = np.zeros(n, dtype=ATOM_DTYPE)
    recs["name"], recs["resName"], recs["element"] = b"_CA_", b"SER", b"_C"
    recs["chainID1"], recs["chainID2"], recs["altLoc"], recs["iCode"] = 0x20, 0x41, 0x20, 0x20
    recs["serial"] = rng.integers(-200000, 2000000, n)
    recs["serial"][:4] = (0, -2147483648, 2147483647, 99999)
    recs["resSeq"] = rng.integers(-2000, 20000, n)
    xyz = rng.uniform(-500, 500, (n, 3))
    xyz[len(special):len(special) + 5000] = np.round(xyz[len(special):len(special) + 5000], 3)   # PDB precision
    xyz[:len(special), 0] = special
    xyz[:len(special), 1] = special[::-1]
    recs["x"] = xyz
    recs["occupancy"] = rng.uniform(0, 1, n).astype(np.float32)
    recs["tempFactor"] = rng.uniform(0, 120, n).astype(np.float32)
    with np.errstate(over="ignore"):
        recs["occupancy"][:len(special)] = special.astype(np.float32)
    recs["tempFactor"][:200] = (rng.integers(0, 8000, 200) / 8.0 + 0.005).astype(np.float32)
    got = _capi.format_atoms_pdb(recs)
    want = "".join(_format_atom(recs[k], recs["x"][k]) for k in range(n))
    if got != want:
        g, w = got.splitlines(), want.splitlines()
        bad = [k for k in range(min(len(g), len(w))) if g[k] != w[k]][:3]
        raise AssertionError("\n".join(f"{k}: {g[k]!r} != {w[k]!r}" for k in bad) or "line counts differ")
    # a block that ends inside the 2 KiB reserve of the output buffer goes through the bounded path
    small = recs[len(special):len(special) + 3]
    assert _capi.format_atoms_pdb(small) == "".join(_format_atom(small[k], small["x"][k]) for k in range(3))


def test_load_many_equals_load_in_input_order():
    """`Molecule.load_many` (threads over the GIL-free native parsers) == `Molecule.load` per file,
    in input order, PDB and mmCIF mixed, whatever the number of threads; errors propagate."""
    import warnings
    names = ["1AMY.pdb", "pdb1lnb.ent", "1AMY.cif", "5ayx.EF.pdb", "pdb1a0p.ent"] * 5
    files = [os.path.join(DATA, n) for n in names]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = [Molecule.load(f, skip_hetatm=True) for f in files]
        for jobs in (1, 3, 64, None):
            got = Molecule.load_many(files, jobs=jobs, skip_hetatm=True)
            assert len(got) == len(want)
            for a, b in zip(got, want):
                assert a.id == b.id and a.name == b.name and a.date == b.date
                assert a._atoms.tobytes() == b._atoms.tobytes()
        named = Molecule.load_many(files[:4], jobs=2, ids=["a", None, "c", "d"])
        assert [m.id for m in named] == ["a", want[1].id, "c", "d"]
        with pytest.raises(ValueError):
            Molecule.load_many(files[:2], ids=["a"])
        assert Molecule.load_many([]) == []
        with pytest.raises(FileNotFoundError):
            Molecule.load_many(files[:3] + [os.path.join(DATA, "missing.pdb")] + files[3:], jobs=4)


def test_pdb_capacity_estimate_is_retried():
    """The record capacity is estimated from the file size (54 characters per ATOM line); files made
    of shorter lines make the parser report what it needs and the load is repeated."""
    text = "ATOM      1  N   GLN A   1\n" * 500 + "HEADER    X\n"
    mol = Molecule.loads(text)
    assert len(mol) == 500 and mol[499].residue_number == 1
    assert mol._atoms.tobytes() == Molecule._parse_pdb_python(text.encode())._atoms.tobytes()


def test_threaded_packing_is_identical(monkeypatch):
    """jb_molecules_pack cuts batches of more than 262,144 atoms into runs of structures, one host
    thread each (JB_PACK_THREADS): same staging image for every thread count, empty structures
    and a thread count above the number of structures included."""
    from pyjess_b200 import synth
    from pyjess_b200._tables import MoleculeBatch
    _, mols, _, _ = synth.workload("cfg3", scale=0.13)
    mols = mols + [Molecule()] + mols[:5]
    assert sum(len(m) for m in mols) > 262144
    ref = None
    for threads in ("1", "2", "7", "64"):
        monkeypatch.setenv("JB_PACK_THREADS", threads)
        b = MoleculeBatch(mols, serial=True)
        cur = [b.atom_offset.copy()] + [np.asarray(getattr(b, k))[:b.n_atoms].copy()
                                        for k in ("x", "y", "z", "name", "resname", "resseq", "chain", "serial")]
        ref = ref or cur
        assert all(np.array_equal(u, v) for u, v in zip(ref, cur)), threads


def test_pdb_loader_fuzz_against_the_line_by_line_restatement():
    """Random edits of a PDB excerpt (blanks, digits, NULs, lone carriage returns, truncated
    lines): the product path (native one-pass parser + header decoding) and the line-by-line
    restatement of the reference's text-mode loader give the same records, id, name and date."""
    with open(os.path.join(DATA, "pdb1lnb.ent")) as f:
        lines = f.read().split("\n")
    base = "\n".join(lines[:3] + [l for l in lines if l.startswith(("ATOM", "HETATM", "ENDMDL", "HEADER"))][:60]) + "\n"
    rng = np.random.default_rng(5)
    alphabet = list(" \n\r\t.-+0123456789eExXATOMHETNDLRabc_\0")

    def outcome(loader, data, **kw):
        try:
            with np.errstate(over="ignore"):
                m = loader(data, **kw)
            return m._atoms.tobytes(), m.id, m.name, m.date
        except Exception as err:  # noqa: BLE001
            return type(err).__name__

    for it in range(500):
        s = list(base)
        for _ in range(int(rng.integers(1, 5))):
            op, pos = int(rng.integers(0, 3)), int(rng.integers(0, len(s)))
            if op == 0:
                s[pos] = alphabet[int(rng.integers(len(alphabet)))]
            elif op == 1:
                del s[pos]
            else:
                s.insert(pos, alphabet[int(rng.integers(len(alphabet)))])
        data = "".join(s).encode()
        kw = [{}, {"skip_hetatm": True}, {"ignore_endmdl": True}][it % 3]
        full = dict(ignore_endmdl=kw.get("ignore_endmdl", False), skip_hetatm=kw.get("skip_hetatm", False))
        got = outcome(lambda d, **k: Molecule._parse_pdb(d, id=None, date=None, name=None, format="pdb", **k), data, **full)
        want = outcome(Molecule._parse_pdb_python, data, **full)
        assert got == want, data
