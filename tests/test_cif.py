"""mmCIF ingest without gemmi (pyjess_b200/_cif.py) against the reference's OWN mmCIF tests
(src/pyjess/tests/test_molecule.py:254-397, skipped upstream when gemmi is missing): same files
(tests/data/1AMY.cif / 1AMY.pdb from the reference's test data), same assertions -- the CIF
molecule equals the PDB molecule atom by atom -- plus tokenizer cases (quotes, text fields,
comments, pairs instead of loops, missing columns, several blocks)."""
import datetime
import io
import itertools
import os
import textwrap
import warnings

import numpy as np
import pytest

from pyjess_b200 import Molecule
from pyjess_b200 import _cif

from _util import DATA

PDB, CIF = os.path.join(DATA, "1AMY.pdb"), os.path.join(DATA, "1AMY.cif")


def assert_atom_equal(a1, a2, hetero=False):
    """assertAtomEqual / assertHeteroAtomEqual (test_molecule.py:46-77)."""
    if not hetero:
        assert a1.serial == a2.serial
        assert a1.chain_id == a2.chain_id
        assert a1.residue_number == a2.residue_number
    assert a1.name == a2.name
    assert a1.altloc == a2.altloc
    assert a1.residue_name == a2.residue_name
    assert a1.insertion_code == a2.insertion_code
    assert a1.element == a2.element
    for f in ("x", "y", "z", "occupancy", "temperature_factor", "charge"):
        assert getattr(a1, f) == pytest.approx(getattr(a2, f), abs=1e-5)


def _load_cif(**kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with open(CIF) as f:
            return Molecule.load(f, format="cif", **kw)


def test_cif_dump():
    """test_molecule.py:255-277"""
    with open(CIF) as f:
        text = f.read()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cif_molecule = Molecule.loads(text, format="cif")
    expected = textwrap.dedent(
        """
        HEADER    HYDROLASE (O-GLYCOSYL)                  13-MAY-95   1AMY
        ATOM      1  N   GLN A   1       6.240  48.686  17.460  1.00 27.79           N
        ATOM      2  CA  GLN A   1       5.440  49.851  17.773  1.00 16.62           C
        ATOM      3  C   GLN A   1       6.628  50.721  18.086  1.00 14.24           C
        ATOM      4  O   GLN A   1       7.313  50.396  19.052  1.00 11.61           O
        ATOM      5  CB  GLN A   1       4.588  49.715  19.030  1.00 18.54           C
        """
    ).strip()
    assert cif_molecule[0:5].dumps().strip() == expected


def test_cif_vs_pdb():
    """test_molecule.py:280-303"""
    cif_molecule = _load_cif()
    with open(PDB) as f:
        pdb_molecule = Molecule.loads(f.read(), format="pdb")
    first5_pdb, first5_cif = pdb_molecule[0:5], cif_molecule[0:5]
    assert first5_pdb == first5_cif
    assert first5_pdb.dumps(write_header=False).strip() == first5_cif.dumps(write_header=False).strip()


@pytest.mark.parametrize("via", ["load", "loads"])
def test_load_mmcif_comment(via):
    """test_molecule.py:306-343: comment lines before data_, explicit and detected format."""
    buffer = io.StringIO()
    buffer.write("# this is a comment line \n")
    buffer.write("# and a second comment line \n")
    with open(CIF) as f:
        buffer.write(f.read())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for fmt in ("cif", "detect"):
            if via == "load":
                buffer.seek(0)
                mol = Molecule.load(buffer, format=fmt)
            else:
                mol = Molecule.loads(buffer.getvalue(), format=fmt)
            assert len(mol) == 3339


@pytest.mark.parametrize("use_author", [False, True])
def test_load_consistency_no_skip_hetatm(use_author):
    """test_molecule.py:346-372"""
    pdb_molecule = Molecule.load(PDB, format="pdb")
    cif_molecule = _load_cif(use_author=use_author)
    assert len(cif_molecule) == len(pdb_molecule) == 3339
    for pdb_atom, cif_atom in itertools.islice(zip(pdb_molecule, cif_molecule), 3184):
        assert_atom_equal(pdb_atom, cif_atom)
    for pdb_atom, cif_atom in itertools.islice(zip(pdb_molecule, cif_molecule), 3184, 3339):
        assert_atom_equal(pdb_atom, cif_atom, hetero=True)


@pytest.mark.parametrize("use_author", [False, True])
def test_load_consistency_skip_hetatm(use_author):
    """test_molecule.py:375-397"""
    pdb_molecule = Molecule.load(PDB, format="pdb", skip_hetatm=True)
    cif_molecule = _load_cif(skip_hetatm=True, use_author=use_author)
    assert len(cif_molecule) == len(pdb_molecule) == 3184
    for pdb_atom, cif_atom in zip(pdb_molecule, cif_molecule):
        assert_atom_equal(pdb_atom, cif_atom)


def test_hetatm_without_residue_number_warns_and_numbers_like_the_reference():
    """_io.py:185-199: label_seq_id '.' on a HETATM row -> UserWarning, the running maximum as
    residue number (then incremented); use_author reads auth_seq_id instead (no warning)."""
    with pytest.warns(UserWarning, match="HETATM line found without residue number"):
        with open(CIF) as f:
            mol = Molecule.load(f, format="cif")
    het = [a for a in mol][3184:]
    assert [a.residue_number for a in het[:3]] == [403, 404, 405]          # max label_seq_id is 403
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        with open(CIF) as f:
            auth = Molecule.load(f, format="cif", use_author=True)
    assert [a.residue_number for a in [a for a in auth][3184:3187]] == [500, 501, 502]
    assert mol.id == "1AMY" and mol.name == "HYDROLASE (O-GLYCOSYL)"
    assert mol.date == datetime.datetime(1995, 5, 13)                       # earliest revision_date


MINI = """\
# comment
data_mini
_entry.id 'MINI'
_struct.title
;A title in a
text field
;
loop_
_atom_site.group_PDB
_atom_site.id
_atom_site.type_symbol
_atom_site.label_atom_id
_atom_site.label_alt_id
_atom_site.label_comp_id
_atom_site.label_asym_id
_atom_site.label_seq_id
_atom_site.Cartn_x
_atom_site.Cartn_y
_atom_site.Cartn_z
_atom_site.occupancy
_atom_site.B_iso_or_equiv
_atom_site.pdbx_PDB_model_num
_atom_site.note
ATOM 1 N N . GLY A 1 1.0 2.0 3.0 1.00 10.0 1 'two words' # trailing comment
ATOM 2 C CA A GLY A 1 1.5 2.5 3.5 . . 1 'it's quoted'
HETATM 3 O "O5'" . GLY A 2 4.0 5.0
6.0 0.50 20.0 1
;text
field
;
ATOM 4 N N . GLY A 1 1.0 2.0 3.0 1.00 10.0 2 x#not-a-comment
"""


def test_tokenizer_cases():
    mol = Molecule.loads(MINI)                       # detected: comment line, then data_
    assert mol.id == "mini" and mol.date is None
    assert mol.name == ";A title in a\ntext field\n;"            # raw value, like gemmi's find_value
    atoms = list(mol)
    assert len(atoms) == 3                            # model 2 stops the scan
    assert [a.name for a in atoms] == ["N", "CA", "O5'"]
    assert atoms[1].altloc == "A" and atoms[0].altloc == " "
    assert atoms[1].occupancy == 0.0 and atoms[1].temperature_factor == 0.0
    assert (atoms[2].x, atoms[2].y, atoms[2].z) == (4.0, 5.0, 6.0)
    assert atoms[0].insertion_code == " " and atoms[0].charge == 0          # optional columns absent
    assert len(Molecule.loads(MINI, ignore_endmdl=True)) == 4
    assert len(Molecule.loads(MINI, skip_hetatm=True)) == 2
    blocks = _cif.read_blocks(MINI)
    assert blocks[0].find_value("_entry.id") == "'MINI'"
    assert blocks[0].find("_atom_site.", ["id", "?nope"])[0] == ["1", None]
    assert [r[0] for r in blocks[0].find("_atom_site.", ["note"])] == ["'two words'", "'it's quoted'", ";text\nfield\n;",
                                                                       "x#not-a-comment"]


def test_errors():
    with pytest.raises(ValueError, match="missing columns"):
        Molecule.loads(MINI.replace("_atom_site.Cartn_z\n", "_atom_site.Cartn_zz\n"), format="cif")
    with pytest.raises(ValueError, match="missing columns"):
        Molecule.loads("data_empty\n_entry.id X\n", format="cif")
    with pytest.raises(RuntimeError, match="single data block"):
        Molecule.loads(MINI + "data_second\n_entry.id Y\n", format="cif")
    with pytest.raises(ValueError, match="CIF syntax error"):
        Molecule.loads("data_x\nloop_\n_a.b\n_a.c\n1 2 3\n", format="cif")
    with pytest.raises(ValueError, match="invalid value for `format`"):
        Molecule.loads(MINI, format="xyz")
    with pytest.raises(ValueError, match="mmCIF data tags found"):
        Molecule.loads(MINI, format="pdb")


@pytest.mark.parametrize("kw", [{}, {"skip_hetatm": True}, {"ignore_endmdl": True}])
@pytest.mark.parametrize("columns", ["_PRIMARY", "_AUTHOR"])
def test_column_builder_equals_the_reference_row_loop(kw, columns):
    """`_records_from_columns` (what the loader runs) == `_records_from_rows` (the reference's loop
    as written, one `Atom` per row), byte for byte, on 1AMY.cif and on the tokenizer example."""
    for text in (open(CIF).read(), MINI):
        block = _cif.read_blocks(text)[0]
        cols = getattr(_cif, columns)
        if block.find("_atom_site.", cols) is None:
            continue                                   # MINI has no auth_* columns
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            fast = _cif._records_from_columns(block.find_columns("_atom_site.", cols), **kw)
            slow = _cif._records_from_rows(block.find("_atom_site.", cols), **kw)
        assert len(fast) == len(slow) > 0 and fast.tobytes() == slow.tobytes()


def test_line_tokenizer_equals_the_grammar():
    """The line-wise fast path of `_tokens` against the token grammar applied to the whole text."""
    for text in (open(CIF).read(), MINI, MINI.replace("\n", "\r\n")):
        want = [t for t in _cif._TOKEN.findall(text.replace("\r\n", "\n")) if not t.startswith("#")]
        assert _cif._tokens(text) == want
    with pytest.raises(ValueError, match="unterminated text field"):
        _cif._tokens("data_x\n_a.b\n;never closed\n")


def test_invalid_fields_raise_like_atom():
    with pytest.raises(ValueError, match="Invalid residue name"):
        Molecule.loads(MINI.replace(" GLY A 2 ", " GLYX A 2 "), format="cif")
    with pytest.raises(ValueError):
        Molecule.loads(MINI.replace("ATOM 1 N N . GLY A 1 ", "ATOM 1 N N . GLY A . "), format="cif")   # int(".")


# ---- the native one-pass reader (jb_cif_parse) against the generic Python loader --------------

def _both(text, **kw):
    """(info, native records, records of the generic loader or the exception it raises)."""
    from pyjess_b200 import _capi
    from pyjess_b200._model import ATOM_DTYPE
    data = text.encode("utf-8")
    info, native = _capi.cif_parse(data, ATOM_DTYPE, **kw)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            blocks = _cif.read_blocks(text)
            if len(blocks) != 1:
                raise RuntimeError("single data block expected")
            cols = blocks[0].find_columns("_atom_site.", _cif._AUTHOR if kw.get("use_author") else _cif._PRIMARY)
            if not cols:
                raise ValueError("missing columns in CIF files")
            generic = _cif._records_from_columns(cols, skip_hetatm=kw.get("skip_hetatm", False),
                                                 ignore_endmdl=kw.get("ignore_endmdl", False))
    except Exception as err:  # noqa: BLE001
        generic = err
    return info, native, generic


def _synthetic_cif(seed, n=400):
    """An _atom_site loop with every field shape the reader distinguishes: two-character chains,
    negative / zero-padded / signed numbers, exponents, quoted names, alternate locations,
    insertion codes, charges, '.' and '?' place holders, HETATM rows with and without numbers,
    a second model, comments and rows broken over two lines."""
    rng = np.random.default_rng(seed)
    tags = ["group_PDB", "id", "type_symbol", "label_atom_id", "label_alt_id", "label_comp_id", "label_asym_id",
            "label_entity_id", "label_seq_id", "pdbx_PDB_ins_code", "Cartn_x", "Cartn_y", "Cartn_z", "occupancy",
            "B_iso_or_equiv", "pdbx_formal_charge", "auth_seq_id", "auth_comp_id", "auth_asym_id", "auth_atom_id",
            "pdbx_PDB_model_num"]
    out = ["# made by tests/test_cif.py", "data_SYN%d" % seed, "_entry.id SYN", "loop_"] + ["_atom_site." + t for t in tags]
    names = ["N", "CA", "C", "O", "CB", '"O5\'"', '"C1\'"', "OXT", "HG21", "1HB"]
    for i in range(n):
        het = rng.random() < 0.15
        model = "1" if i < n - 30 else "2"
        seq = "." if het and rng.random() < 0.5 else rng.choice(["%d", "%03d", "+%d", "-%d"]) % int(rng.integers(0, 900))
        x, y, z = rng.uniform(-99, 99, 3)
        fx = rng.choice(["%.3f", "%.5f", "%.1e", "%+.2f", "%d."])
        row = ["HETATM" if het else "ATOM", str(i + 1 if rng.random() < 0.9 else int(rng.integers(1, 2 ** 31 - 1))),
               str(rng.choice(["C", "N", "O", "S", "CA", "FE"])), str(rng.choice(names)), str(rng.choice([".", ".", "A", "B"])),
               str(rng.choice(["ALA", "GLY", "HOH", "CA", "A"])), str(rng.choice(["A", "B", "AA", "A1", "z"])), "1", seq,
               str(rng.choice(["?", "?", "A", "P"])), fx % x if "d" not in fx else fx % int(x), "%.3f" % y, "%.3f" % z,
               str(rng.choice(["1.00", "0.50", ".", "1"])), str(rng.choice(["20.11", ".", "0.0", "1.5e1"])),
               str(rng.choice(["?", "?", "0", "-1", "+2"])), str(int(rng.integers(-50, 9000))), "ALA", str(rng.choice(["A", "B"])),
               str(rng.choice(["N", "CA", "O"])), model]
        if rng.random() < 0.1:
            out.append(" ".join(row[:9]))
            out.append("  " + " ".join(row[9:]) + "   # wrapped row")
        else:
            out.append((" " * int(rng.integers(1, 4))).join(row))
    out += ["#", "loop_", "_pdbx_audit_revision_history.ordinal", "_pdbx_audit_revision_history.revision_date",
            "1 2001-02-03", "2 1999-12-31", "_struct.title 'a synthetic; entry'", ""]
    return "\n".join(out)


@pytest.mark.parametrize("kw", [{}, {"skip_hetatm": True}, {"ignore_endmdl": True}, {"use_author": True}])
@pytest.mark.parametrize("seed", range(4))
def test_native_reader_equals_the_generic_loader(seed, kw):
    for text in (_synthetic_cif(seed), _synthetic_cif(seed).replace("\n", "\r\n"), MINI, open(CIF).read()):
        if kw.get("use_author") and text is MINI:
            continue
        info, native, generic = _both(text, **kw)
        assert info.fallback == 0 and not isinstance(generic, Exception), (info.fallback, generic)
        assert len(native) == len(generic) > 0 and native.tobytes() == generic.tobytes()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mol = Molecule.loads(_synthetic_cif(seed), **kw)           # end to end: metadata next to the native table
    assert mol.id == "SYN%d" % seed and mol.name == "a synthetic; entry" and mol.date == datetime.datetime(1999, 12, 31)


@pytest.mark.parametrize("edit,code", [
    (("ATOM 1 N N . GLY A 1 ", "ATOM 1 N N . GLY A 1_0 "), 4),                 # int() accepts underscores: Python decides
    (("ATOM 1 N N . GLY A 1 ", "ATOM 1 N N . GLY A . "), 4),                   # ATOM row without a number: int('.') raises
    ((" 1.0 2.0 3.0 1.00 10.0 1 'two words'", " nan 2.0 3.0 1.00 10.0 1 'two words'"), 4),
    (("ATOM 2 C CA A GLY", "ATOM 2 C CALPHA A GLY"), 4),                       # Atom() rejects the name
    (("ATOM 2 C CA A GLY", "ATOM 2 C CA AB GLY"), 4),                          # alternate location of two characters
    (("_atom_site.Cartn_z\n", "_atom_site.Cartn_zz\n"), 3),                    # missing column
    (("_atom_site.id\n", "_atom_site.id\n_atom_site.id\n"), 3),                # duplicate tag (and a ragged loop)
    (("data_mini", "data_mini\ndata_second"), 2),
    (("# comment\ndata_mini", "_stray.tag 1\ndata_mini"), 1),
    ((";A title in a", ";A title in a\n;\n;never closed"), 1),
])
def test_native_reader_leaves_the_hard_cases_to_python(edit, code):
    text = MINI.replace(*edit)
    assert text != MINI
    info, native, generic = _both(text)
    assert info.fallback == code
    # ... and the public loader answers like the generic one (same exception type or same records)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = Molecule.loads(text, format="cif")._atoms
    except Exception as err:  # noqa: BLE001
        got = err
    if isinstance(generic, Exception):
        assert type(got) is type(generic) or isinstance(got, (ValueError, RuntimeError))
    else:
        assert got.tobytes() == generic.tobytes()


def test_native_reader_capacity_retry_and_warning_count():
    from pyjess_b200 import _capi
    from pyjess_b200._model import ATOM_DTYPE
    loop = MINI[MINI.index("loop_"):MINI.index("HETATM 3")].replace("# trailing comment", "")
    text = "data_mini\n_entry.id X\n" + " ".join(loop.split("\n")) * 1 + " ATOM 3 O O . GLY A 2 4.0 5.0 6.0 0.50 20.0 1 z"
    assert "\nATOM" not in text                                               # every row on one line: nothing to count
    info, native = _capi.cif_parse(text.encode(), ATOM_DTYPE)
    assert info.fallback == 0 and len(native) == 3
    with pytest.warns(UserWarning) as caught:
        warnings.simplefilter("always")
        mol = Molecule.load(CIF)
    assert len(mol) == 3339 and len([w for w in caught if "HETATM line found" in str(w.message)]) == 155


def test_native_reader_fuzz_never_disagrees_with_the_generic_loader():
    """Random edits (replace / delete / insert of syntax-relevant characters) of three CIF documents:
    the native reader either hands the file over (`fallback`) or returns exactly the records of the
    generic loader -- it never accepts what the generic loader rejects, and never crashes."""
    rng = np.random.default_rng(2026)
    bases = [MINI, _synthetic_cif(1, 40), _synthetic_cif(2, 25).replace("\n", "\r\n")]
    alphabet = list(" \n\t#;'\"._?-+0123456789eEATOMHLYGdatloop_")
    accepted = 0
    for it in range(1500):
        s = list(bases[it % len(bases)])
        for _ in range(int(rng.integers(1, 4))):
            op, pos = int(rng.integers(0, 3)), int(rng.integers(0, len(s)))
            if op == 0:
                s[pos] = alphabet[int(rng.integers(len(alphabet)))]
            elif op == 1:
                del s[pos]
            else:
                s.insert(pos, alphabet[int(rng.integers(len(alphabet)))])
        text = "".join(s)
        kw = [{}, {"skip_hetatm": True}, {"ignore_endmdl": True}][it % 3]
        info, native, generic = _both(text, **kw)
        if info.fallback == 0:
            accepted += 1
            assert not isinstance(generic, Exception), (text, generic)
            assert native.tobytes() == generic.tobytes(), text
    assert accepted > 300
