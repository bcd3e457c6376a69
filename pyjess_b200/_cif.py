"""mmCIF structure ingest without gemmi (second half).

The reference reads mmCIF through `gemmi` (``CIFMoleculeParser``, src/pyjess/_io.py:105-245):
``block.find("_atom_site.", columns)`` over the sole data block, one `Atom` per row, and the
identifier / name / date from ``block.name``, ``_struct_keywords.pdbx_keywords`` /
``_struct.title`` / ``_entry.id`` and the earliest ``_pdbx_audit_revision_history.revision_date``.
This module is that loader over a small CIF 1.1 tokenizer of its own (data blocks, pairs,
loops, quoted strings, ``;`` text fields, comments) that keeps values RAW like gemmi does
(quotes included: the loader strips them exactly where the reference does).

Parity is pinned by the reference's own mmCIF tests (src/pyjess/tests/test_molecule.py:254-397,
skipped upstream when gemmi is absent): the 3,339 atoms of ``1AMY.cif`` equal the atoms of
``1AMY.pdb`` field by field, with and without ``use_author`` / ``skip_hetatm``; leading comment
lines and format detection; the dumped header -- see tests/test_cif.py.
"""
from __future__ import annotations

import datetime
import re
import warnings
from typing import Dict, List, Optional, Tuple

import numpy as np

# one token: a `;` text field (only at the start of a line), a quoted string (the closing quote
# must be followed by white space), a comment, or a run of non-blanks
_TOKEN = re.compile(
    r"""^;(?:.|\n)*?^;          # text field
      | '(?:[^'\n]|'(?=\S))*'(?=\s|\Z)   # single-quoted
      | "(?:[^"\n]|"(?=\S))*"(?=\s|\Z)   # double-quoted
      | \#[^\n]*                # comment
      | [^\s]+                  # bare word
    """, re.M | re.X)


class _Block:
    """One data block: `pairs[tag] = value`, `loops = [(tags, values)]` (row-major, raw tokens)."""

    def __init__(self, name: str):
        self.name = name
        self.pairs: Dict[str, str] = {}
        self.loops: List[Tuple[List[str], List[str]]] = []

    def find(self, prefix: str, columns: List[str]) -> Optional[List[List[Optional[str]]]]:
        """``gemmi.cif.Block.find(prefix, tags)``: rows of the loop (or the pairs, as a single
        row) holding the tags; a leading ``?`` marks an optional column (None when absent);
        `None` when a required column is missing."""
        wanted = [(c[1:], True) if c.startswith("?") else (c, False) for c in columns]
        first = next((prefix + c for c, optional in wanted if not optional), None)
        if first is None:
            return None
        lower = first.lower()
        for tags, values in self.loops:
            index = {t.lower(): i for i, t in enumerate(tags)}
            if lower not in index:
                continue
            pos = []
            for c, optional in wanted:
                i = index.get((prefix + c).lower())
                if i is None and not optional:
                    return None
                pos.append(i)
            width = len(tags)
            return [[None if i is None else values[r + i] for i in pos] for r in range(0, len(values) - width + 1, width)]
        pairs = {t.lower(): v for t, v in self.pairs.items()}
        if lower not in pairs:
            return None
        row = []
        for c, optional in wanted:
            v = pairs.get((prefix + c).lower())
            if v is None and not optional:
                return None
            row.append(v)
        return [row]

    def find_columns(self, prefix: str, columns: List[str]) -> Optional[List[Optional[List[str]]]]:
        """`find`, column-wise: one list of raw values per requested tag (None for an absent
        optional one); no per-row lists."""
        wanted = [(c[1:], True) if c.startswith("?") else (c, False) for c in columns]
        first = next((prefix + c for c, optional in wanted if not optional), None)
        if first is None:
            return None
        for tags, values in self.loops:
            index = {t.lower(): i for i, t in enumerate(tags)}
            if first.lower() not in index:
                continue
            out = []
            for c, optional in wanted:
                i = index.get((prefix + c).lower())
                if i is None and not optional:
                    return None
                out.append(None if i is None else values[i::len(tags)])
            return out
        rows = self.find(prefix, columns)          # pairs: a one-row table
        if not rows:
            return None
        return [None if v is None else [v] for v in rows[0]]

    def find_value(self, tag: str) -> Optional[str]:
        """``Block.find_value``: the value of a pair, or of a one-row loop column."""
        lower = tag.lower()
        for t, v in self.pairs.items():
            if t.lower() == lower:
                return v
        for tags, values in self.loops:
            for i, t in enumerate(tags):
                if t.lower() == lower and len(values) == len(tags):
                    return values[i]
        return None


def _tokens(text: str) -> List[str]:
    """All tokens of the document, comments dropped.  Lines without quotes, `#` or a leading `;`
    (nearly all of an `_atom_site` loop) are cut with `str.split`; the others go through `_TOKEN`."""
    if "\r" in text:
        text = text.replace("\r\n", "\n").replace("\r", "\n")
    out: List[str] = []
    lines = text.split("\n")
    i, n = 0, len(lines)
    while i < n:
        line = lines[i]
        if line.startswith(";"):                       # text field: up to the next line starting with ';'
            j = i + 1
            while j < n and not lines[j].startswith(";"):
                j += 1
            if j >= n:
                raise ValueError("CIF syntax error: unterminated text field")
            out.append("\n".join(lines[i:j]) + "\n;")
            rest = lines[j][1:]
            if rest.strip():
                out.extend(t for t in _TOKEN.findall(rest) if not t.startswith("#"))
            i = j + 1
            continue
        if "'" in line or '"' in line or "#" in line:
            out.extend(t for t in _TOKEN.findall(line) if not t.startswith("#"))
        else:
            out.extend(line.split())
        i += 1
    return out


def read_blocks(text: str) -> List[_Block]:
    """Tokenise a CIF document into its data blocks."""
    blocks: List[_Block] = []
    block: Optional[_Block] = None
    tokens = _tokens(text)
    i, n = 0, len(tokens)
    while i < n:
        tok = tokens[i]
        low = tok[:5].lower()
        if low == "data_":
            block = _Block(tok[5:])
            blocks.append(block)
            i += 1
        elif block is None:
            raise ValueError(f"CIF syntax error: {tok[:20]!r} before any data block")
        elif tok.lower() == "loop_":
            i += 1
            tags = []
            while i < n and tokens[i].startswith("_"):
                tags.append(tokens[i])
                i += 1
            start = i
            while i < n and (tokens[i][0] not in "_dDlLsSgG" or not _is_keyword(tokens[i])):
                i += 1
            if not tags:
                raise ValueError("CIF syntax error: loop_ without tags")
            values = tokens[start:i]
            if len(values) % len(tags):
                raise ValueError(f"CIF syntax error: loop {tags[0]} holds {len(values)} values for {len(tags)} tags")
            block.loops.append((tags, values))
        elif tok.startswith("_"):
            if i + 1 >= n or _is_keyword(tokens[i + 1]):
                raise ValueError(f"CIF syntax error: tag {tok} without a value")
            block.pairs[tok] = tokens[i + 1]
            i += 2
        elif low == "save_":       # save frames (dictionaries): skipped
            i += 1
        else:
            raise ValueError(f"CIF syntax error: unexpected value {tok[:20]!r}")
    return blocks


def _is_keyword(tok: str) -> bool:
    if tok.startswith("_"):
        return True
    low = tok[:5].lower()
    return low in ("data_", "loop_", "save_") or tok.lower() in ("stop_", "global_")


_PRIMARY = ["id", "type_symbol", "label_atom_id", "label_alt_id", "label_comp_id", "label_asym_id", "label_seq_id",
            "?pdbx_PDB_ins_code", "Cartn_x", "Cartn_y", "Cartn_z", "occupancy", "B_iso_or_equiv", "?pdbx_formal_charge",
            "?group_PDB", "pdbx_PDB_model_num"]
_AUTHOR = ["id", "type_symbol", "auth_atom_id", "label_alt_id", "auth_comp_id", "auth_asym_id", "auth_seq_id",
           "?pdbx_PDB_ins_code", "Cartn_x", "Cartn_y", "Cartn_z", "occupancy", "B_iso_or_equiv", "?pdbx_formal_charge",
           "?group_PDB", "pdbx_PDB_model_num"]


def load_molecule(text: str, molecule_type, *, id=None, date=None, name=None, use_author=False, skip_hetatm=False,
                  ignore_endmdl=False):
    """``CIFMoleculeParser._load_block`` (src/pyjess/_io.py:167-245) on `text` (str or bytes).

    Well-formed files take the native one-pass reader of the engine library (`jb_cif_parse`,
    csrc/ingest.cu) for the atom table -- the Python tokenizer then only sees the handful of items behind
    the identifier, name and date (their byte ranges come back with the table); anything else (and every error) goes through the
    generic loader below."""
    arr = None
    if isinstance(text, (bytes, bytearray)):
        data = bytes(text)
        from . import _capi
        from ._model import ATOM_DTYPE
        info, records = _capi.cif_parse(data, ATOM_DTYPE, skip_hetatm=skip_hetatm, ignore_endmdl=ignore_endmdl,
                                        use_author=use_author)
        if not info.fallback:
            arr = records
            for _ in range(info.hetatm_unnumbered):
                warnings.warn(_HETATM_WARNING, UserWarning, stacklevel=4)
            # what is left to read: the block name and the few items behind id / name / date
            pieces = [b"data_" + data[info.name_offset:info.name_offset + info.name_length]]
            pieces += [data[info.meta[k][0]:info.meta[k][1]] for k in range(info.n_meta)]
            text = b"\n".join(pieces).decode("utf-8") + "\n"
        else:
            text = data.decode("utf-8")
    blocks = read_blocks(text)
    if len(blocks) != 1:
        # gemmi: Document.sole_block()
        raise RuntimeError(f"single data block expected, got {len(blocks)}")
    block = blocks[0]
    if arr is None:
        if not block.find_columns("_atom_site.", _AUTHOR if use_author else _PRIMARY):
            raise ValueError("missing columns in CIF files")
        arr = _records_from_columns(block.find_columns("_atom_site.", _AUTHOR if use_author else _PRIMARY),
                                    skip_hetatm=skip_hetatm, ignore_endmdl=ignore_endmdl)
    if id is None:
        id = block.name
    if name is None:
        entry_id = block.find_value("_entry.id")
        keywords = block.find_value("_struct_keywords.pdbx_keywords")
        title = block.find_value("_struct.title")
        if keywords:
            name = keywords.strip("'")
        elif title:
            name = title.strip("'")
        elif entry_id:
            name = entry_id.strip("'")
    if date is None:
        dates = block.find("_pdbx_audit_revision_history.", ["revision_date"])
        if dates:
            # earliest revision = deposition date (a datetime, like the reference's)
            date = min(datetime.datetime.strptime(row[0], "%Y-%m-%d") for row in dates)
    return molecule_type._from_records(arr, id=id, date=date, name=name)


_HETATM_WARNING = ("HETATM line found without residue number. Consider parsing with use_author=True to use "
                   "author-defined residue numbers, or skip_hetatm=True to disable parsing of HETATM altogether.")


def _records_from_rows(table, *, skip_hetatm=False, ignore_endmdl=False) -> np.ndarray:
    """The reference's loop as written (src/pyjess/_io.py:176-219): one `Atom` per row.  Kept as
    the check of `_records_from_columns` (tests/test_cif.py); the loader uses the column form."""
    from ._model import ATOM_DTYPE, Atom
    max_residue_number = 0
    records = []
    for row in table:
        # by default the first model only (_atom_site.pdbx_PDB_model_num)
        if not ignore_endmdl and row[15] != "1":
            break
        if row[14] != "ATOM" and (row[14] != "HETATM" or skip_hetatm):
            continue
        if row[6] == "." and row[14] == "HETATM":
            warnings.warn(_HETATM_WARNING, UserWarning, stacklevel=5)
            residue_number = max_residue_number
            max_residue_number += 1
        else:
            residue_number = int(row[6])
            max_residue_number = max(residue_number, max_residue_number)
        atom = Atom(
            serial=int(row[0]),
            element=row[1],
            name=row[2].strip('"'),
            altloc=" " if row[3] == "." else row[3],
            residue_name=row[4],
            chain_id=row[5],
            residue_number=residue_number,
            insertion_code=" " if row[7] is None or row[7] == "?" else row[7],
            x=float(row[8]),
            y=float(row[9]),
            z=float(row[10]),
            occupancy=0.0 if row[11] == "." else float(row[11]),
            temperature_factor=0.0 if row[12] == "." else float(row[12]),
            charge=0 if row[13] is None or row[13] == "?" else int(row[13]),
        )
        records.append(atom._rec)
    return np.concatenate(records) if records else np.zeros(0, dtype=ATOM_DTYPE)


class _Cached(dict):
    """value -> encoded field, computed once per distinct raw string of a column."""

    def __init__(self, encode):
        super().__init__()
        self.encode = encode

    def __missing__(self, key):
        value = self[key] = self.encode(key)
        return value


def _records_from_columns(cols, *, skip_hetatm=False, ignore_endmdl=False) -> np.ndarray:
    """Same records as `_records_from_rows`, column by column: the row filter is a mask, numbers
    are converted per column, every distinct name / residue / chain string is validated and
    encoded once (through `Atom`, so errors and encodings are the reference-facing ones)."""
    from ._model import ATOM_DTYPE, Atom
    n = len(cols[0])
    model, group = cols[15], cols[14]
    if not ignore_endmdl:
        n = next((i for i, m in enumerate(model) if m != "1"), n)
    if group is None:                      # no group_PDB column: no row is ATOM or HETATM
        return np.zeros(0, dtype=ATOM_DTYPE)
    keep = [i for i in range(n) if group[i] == "ATOM" or (group[i] == "HETATM" and not skip_hetatm)]
    if not keep:
        return np.zeros(0, dtype=ATOM_DTYPE)
    take = (lambda col: [col[i] for i in keep]) if len(keep) != len(group) else (lambda col: col)
    seq, grp = take(cols[6]), take(group)
    if "." in seq:
        # the running maximum of the reference loop (_io.py:185-199)
        numbers, highest = [], 0
        for s, g in zip(seq, grp):
            if s == "." and g == "HETATM":
                warnings.warn(_HETATM_WARNING, UserWarning, stacklevel=5)
                numbers.append(highest)
                highest += 1
            else:
                v = int(s)
                numbers.append(v)
                highest = max(v, highest)
    else:
        numbers = list(map(int, seq))

    def template(**kw):                    # one validated, encoded record per distinct string
        base = dict(serial=0, name="X", residue_name="X", chain_id="A", residue_number=0, x=0.0, y=0.0, z=0.0)
        base.update(kw)
        return Atom(**base)._rec[0]

    arr = np.zeros(len(keep), dtype=ATOM_DTYPE)
    arr["serial"] = list(map(int, take(cols[0])))
    enc = _Cached(lambda s: template(element=s)["element"])
    arr["element"] = [enc[s] for s in take(cols[1])]
    enc = _Cached(lambda s: template(name=s.strip('"'))["name"])
    arr["name"] = [enc[s] for s in take(cols[2])]
    enc = _Cached(lambda s: template(altloc=" " if s == "." else s)["altLoc"])
    arr["altLoc"] = [enc[s] for s in take(cols[3])]
    enc = _Cached(lambda s: template(residue_name=s)["resName"])
    arr["resName"] = [enc[s] for s in take(cols[4])]
    enc = _Cached(lambda s: (lambda r: (r["chainID1"], r["chainID2"]))(template(chain_id=s)))
    chains = [enc[s] for s in take(cols[5])]
    arr["chainID1"], arr["chainID2"] = [c[0] for c in chains], [c[1] for c in chains]
    arr["resSeq"] = numbers
    if cols[7] is None:
        arr["iCode"] = 0x20
    else:
        enc = _Cached(lambda s: template(insertion_code=" " if s == "?" else s)["iCode"])
        arr["iCode"] = [enc[s] for s in take(cols[7])]
    for k, c in enumerate((8, 9, 10)):
        arr["x"][:, k] = list(map(float, take(cols[c])))
    arr["occupancy"] = [0.0 if s == "." else float(s) for s in take(cols[11])]
    arr["tempFactor"] = [0.0 if s == "." else float(s) for s in take(cols[12])]
    if cols[13] is not None:
        enc = _Cached(lambda s: template(charge=0 if s == "?" else int(s))["charge"])
        arr["charge"] = [enc[s] for s in take(cols[13])]
    arr["segID"] = template()["segID"]
    return arr


def looks_like_cif(data: bytes) -> bool:
    """Format detection of ``Molecule.load`` (src/pyjess/_jess.pyx:407-418): skip the lines
    that start with ``#``, then look for ``data_`` / ``loop_``."""
    pos = 0
    while data.startswith(b"#", pos):
        end = data.find(b"\n", pos)
        if end < 0:
            return False
        pos = end + 1
    return data.startswith((b"data_", b"loop_"), pos)
