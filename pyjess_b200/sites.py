"""Best hit per matched site: the post-filter of ``vendor/jess/filter_output.py``.

The reference ships a script that reads the text the Jess / pyjess command line writes and
keeps, for every run of consecutive hits of one (target, template) pair, ONE hit per
distinct list of matched residues -- the one with the lowest ``log(E)`` as printed (two
decimals), the first of equal scores -- and drops self matches (``template == target``)
(filter_output.py:80-126).  Here the same selection runs

* over the engine's hit table without materialising `Hit` objects (`best_sites` on the
  queries returned by `Jess.query` / `Jess.query_many`): residue identifiers are one gather
  through a per-structure table, the grouping is one row-wise `numpy.unique`, the choice of
  the best record one `lexsort`;
* over text (`filter_text`), byte-identical to the script on the same input.

Both go through `select_sites`, the array form of filter_output.py:80-126.
"""
from __future__ import annotations

import math
import typing

import numpy as np

__all__ = ["select_sites", "best_sites", "filter_text", "write_hits"]

#: `min_score = 999` of get_highest_scored_match (filter_output.py:81): a hit scoring 999 or
#: more is never selected (the script then fails on ``None.template``; here its site is skipped)
SCORE_LIMIT = 999.0


def select_sites(block: np.ndarray, sites: np.ndarray, score: np.ndarray, self_match: np.ndarray,
                 all_instances: bool = True) -> np.ndarray:
    """Rows the script prints, in the order it prints them.

    `block[i]`  run id of hit i (consecutive hits of one (target, template) pair share it,
                non-decreasing): the unit `main` flushes (filter_output.py:118-121);
    `sites[i]`  integer residue identifier of every matched atom (-1 padding): the script's
                ``resname_chain_resid`` strings (add_residue, :62-68);
    `score[i]`  the score the script compares (``fields[-1]`` of the REMARK line, :56);
    `self_match[i]`  template name == target name (:92-94).
    """
    block = np.asarray(block, dtype=np.int64)
    sites = np.asarray(sites, dtype=np.int64)
    score = np.asarray(score, dtype=np.float64)
    self_match = np.asarray(self_match, dtype=bool)
    n = len(block)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    if sites.ndim != 2 or len(sites) != n or len(score) != n or len(self_match) != n:
        raise ValueError("select_sites: arrays of different lengths")
    width = sites.shape[1]
    # Match.residues: the DISTINCT residues in atom order (`if res not in self.residues`)
    dup = sites < 0
    for k in range(1, width):
        dup[:, k] |= (sites[:, :k] == sites[:, k:k + 1]).any(axis=1)
    order = np.argsort(dup, axis=1, kind="stable")
    canon = np.where(np.take_along_axis(dup, order, axis=1), -1, np.take_along_axis(sites, order, axis=1))
    # matches[':'.join(residues)] inside one flush unit
    key = np.concatenate([block[:, None], canon], axis=1)
    _, group = np.unique(key, axis=0, return_inverse=True)
    group = group.reshape(-1)
    n_groups = int(group.max()) + 1
    index = np.arange(n, dtype=np.int64)
    # get_highest_scored_match: strict `<` from 999 -> lowest score, first of equals; NaN never wins
    eligible = score < SCORE_LIMIT
    rank = np.lexsort((index, np.where(eligible, score, np.inf), group))
    head = np.ones(n, dtype=bool)
    head[1:] = group[rank][1:] != group[rank][:-1]
    best = rank[head]                                    # one row per group, group ascending
    first = np.full(n_groups, n, dtype=np.int64)
    np.minimum.at(first, group, index)                   # dict insertion order = first occurrence
    keep = eligible[best] & ~self_match[best]
    best, first = best[keep], first[keep]
    best = best[np.argsort(first, kind="stable")]
    if all_instances or len(best) == 0:
        return best
    # print_best_match(all_instances=False): the best of the unique matches of each flush unit
    b = block[best]
    pos = np.arange(len(best))
    rank = np.lexsort((pos, score[best], b))
    head = np.ones(len(best), dtype=bool)
    head[1:] = b[rank][1:] != b[rank][:-1]
    return best[np.sort(rank[head])]


def _printed(values: np.ndarray, fmt: str) -> np.ndarray:
    """The numbers as the script reads them back from the text (``float(fields[..])``)."""
    if len(values) == 0:
        return np.zeros(0, dtype=np.float64)
    return np.char.mod(fmt, values).astype(np.float64)


def _left_justified(cols: np.ndarray) -> np.ndarray:
    """`str.strip` on fixed-width byte columns, as far as equality goes: '_'/NUL are blanks in the
    dumped text, leading blanks are shifted out, trailing ones are padding either way."""
    cols = np.where((cols == 0) | (cols == 0x5F), 0x20, cols).astype(np.int64)
    for _ in range(cols.shape[1] - 1):
        lead = cols[:, 0] == 0x20
        cols = np.where(lead[:, None], np.concatenate([cols[:, 1:], np.full((len(cols), 1), 0x20)], axis=1), cols)
    return cols


def _residue_ids(records: np.ndarray) -> np.ndarray:
    """One integer per atom, equal exactly when the script's ``resname_chain_resid`` strings
    are: columns 18-20, 21-22 and 23-26 of the dumped line, each stripped (:63-66).  Injective
    over the residue numbers that fit the fixed columns (`_fixed_columns`): 5 bytes + 14 bits."""
    name = _left_justified(np.frombuffer(np.ascontiguousarray(records["resName"]).tobytes(), dtype=np.uint8).reshape(-1, 4)[:, :3])
    chain = _left_justified(np.stack([records["chainID1"], records["chainID2"]], axis=1))
    key = (name[:, 0] << 48) | (name[:, 1] << 40) | (name[:, 2] << 32) | (chain[:, 0] << 24) | (chain[:, 1] << 16)
    return key | (records["resSeq"].astype(np.int64) + 1024)


def _residue_ids_from_text(records: np.ndarray, atoms: np.ndarray, ids: dict) -> np.ndarray:
    """Residue identifiers of the matched atoms `atoms` ([hit][atom], -1 padding) read from the
    dumped lines themselves: a serial above 99,999 or a residue number above 9,999 widens its
    field and shifts the columns the script slices, so only the text tells what it sees."""
    from . import _capi
    used = np.unique(atoms[atoms >= 0])
    lines = _capi.format_atoms_pdb(np.ascontiguousarray(records[used])).splitlines()
    key = np.empty(len(used), dtype=np.int64)
    for j, line in enumerate(lines):
        line = line.strip()
        res = (line[17:20].strip(), line[20:22].strip(), line[22:26].strip())
        key[j] = (1 << 62) + ids.setdefault(res, len(ids))   # above every packed key (< 2**56)
    lookup = np.full(len(records), -1, dtype=np.int64)
    lookup[used] = key
    return np.where(atoms >= 0, lookup[np.maximum(atoms, 0)], -1)


def _fixed_columns(records: np.ndarray) -> bool:
    """Whether every dumped line keeps the fixed columns the script slices: `%5i` serial and
    `%4i` residue number do not overflow their fields (_jess.pyx:1273-1291)."""
    serial, resseq = records["serial"], records["resSeq"]
    return bool(((serial >= -9999) & (serial <= 99999) & (resseq >= -999) & (resseq <= 9999)).all())


def best_sites(queries, all_instances: bool = True) -> list:
    """`Hit`s the script keeps out of the hits of `queries`, in the order it prints them.

    `queries`: one `Query` or a sequence of them (``Jess.query_many``), not yet iterated and
    without ``max_candidates`` (that stream is filtered while it is replayed: dump it and use
    `filter_text`).  Equivalent to dumping every hit in iteration order, piping the text through
    ``filter_output.py`` (``main(all_instances)``) and reading the hits back.
    """
    from ._jess import Query
    if isinstance(queries, Query):
        queries = [queries]
    queries = list(queries)
    sites, log_es, selfs, owners, pair_of_row = [], [], [], [], []
    mol_names: dict = {}
    tpl_names: dict = {}
    text_ids: dict = {}        # residue strings of structures whose numbers overflow the fixed columns
    per_jess: dict = {}        # id(jess) -> (dim[t], name id[t]) of its templates
    for q, query in enumerate(queries):
        if not isinstance(query, Query):
            raise TypeError(f"best_sites expects Query objects, found {type(query).__name__}")
        if query._run.all_tuples or query._pos:
            raise RuntimeError("best_sites needs fresh queries without max_candidates")
        table, rows = query.hit_table()
        if len(rows) == 0:
            continue
        mol = query.molecule
        if mol.id is None:
            raise RuntimeError("cannot filter hits where `molecule.id` is `None`")
        if id(query.jess) not in per_jess:
            templates = query.jess._templates
            if any(t.id is None for t in templates):
                raise RuntimeError("cannot filter hits where `template.id` is `None`")
            per_jess[id(query.jess)] = (np.asarray([float(t._dim) for t in templates], dtype=np.float64),
                                        np.asarray([tpl_names.setdefault(t.id, len(tpl_names)) for t in templates],
                                                   dtype=np.int64))
        dim, tpl_name = per_jess[id(query.jess)]
        tpl = table.template[rows].astype(np.int64)
        atoms = table.atoms[rows].astype(np.int64)
        if _fixed_columns(mol._atoms):
            res = _residue_ids(mol._atoms)
            sites.append(np.where(atoms >= 0, res[np.maximum(atoms, 0)], -1))
        else:
            sites.append(_residue_ids_from_text(mol._atoms, atoms, text_ids))
        # Hit.log_evalue (TessTemplate_logE) evaluated in the same order of operations
        log_es.append(-8.5 + 2.5 * table.rmsd[rows] + -0.7 * dim[tpl] + math.log(float(len(mol))))
        # Match.from_remark splits the REMARK line on blanks: names are compared as printed
        selfs.append(tpl_name[tpl] == tpl_names.get(mol.id, -1))
        pair_of_row.append(mol_names.setdefault(mol.id, len(mol_names)) * (1 << 32) + tpl_name[tpl])
        owners.append((q, rows))
    if not sites:
        return []
    width = max(s.shape[1] for s in sites)
    sites = [np.pad(s, ((0, 0), (0, width - s.shape[1])), constant_values=-1) for s in sites]
    pair_of_row = np.concatenate(pair_of_row)
    # flush units: runs of CONSECUTIVE hits with the same (target, template) names
    block = np.zeros(len(pair_of_row), dtype=np.int64)
    block[1:] = np.cumsum(pair_of_row[1:] != pair_of_row[:-1])
    chosen = select_sites(block, np.concatenate(sites), _printed(np.concatenate(log_es), "%4.2f"), np.concatenate(selfs), all_instances)
    starts = np.cumsum([0] + [len(rows) for _, rows in owners])
    owner = np.searchsorted(starts, chosen, side="right") - 1
    all_rows = np.concatenate([rows for _, rows in owners])[chosen]
    return [queries[owners[o][0]]._make_hit(r) for o, r in zip(owner.tolist(), all_rows.tolist())]


def filter_text(lines: typing.Iterable[str], out: typing.TextIO, all_instances: bool = True) -> int:
    """``filter_output.py`` over the text of the command line: same bytes on `out`.
    Returns the number of hits written."""
    remarks, atoms_of, residues_of, targets, templates, scores = [], [], [], [], [], []
    current = None
    for line in lines:
        line = line.strip()
        if line.startswith("REMARK"):
            fields = line.split()
            # Match.from_remark (:54-57); float() as in Match.__init__
            current = (fields[1], fields[3], float(fields[2]), float(fields[-1]), line, [], [])
        if line.startswith("ATOM") or line.startswith("HETATM"):
            current[5].append(line)
            res = (line[17:20].strip(), line[20:22].strip(), line[22:26].strip())
            if res not in current[6]:
                current[6].append(res)
        if line.startswith("ENDMDL"):
            targets.append(current[0]); templates.append(current[1]); scores.append(current[3])
            remarks.append(current[4]); atoms_of.append(current[5]); residues_of.append(current[6])
    n = len(remarks)
    if n == 0:
        return 0
    ids: dict = {}
    width = max(1, max(len(r) for r in residues_of))
    sites = np.full((n, width), -1, dtype=np.int64)
    for i, residues in enumerate(residues_of):
        for k, res in enumerate(residues):
            # the script joins the strings with ':' and '_': keep its (loss-free here) identity
            sites[i, k] = ids.setdefault("_".join(res), len(ids))
    pair = np.zeros(n, dtype=np.int64)
    names: dict = {}
    for i in range(n):
        pair[i] = names.setdefault((targets[i], templates[i]), len(names))
    block = np.zeros(n, dtype=np.int64)
    block[1:] = np.cumsum(pair[1:] != pair[:-1])
    same = np.asarray([targets[i] == templates[i] for i in range(n)], dtype=bool)
    chosen = select_sites(block, sites, np.asarray(scores, dtype=np.float64), same, all_instances)
    for i in chosen:
        out.write("{}\n{}\nENDMDL\n\n".format(remarks[i], "\n".join(atoms_of[i])))
    return len(chosen)


def write_hits(hits, file: typing.TextIO, transform: bool = True) -> None:
    """The selected hits in the script's output format (``Match.__repr__`` + print, :47-48):
    the `Hit.dump` block with every line stripped, followed by an empty line."""
    for hit in hits:
        lines = [line.strip() for line in hit.dumps(format="pdb", transform=transform).splitlines()]
        file.write("\n".join(lines))
        file.write("\n\n")


def main(argv: typing.Optional[typing.List[str]] = None) -> int:
    """``python -m pyjess_b200.sites < hits.txt``: drop-in for ``filter_output.py`` on the command
    line's text (``--best``: the script's other mode, one hit per (target, template) run)."""
    import argparse
    import sys
    parser = argparse.ArgumentParser(prog="python -m pyjess_b200.sites", description=main.__doc__)
    parser.add_argument("--best", action="store_true", help="only the best hit of every (target, template) run")
    args = parser.parse_args(argv)
    filter_text(sys.stdin, sys.stdout, all_instances=not args.best)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
