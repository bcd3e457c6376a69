"""Best hit per matched site (pyjess_b200.sites) against the reference's own post-filter,
``vendor/jess/filter_output.py``.

Goldens: tests/golden/sites_*.txt, written by the UNMODIFIED script (make_golden_sites.py)
from the reference command line's output (cli_*.txt) and from an adversarial synthetic stream
(score ties, self matches, recurring pairs, HETATM, residues differing only in chain).
"""
import io
import os

import numpy as np
import pytest

from pyjess_b200.sites import filter_text, select_sites

from _util import GOLD

CASES = ["cli_default", "cli_flags", "cli_maxcand", "synthetic"]


def _input(name):
    path = os.path.join(GOLD, "sites_synthetic_in.txt" if name == "synthetic" else name + ".txt")
    with open(path) as f:
        return f.read()


def _golden(name, tag):
    with open(os.path.join(GOLD, f"sites_{name}_{tag}.txt")) as f:
        return f.read()


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tag,flag", [("all", True), ("best", False)])
def test_filter_text_matches_script(name, tag, flag):
    out = io.StringIO()
    n = filter_text(io.StringIO(_input(name)), out, all_instances=flag)
    assert out.getvalue() == _golden(name, tag)
    assert n == _golden(name, tag).count("REMARK")


def _script_restatement(block, sites, score, self_match, all_instances):
    """filter_output.py:80-126 hit by hit (dict of lists, strict `<` from 999)."""
    out = []

    def best_of(items):
        low, best = 999, None
        for i in items:
            if score[i] < low:
                low, best = score[i], i
        return best

    def flush(groups):
        unique = []
        for members in groups.values():
            b = best_of(members)
            if b is None or self_match[b]:
                continue
            unique.append(b)
        if all_instances:
            out.extend(unique)
        else:
            b = best_of(unique)
            if b is not None:
                out.append(b)

    groups, prev = {}, None
    for i in range(len(block)):
        if prev is not None and block[i] != block[prev]:
            flush(groups)
            groups = {}
        residues = []
        for r in sites[i]:
            if r >= 0 and r not in residues:
                residues.append(int(r))
        groups.setdefault(tuple(residues), []).append(i)
        prev = i
    if prev is not None:
        flush(groups)
    return out


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("flag", [True, False])
def test_select_sites_random(seed, flag):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 600))
    block = np.sort(rng.integers(0, 6, size=n))
    width = int(rng.integers(1, 7))
    sites = rng.integers(-1, 5, size=(n, width))
    score = rng.choice([-2.0, -1.0, -1.0, 0.25, 3.0, 999.0, 1200.0, np.nan], size=n)
    self_match = np.isin(block, rng.integers(0, 6, size=2))
    got = select_sites(block, sites, score, self_match, flag)
    assert got.tolist() == _script_restatement(block, sites, score, self_match, flag)


def test_select_sites_empty_and_shapes():
    assert select_sites([], np.zeros((0, 3)), [], [], True).tolist() == []
    with pytest.raises(ValueError):
        select_sites([0, 0], np.zeros((1, 3)), [0.0, 0.0], [False, False])
    # [a, a, b] and [a, b, b] are the same residue list; [b, a] is another one
    got = select_sites([0, 0, 0], [[1, 1, 2], [1, 2, 2], [2, 1, -1]], [0.5, 0.25, 0.1], [False] * 3)
    assert got.tolist() == [1, 2]


def _fake_queries(n_hits, seed, wide=False):
    """`Query` objects over a hand-made hit table (the host layer only: no engine call).
    `wide`: serial and residue numbers that overflow the fixed PDB columns in some structures."""
    import types
    from pyjess_b200 import Template, synth
    from pyjess_b200._jess import Jess, Query, _BatchRun
    tpl, mols, thr, _ = synth.workload("cfg3", scale=0.012)
    tpl = [Template(list(t), id=mols[1].id if i == 2 else t.id) for i, t in enumerate(tpl)]      # a self match
    if wide:
        mols = [type(m)._from_records(m._atoms.copy(), id=m.id) for m in mols]
        for m in mols[::2]:
            m._atoms["serial"] += 99990                      # six-digit serials shift every later column
            m._atoms["resSeq"][m._atoms["resSeq"] > 150] += 9900     # five-digit residue numbers too
    jess = Jess.__new__(Jess)
    jess._templates = tpl
    rng = np.random.default_rng(seed)
    mol = np.sort(rng.integers(0, len(mols), n_hits))
    t = rng.integers(0, len(tpl), n_hits)
    order = np.lexsort((-t, mol))                      # iteration order: templates descending per structure
    mol, t = mol[order], t[order]
    atoms = np.full((n_hits, 18), -1, dtype=np.int32)
    for i in range(n_hits):
        n = len(tpl[t[i]])
        atoms[i, :n] = rng.integers(0, 60) + rng.integers(0, 24, n)        # few residues: many repeated sites
    rot = np.tile(np.eye(3), (n_hits, 1, 1))
    table = types.SimpleNamespace(count=n_hits, molecule=mol.astype(np.int32), template=t.astype(np.int32),
                                  atoms=atoms, rmsd=rng.integers(0, 40, n_hits) / 20.0, rotation=rot,
                                  centre0=np.zeros((n_hits, 3)), centre1=np.zeros((n_hits, 3)))
    run = _BatchRun(jess, mols, (2.0, 1.5, 1.5, None, True, False), False)
    run.done, run.table = True, table
    bounds = np.searchsorted(mol, np.arange(len(mols) + 1))
    run._slices = [np.arange(bounds[m], bounds[m + 1]) for m in range(len(mols))]
    return [Query(jess, m, 2.0, False, None, None, run, i) for i, m in enumerate(mols)]


@pytest.mark.parametrize("wide", [False, True])
@pytest.mark.parametrize("flag", [True, False])
def test_best_sites_over_a_hit_table_equals_the_text_route(flag, wide):
    """`best_sites` (arrays of the hit table) == `filter_text` (pinned to the script above) on
    the dump of every hit, `Hit.dump` text included."""
    from pyjess_b200 import sites
    text = io.StringIO()
    for query in _fake_queries(3000, 5, wide):
        for hit in query:
            hit.dump(text)
            text.write("\n")
    text.seek(0)
    want = io.StringIO()
    n = sites.filter_text(text, want, all_instances=flag)
    hits = sites.best_sites(_fake_queries(3000, 5, wide), all_instances=flag)
    got = io.StringIO()
    sites.write_hits(hits, got)
    assert 0 < n == len(hits) < 3000
    assert got.getvalue() == want.getvalue()


def test_best_sites_rejects_what_it_cannot_filter():
    from pyjess_b200 import sites
    queries = _fake_queries(50, 1)
    next(q for q in queries if len(q.hit_table()[1]))._pos = 1
    with pytest.raises(RuntimeError):
        sites.best_sites(queries)
    with pytest.raises(TypeError):
        sites.best_sites([object()])
    assert sites.best_sites([]) == []


@pytest.mark.parametrize("tag,flags", [("all", []), ("best", ["--best"])])
def test_module_entry_point_is_a_drop_in_for_the_script(tag, flags):
    import subprocess
    import sys
    root = os.path.dirname(GOLD.rstrip(os.sep).rsplit(os.sep, 1)[0])
    proc = subprocess.run([sys.executable, "-m", "pyjess_b200.sites", *flags], input=_input("synthetic"), text=True,
                          capture_output=True, cwd=root, timeout=120)
    assert proc.returncode == 0, proc.stderr
    assert proc.stdout == _golden("synthetic", tag)
