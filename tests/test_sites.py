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
