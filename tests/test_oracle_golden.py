"""Pins the CPU oracle (oracle/jess_oracle.c) against the reference's golden
vectors (tests/golden/reference_hits.json, made by the reference's own Python
API) and, when present, against the unmodified reference C library
(oracle/_ref/libjess_ref.so).  CPU only."""
import numpy as np
import pytest

from _util import assert_matches_golden, case_inputs, compare_tables, golden_cases
from pyjess_b200._tables import MoleculeBatch, TemplateTable

CASES = sorted(c for c in golden_cases() if not golden_cases()[c].get("single_atom"))
# (best_match over single-atom templates is decided by the host layer AFTER the first-candidate drop: tests/test_api_*.py)
SINGLE = sorted(c for c in golden_cases() if golden_cases()[c].get("single_atom")
                and not golden_cases()[c]["kwargs"].get("best_match"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(oracle_lib, name):
    case = golden_cases()[name]
    templates, mols, params = case_inputs(case)
    table, stats = oracle_lib.query(TemplateTable(templates), MoleculeBatch(mols), params, threads=4)
    assert table.count == case["n_hits"]
    assert_matches_golden(table, case, best_match=bool(params.best_match))
    assert stats["hits"] == table.count


@pytest.mark.parametrize("name", CASES)
def test_reference_library_matches_golden(oracle_lib, name):
    if not oracle_lib.ref_available():
        pytest.skip("oracle/_ref/libjess_ref.so not built (no reference checkout)")
    case = golden_cases()[name]
    templates, mols, params = case_inputs(case)
    tt, mb = TemplateTable(templates), MoleculeBatch(mols)
    ref, _ = oracle_lib.ref_query(tt, mb, params, threads=4)
    assert_matches_golden(ref, case, best_match=bool(params.best_match))
    ours, _ = oracle_lib.query(tt, mb, params, threads=4)
    if not params.best_match:
        assert compare_tables(ours, ref, params.rmsd_threshold) == []
        # same arithmetic, same libm: the restatement is bit-identical here
        assert np.array_equal(ours.rmsd, ref.rmsd)


@pytest.mark.parametrize("name", SINGLE)
def test_single_atom_templates_report_every_candidate_but_one(oracle_lib, name):
    """n == 1: oracle (and engine ABI) list ALL candidates; the reference lists all but the
    first of its own candidate-list order -- per (molecule, single-atom template) exactly one
    hit fewer, and a subset."""
    case = golden_cases()[name]
    templates, mols, params = case_inputs(case)
    table, _ = oracle_lib.query(TemplateTable(templates), MoleculeBatch(mols), params, threads=2)
    gold = {(h.get("molecule", 0), h["template"], tuple(h["atoms"])) for h in case["hits"]}
    ours = set(table.keys())
    assert gold <= ours
    extra = ours - gold
    singles = [i for i, t in enumerate(templates) if len(t) == 1]
    assert all(k[1] in singles for k in extra)
    pairs = {(k[0], k[1]) for k in ours if k[1] in singles}
    assert len(extra) == len(pairs)
