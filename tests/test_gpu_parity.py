"""GPU parity tests proper: the CUDA engine, called through the C ABI, against
(a) the reference's golden vectors and (b) the CPU oracle on the same inputs.
Hit key sets must be bit-exact; RMSD / rotation / centroids within TOL = 1e-6."""
import numpy as np
import pytest

from _util import TOL, assert_matches_golden, case_inputs, compare_tables, golden_cases
from pyjess_b200._tables import MoleculeBatch, TemplateTable

pytestmark = pytest.mark.gpu

CASES = sorted(golden_cases())
COUNTERS = ("pairs", "live_pairs", "candidate_atoms", "frontier_items", "frontier_ints", "candidate_tests",
            "tuples", "nan_dropped")


def run_engine(engine, templates, mols, params):
    from pyjess_b200 import _capi
    tt = TemplateTable(templates)
    dev = _capi.DeviceTemplates(engine, tt)
    try:
        return _capi.query_host(engine, dev, MoleculeBatch(mols), params)
    finally:
        dev.close()


@pytest.mark.parametrize("name", CASES)
def test_engine_matches_reference_golden(engine, name):
    case = golden_cases()[name]
    templates, mols, params = case_inputs(case)
    table, stats = run_engine(engine, templates, mols, params)
    assert table.count == case["n_hits"]
    assert_matches_golden(table, case, best_match=bool(params.best_match))


@pytest.mark.parametrize("name", CASES)
def test_engine_matches_oracle(engine, oracle_lib, name):
    case = golden_cases()[name]
    templates, mols, params = case_inputs(case)
    table, stats = run_engine(engine, templates, mols, params)
    ours, ostats = oracle_lib.query(TemplateTable(templates), MoleculeBatch(mols), params, threads=8)
    if params.best_match:
        assert table.count == ours.count
        assert np.abs(np.sort(table.rmsd) - np.sort(ours.rmsd)).max() <= TOL
        return
    assert compare_tables(table, ours, params.rmsd_threshold) == []
    # the algorithm-independent work counters must agree exactly (they define the roofline bytes)
    for key in COUNTERS:
        assert stats[key] == ostats[key], key
    assert stats["hits"] == table.count
