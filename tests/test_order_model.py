"""The pure-Python model of the reference's iteration order (tests/_order_model.py)
reproduces the ORDER of the reference's own hits (tests/golden/reference_hits.json stores
them as yielded): pins the theory csrc/order.cuh implements.  CPU only."""
import pytest

from _order_model import reference_sort_key
from _util import case_inputs, golden_cases

ORDER_CASES = ["amy_133_reorder", "amy_133_noreorder", "5ayx_412_ignore_atoms", "amy_doctest", "lnb_t01_2_5_5",
               "single_glu_asp_pdb1lnb", "single_any_ca_pdb1lnb", "single_mixed_lnb", "cfg2_examples_all"]


@pytest.mark.parametrize("name", ORDER_CASES)
def test_model_reproduces_reference_order(name):
    case = golden_cases()[name]
    templates, mols, _ = case_inputs(case)
    reorder = case["kwargs"].get("reorder", True)
    runs = {}
    for h in case["hits"]:
        runs.setdefault((h.get("molecule", 0), h["template"]), []).append(tuple(h["atoms"]))
    # templates come in reverse insertion order, structures in input order
    seen = [k for k in dict.fromkeys((h.get("molecule", 0), h["template"]) for h in case["hits"])]
    assert seen == sorted(seen, key=lambda k: (k[0], -k[1]))
    for (m, t), tuples in runs.items():
        if len(tuples) < 2:
            continue
        key = reference_sort_key(templates[t], mols[m], reorder)
        assert sorted(tuples, key=key) == tuples, (name, m, t)


# ---- max_candidates: the host replay of Query.__next__ over a CPU-built tuple stream ---------
import json
import os

import numpy as np

import oracle
from _util import DATA, GOLD, load_molecule
from pyjess_b200 import Jess, Template
from pyjess_b200._jess import Query, _BatchRun
from pyjess_b200._tables import HitTable, MoleculeBatch, TemplateTable, query_params


def max_candidates_cases():
    with open(os.path.join(GOLD, "reference_max_candidates.json")) as f:
        return json.load(f)["cases"]


_streams = {}


def tuple_stream(case):
    """All tuples of the case (oracle, no RMSD filter) in the reference's iteration order."""
    kw = case["kwargs"]
    key = (tuple(case["templates"]), case["molecule"], tuple(case["params"]), kw["reorder"], kw.get("ignore_chain"))
    if key not in _streams:
        templates = [Template.load(os.path.join(DATA, t)) for t in case["templates"]]   # distinct objects
        mol = load_molecule(case["molecule"])
        qp = query_params(1e300, case["params"][1], case["params"][2], ignore_chain=kw.get("ignore_chain"))
        table, _ = oracle.query(TemplateTable(templates), MoleculeBatch([mol]), qp)
        rows = []
        for t in sorted(set(table.template.tolist()), reverse=True):
            sel = np.nonzero(table.template == t)[0]
            rk = reference_sort_key(templates[t], mol, kw["reorder"])
            rows.extend(sorted(sel.tolist(), key=lambda i: rk(table.atoms[i])))
        rows = np.asarray(rows, dtype=np.int64)
        _streams[key] = (templates, mol, HitTable.from_arrays(*(a[rows] for a in table.arrays())))
    return _streams[key]


@pytest.mark.parametrize("case", max_candidates_cases(), ids=lambda c: c["name"])
def test_replay_matches_reference_max_candidates(case):
    templates, mol, table = tuple_stream(case)
    kw = case["kwargs"]
    jess = Jess(templates)
    run = _BatchRun(jess, [mol], (case["params"][0], case["params"][1], case["params"][2], kw.get("ignore_chain"),
                                  kw["reorder"], kw["best_match"]), all_tuples=True)
    run.done, run.table, run._slices = True, table, [np.arange(table.count)]
    q = Query(jess, mol, case["params"][0], kw["best_match"], kw.get("ignore_chain"), kw["max_candidates"], run, 0)
    def tid(t):
        return next(i for i, x in enumerate(jess._templates) if x is t)
    serial = mol._atoms["serial"]
    got = [(tid(h.template()), [a.serial for a in h.atoms(transform=False)], h.rmsd) for h in q]
    want = [(h["template"], [int(serial[i]) for i in h["atoms"]], h["rmsd"]) for h in case["hits"]]
    assert [g[:2] for g in got] == [w[:2] for w in want]
    assert np.allclose([g[2] for g in got], [w[2] for w in want], rtol=0, atol=1e-6)


# ---- the model against the UNMODIFIED reference library on synthetic structures -------------
@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/libjess_ref.so not built")
@pytest.mark.parametrize("reorder", [True, False])
def test_model_matches_reference_library_on_synthetic(reorder):
    from pyjess_b200 import synth
    tpl, mols, params, kw = synth.workload("cfg3dense", scale=0.06)
    tpl, mols = tpl[:6], mols[:2]
    rng = np.random.default_rng(7)
    for m in mols:   # scrambled, partly duplicated serials: the residue index sorts by them
        s = rng.permutation(len(m)).astype(np.int32) // 2
        m._atoms["serial"] = s
    qp = query_params(*params, reorder=1 if reorder else 0, hit_order="reference" if reorder else "reference-noreorder")
    table, _ = oracle.ref_query(TemplateTable(tpl), MoleculeBatch(mols, serial=True), qp)
    assert table.count > 200
    runs = {}
    for i in range(table.count):
        runs.setdefault((int(table.molecule[i]), int(table.template[i])), []).append(tuple(table.atoms[i].tolist()))
    assert list(runs) == sorted(runs, key=lambda k: (k[0], -k[1]))
    for (m, t), tuples in runs.items():
        key = reference_sort_key(tpl[t], mols[m], reorder)
        assert sorted(tuples, key=key) == tuples, (m, t)
