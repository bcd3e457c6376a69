"""Shared helpers of the parity tests: golden loading, hit comparison."""
import json
import os

import numpy as np

from pyjess_b200 import Molecule, Template
from pyjess_b200._model import ATOM_DTYPE
from pyjess_b200._tables import HitTable, MoleculeBatch, TemplateTable, query_params

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "data")
GOLD = os.path.join(HERE, "golden")

#: north-star tolerance: RMSD and transformation within 1e-6 A absolute
TOL = 1e-6

_cache = {}


def golden_cases():
    if "cases" not in _cache:
        with open(os.path.join(GOLD, "reference_hits.json")) as f:
            _cache["cases"] = {c["name"]: c for c in json.load(f)["cases"]}
    return _cache["cases"]


def load_template(name):
    key = ("t", name)
    if key not in _cache:
        _cache[key] = Template.load(os.path.join(DATA, name))
    return _cache[key]


def load_molecule(name):
    key = ("m", name)
    if key not in _cache:
        _cache[key] = Molecule.load(os.path.join(DATA, name))
    return _cache[key]


def cfg2_molecules():
    """The 132 vendor example structures from the packed golden npz."""
    if "cfg2" not in _cache:
        z = np.load(os.path.join(GOLD, "cfg2_structures.npz"))
        off = z["atom_offset"]
        mols = []
        for m in range(len(off) - 1):
            a, b = int(off[m]), int(off[m + 1])
            arr = np.zeros(b - a, dtype=ATOM_DTYPE)
            arr["x"][:, 0] = z["xi"][a:b] / 1000.0
            arr["x"][:, 1] = z["yi"][a:b] / 1000.0
            arr["x"][:, 2] = z["zi"][a:b] / 1000.0
            arr["serial"] = z["serial"][a:b]
            arr["name"] = z["name"][a:b]
            arr["resName"] = z["resname"][a:b]
            arr["resSeq"] = z["resseq"][a:b]
            arr["chainID1"] = z["chain1"][a:b]
            arr["chainID2"] = z["chain2"][a:b]
            arr["altLoc"] = 0x20
            arr["iCode"] = 0x20
            mols.append(Molecule._from_records(arr, id=str(z["ids"][m])[3:7]))
        _cache["cfg2"] = mols
    return _cache["cfg2"]


def case_inputs(case):
    templates = [load_template(t) for t in case["templates"]]
    if case["molecule"].endswith(".npz"):
        mols = cfg2_molecules()
    else:
        mols = [load_molecule(case["molecule"])]
    kw = dict(case["kwargs"])
    params = query_params(*case["params"], ignore_chain=kw.get("ignore_chain"),
                          reorder=kw.get("reorder", True), best_match=kw.get("best_match", False))
    return templates, mols, params


def hit_matrix(rot, c0, c1):
    """T(c1) R T(-c0) as a 4x4 (Hit.transformation, _jess.pyx:2343-2377)."""
    m = np.eye(4)
    m[:3, :3] = rot
    m[:3, 3] = c1 - rot @ c0
    return m


def golden_map(case):
    out = {}
    for h in case["hits"]:
        key = (h.get("molecule", 0), h["template"], tuple(h["atoms"]))
        out.setdefault(key, []).append(h)
    return out


def assert_matches_golden(table: HitTable, case, best_match=False):
    gold = golden_map(case)
    keys = table.keys()
    if best_match:
        # one hit per (molecule, template); identity of the winner is pinned by RMSD
        assert sorted((k[0], k[1]) for k in keys) == sorted((k[0], k[1]) for k in gold for _ in gold[k])
    else:
        assert sorted(keys) == sorted(k for k in gold for _ in gold[k]), "hit key sets differ"
    for i, key in enumerate(keys):
        if key not in gold:
            assert best_match
            continue
        g = gold[key][0]
        assert abs(table.rmsd[i] - g["rmsd"]) <= TOL
        m = hit_matrix(table.rotation[i], table.centre0[i], table.centre1[i])
        assert np.abs(m - np.asarray(g["transformation"])).max() <= TOL


def conditioning(table: HitTable, i: int, templates, mols) -> float:
    """(sigma_max / sigma_min)^2 of the 3x3 covariance of hit i: the reference solves the
    eigenproblem of X^T X (vendor/jess/src/Super.c:417-538), whose smallest eigenvalue it
    resolves only to ~eps * lambda_max; the rotation inherits a relative error of
    ~eps * cond.  Near-coplanar atom sets therefore have NO reproducible rotation in the
    reference itself (its own det(R) is off by 1e-6 and more there)."""
    tpl = templates[int(table.template[i])]
    n = len(tpl)
    x = mols[int(table.molecule[i])]._atoms["x"][table.atoms[i, :n]]
    y = np.array([a._pos for a in tpl], dtype=np.float64)
    sv = np.linalg.svd((x - x.mean(0)).T @ (y - y.mean(0)), compute_uv=False)
    return float((sv[0] / max(sv[-1], 1e-300)) ** 2)


#: rotation tolerance for an ill-conditioned superposition: TOL + ILL * cond (ILL ~ 50 eps)
ILL = 1e-14


def compare_tables(a: HitTable, b: HitTable, rmsd_threshold=None, templates=None, mols=None, report=None):
    """Bit-exact key sets, values within TOL; returns the list of keys that lie
    within TOL of the RMSD threshold in either table (reported separately).
    With `templates`/`mols`, a rotation outside TOL is tolerated only when the hit is
    provably ill-conditioned (|dR| <= TOL + ILL * cond); those keys are appended to
    `report` (listed separately)."""
    ka, kb = a.keys(), b.keys()
    edge = []
    if ka != kb:
        sa, sb = set(ka), set(kb)
        diff = sa ^ sb
        ra = {k: a.rmsd[i] for i, k in enumerate(ka)}
        rb = {k: b.rmsd[i] for i, k in enumerate(kb)}
        for k in diff:
            r = ra.get(k, rb.get(k))
            assert rmsd_threshold is not None and abs(r - rmsd_threshold) <= TOL, f"hit {k} only on one side (rmsd {r})"
            edge.append(k)
    ib = {k: i for i, k in enumerate(kb)}
    for i, k in enumerate(ka):
        j = ib.get(k)
        if j is None:
            continue
        assert abs(a.rmsd[i] - b.rmsd[j]) <= TOL, (k, a.rmsd[i], b.rmsd[j])
        d = np.abs(a.rotation[i] - b.rotation[j]).max()
        if d > TOL:
            assert templates is not None, k
            cond = conditioning(a, i, templates, mols)
            assert d <= TOL + ILL * cond, (k, d, cond)
            if report is not None:
                report.append((k, float(d), cond))
        assert np.abs(a.centre0[i] - b.centre0[j]).max() <= TOL, k
        assert np.abs(a.centre1[i] - b.centre1[j]).max() <= TOL, k
    return edge
