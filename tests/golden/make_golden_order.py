#!/usr/bin/env python3
"""Golden fixtures for the ORDER-DEPENDENT options of Query (max_candidates), generated
from the REFERENCE itself (baseline/_ref, see make_golden.py):

    python tests/golden/make_golden_order.py   ->  tests/golden/reference_max_candidates.json

Each case stores the sequence of hits exactly as the reference's Query yields them:
(template index, atom indices in template order, rmsd)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import pyjess  # the reference  # noqa: E402
from make_golden import DATA, GOLD, atom_index  # noqa: E402


def run(templates, mol, params, kw):
    jess = pyjess.Jess(templates)
    index = atom_index(mol)
    tid = {id(t): i for i, t in enumerate(templates)}
    out = []
    for hit in jess.query(mol, *params, **kw):
        atoms = [index[(a.serial, a.name, a.residue_number, a.chain_id, a.x, a.y, a.z)]
                 for a in hit.atoms(transform=False)]
        out.append({"template": tid[id(hit.template())], "atoms": atoms, "rmsd": hit.rmsd})
    return out


def main():
    load_t = lambda name: pyjess.Template.load(os.path.join(DATA, name))  # noqa: E731
    load_m = lambda name: pyjess.Molecule.load(os.path.join(DATA, name))  # noqa: E731
    setups = [
        ("amy_133", ["1.3.3.tpl"], "1AMY.pdb", (2.0, 4.0, 4.0)),
        ("amy_133_tight", ["1.3.3.tpl"], "1AMY.pdb", (1.0, 4.0, 4.0)),
        ("amy_two", ["1.3.3.tpl", "4.1.2.tpl"], "1AMY.pdb", (2.0, 2.0, 2.0)),
        ("amy_133_x2", ["1.3.3.tpl", "1.3.3.tpl"], "1AMY.pdb", (1.5, 4.0, 4.0)),
        ("lnb_t01", ["template_01.qry"], "pdb1lnb.ent", (2.0, 5.0, 5.0)),
        ("lnb_examples", ["template_01.qry", "template_02.qry", "template_03.qry"], "pdb1lnb.ent", (2.0, 5.0, 5.0)),
        ("ayx_412_atoms", ["4.1.2.tpl"], "5ayx.EF.pdb", (3.0, 3.0, 3.0)),
    ]
    cases = []
    for name, tfiles, mfile, params in setups:
        templates = [load_t(f) for f in tfiles]
        mol = load_m(mfile)
        for mc in (0, 1, 2, 3, 7, 20, 1000):
            for best in (False, True):
                for reorder in (True, False):
                    kw = {"max_candidates": mc, "best_match": best, "reorder": reorder}
                    if name.startswith("ayx"):
                        kw["ignore_chain"] = "atoms"
                    hits = run(templates, mol, params, kw)
                    cases.append({"name": f"{name}_mc{mc}_best{int(best)}_re{int(reorder)}", "templates": tfiles,
                                  "molecule": mfile, "params": list(params), "kwargs": kw, "hits": hits})
        print(name, "done")
    with open(os.path.join(GOLD, "reference_max_candidates.json"), "w") as f:
        json.dump({"reference": "althonos/pyjess 0.10.1 (baseline/_ref)", "cases": cases}, f)
    print(len(cases), "cases;", sum(len(c["hits"]) for c in cases), "hits")


if __name__ == "__main__":
    main()
