#!/usr/bin/env python3
"""Generate the golden fixtures of tests/golden/ from the REFERENCE itself.

Run in the build container only (needs /root/reference and the reference's Python
package installed in baseline/_ref):

    python tests/golden/make_golden.py

Outputs (all committed):
  tests/data/*                      small reference fixtures (templates, PDB files,
                                    the original-Jess golden output 1AMY+1.3.3.txt)
  tests/golden/cfg2_structures.npz  packed atoms of the 132 vendor example structures
                                    as parsed by the REFERENCE's Molecule.load
  tests/golden/reference_hits.json  hits of the reference's Jess.query for every case
                                    in CASES (full FP64 values, not 3 dp)
Nothing under tests/ reads /root/reference at run time.
"""
import glob
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("JESS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, ROOT)

import pyjess  # the reference  # noqa: E402

DATA = os.path.join(ROOT, "tests", "data")
GOLD = os.path.join(ROOT, "tests", "golden")
EX = os.path.join(REF, "vendor", "jess", "examples")
RT = os.path.join(REF, "src", "pyjess", "tests", "data")


def copy_fixtures():
    os.makedirs(DATA, exist_ok=True)
    for name in ("template_01.qry", "template_02.qry", "template_03.qry"):
        shutil.copyfile(os.path.join(EX, name), os.path.join(DATA, name))
    for name in ("pdb1a0p.ent", "pdb1lnb.ent"):
        shutil.copyfile(os.path.join(EX, "test_pdbs", name), os.path.join(DATA, name))
    for name in ("1.3.3.tpl", "4.1.2.tpl", "1sur.qry", "1AMY.pdb", "5ayx.EF.pdb", "1AMY+1.3.3.txt"):
        shutil.copyfile(os.path.join(RT, name), os.path.join(DATA, name))
    for name in os.listdir(DATA):
        os.chmod(os.path.join(DATA, name), 0o644)


def tok(s: str, n: int) -> bytes:
    return s.encode("ascii")


def pack_reference_molecule(mol):
    """Field arrays of a reference Molecule, tokens re-encoded the way the
    reference stores them internally (Atom.name strips the underscores, so we
    rebuild the 4-char token from the PDB column rule)."""
    n = len(mol)
    out = {
        "xi": np.zeros(n, np.int32), "yi": np.zeros(n, np.int32), "zi": np.zeros(n, np.int32),
        "serial": np.zeros(n, np.int32), "resseq": np.zeros(n, np.int32),
        "chain1": np.zeros(n, np.uint8), "chain2": np.zeros(n, np.uint8),
        "name": np.zeros(n, "S4"), "resname": np.zeros(n, "S3"),
    }
    for i, a in enumerate(mol):
        line = a.dumps()  # reference writes the raw 4-char name column back
        out["name"][i] = line[12:16].replace(" ", "_").encode("ascii")
        out["resname"][i] = line[17:20].replace(" ", "_").encode("ascii")
        for k, v in (("xi", a.x), ("yi", a.y), ("zi", a.z)):
            iv = int(round(v * 1000.0))
            assert iv / 1000.0 == v, (v, iv)
            out[k][i] = iv
        out["serial"][i] = a.serial
        out["resseq"][i] = a.residue_number
        cid = a.chain_id
        out["chain1"][i] = ord(cid[0]) if len(cid) == 2 else 0x20
        out["chain2"][i] = ord(cid[-1]) if cid else 0x30
    return out


def make_cfg2_npz():
    files = sorted(glob.glob(os.path.join(EX, "test_pdbs", "*.ent")))
    cols, ids, offs = {}, [], [0]
    for path in files:
        mol = pyjess.Molecule.load(path)
        p = pack_reference_molecule(mol)
        for k, v in p.items():
            cols.setdefault(k, []).append(v)
        ids.append(os.path.basename(path))
        offs.append(offs[-1] + len(mol))
    arrays = {k: np.concatenate(v) for k, v in cols.items()}
    arrays["atom_offset"] = np.asarray(offs, np.int64)
    arrays["ids"] = np.asarray(ids)
    np.savez_compressed(os.path.join(GOLD, "cfg2_structures.npz"), **arrays)
    return files


def atom_index(mol):
    idx = {}
    for i, a in enumerate(mol):
        key = (a.serial, a.name, a.residue_number, a.chain_id, a.x, a.y, a.z)
        assert key not in idx, key
        idx[key] = i
    return idx


def run_case(templates, mol, rmsd, dist, maxdyn, **kw):
    jess = pyjess.Jess(templates)
    index = atom_index(mol)
    tid = {id(t): i for i, t in enumerate(templates)}
    hits = []
    for hit in jess.query(mol, rmsd, dist, maxdyn, **kw):
        atoms = [index[(a.serial, a.name, a.residue_number, a.chain_id, a.x, a.y, a.z)]
                 for a in hit.atoms(transform=False)]
        m = hit.transformation
        mat = [[float(memoryview(m)[i, j]) for j in range(4)] for i in range(4)]
        hits.append({"template": tid[id(hit.template())], "atoms": atoms, "rmsd": hit.rmsd,
                     "transformation": mat, "determinant": hit.determinant, "log_evalue": hit.log_evalue})
    return hits


def main():
    copy_fixtures()
    files = make_cfg2_npz()
    load_t = lambda name: pyjess.Template.load(os.path.join(DATA, name))  # noqa: E731
    load_m = lambda name: pyjess.Molecule.load(os.path.join(DATA, name))  # noqa: E731
    ex = [load_t(f"template_0{i}.qry") for i in (1, 2, 3)]
    cases = []

    def add(name, tpl_files, mol_file, params, kw, templates=None, mol=None):
        templates = templates or [load_t(f) for f in tpl_files]
        mol = mol or load_m(mol_file)
        hits = run_case(templates, mol, *params, **kw)
        cases.append({"name": name, "templates": tpl_files, "molecule": mol_file, "params": list(params),
                      "kwargs": kw, "n_hits": len(hits), "hits": hits})
        print(f"{name}: {len(hits)} hits")

    ex_files = ["template_01.qry", "template_02.qry", "template_03.qry"]
    add("cfg1_examples_1a0p", ex_files, "pdb1a0p.ent", (2.0, 3.0, 3.0), {})
    for p in ((1, 2, 2), (2, 5, 3), (2, 5, 5)):
        add(f"lnb_t01_{p[0]}_{p[1]}_{p[2]}", ["template_01.qry"], "pdb1lnb.ent", tuple(map(float, p)), {})
    add("amy_133_reorder", ["1.3.3.tpl"], "1AMY.pdb", (2.0, 4.0, 4.0), {"reorder": True})
    add("amy_133_noreorder", ["1.3.3.tpl"], "1AMY.pdb", (2.0, 4.0, 4.0), {"reorder": False})
    for ic in (None, "residues", "atoms"):
        add(f"5ayx_412_ignore_{ic}", ["4.1.2.tpl"], "5ayx.EF.pdb", (3.0, 3.0, 3.0), {"ignore_chain": ic})
    add("amy_doctest", ["1.3.3.tpl", "4.1.2.tpl"], "1AMY.pdb", (2.0, 2.0, 2.0), {})
    add("amy_doctest_best", ["1.3.3.tpl", "4.1.2.tpl"], "1AMY.pdb", (2.0, 2.0, 2.0), {"best_match": True})
    add("lnb_t01_x3_best", ["template_01.qry"] * 3, "pdb1lnb.ent", (1.0, 2.0, 2.0), {"best_match": True})
    add("amy_examples_split", ["template_01.qry", "template_02.qry"], "1AMY.pdb", (2.0, 5.0, 5.0), {})
    # single-atom templates (the reference never reports the first entry of ITS candidate list):
    # exact name, alternate residue names listed BEFORE/AFTER the primary one, a name-only mode
    singles = {
        "single_his.qry": "ATOM      0  ND1 HIS E 142       4.191  -2.584   4.209\n",
        "single_glu_asp.qry": "ATOM      3  OE1 GLU E 143       0.491  -4.480  -0.456 D\n",
        "single_any_ca.qry": "ATOM    100  CA  GLY E   1       0.000   0.000   0.000\n",
    }
    for fname, text in singles.items():
        with open(os.path.join(DATA, fname), "w") as f:
            f.write(text)
    for fname in singles:
        for mfile in ("pdb1lnb.ent", "5ayx.EF.pdb"):
            add(f"{fname[:-4]}_{mfile.split('.')[0]}", [fname], mfile, (2.0, 2.0, 2.0), {})
            cases[-1]["single_atom"] = True
    add("single_mixed_lnb", ["single_his.qry", "template_01.qry", "single_glu_asp.qry"], "pdb1lnb.ent",
        (2.0, 5.0, 5.0), {})
    cases[-1]["single_atom"] = True
    # ... with best_match: every single-atom tuple has RMSD 0, so the reference keeps the FIRST candidate it
    # reports -- its candidate #1, candidate #0 being the one it never reports
    for fname in singles:
        add(f"{fname[:-4]}_pdb1lnb_best", [fname], "pdb1lnb.ent", (2.0, 2.0, 2.0), {"best_match": True})
        cases[-1]["single_atom"] = True
    add("single_mixed_lnb_best", ["single_his.qry", "template_01.qry", "single_glu_asp.qry"], "pdb1lnb.ent",
        (2.0, 5.0, 5.0), {"best_match": True})
    cases[-1]["single_atom"] = True

    # cfg 2: the three example templates against all 132 structures
    cfg2 = []
    for m, path in enumerate(files):
        mol = pyjess.Molecule.load(path)
        for h in run_case(ex, mol, 2.0, 3.0, 3.0):
            h["molecule"] = m
            cfg2.append(h)
    print("cfg2:", len(cfg2), "hits")
    cases.append({"name": "cfg2_examples_all", "templates": ex_files, "molecule": "cfg2_structures.npz",
                  "params": [2.0, 3.0, 3.0], "kwargs": {}, "n_hits": len(cfg2), "hits": cfg2})

    with open(os.path.join(GOLD, "reference_hits.json"), "w") as f:
        json.dump({"reference": "althonos/pyjess 0.10.1 (baseline/_ref)", "cases": cases}, f)


if __name__ == "__main__":
    main()
