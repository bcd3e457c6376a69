#!/usr/bin/env python3
"""Cut the rigid residue geometries used by pyjess_b200.synth out of tests/data/1AMY.pdb:
for each of the 20 standard residue types, the heavy atoms of the first COMPLETE
instance, C-alpha centred (SURVEY 8d "Protein generator").  Output:
pyjess_b200/data/residue_library.json (committed)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyjess_b200 import Molecule  # noqa: E402

HEAVY = {
    "ALA": 5, "ARG": 11, "ASN": 8, "ASP": 8, "CYS": 6, "GLN": 9, "GLU": 9, "GLY": 4, "HIS": 10, "ILE": 8,
    "LEU": 8, "LYS": 9, "MET": 8, "PHE": 11, "PRO": 7, "SER": 6, "THR": 7, "TRP": 14, "TYR": 12, "VAL": 7,
}

mol = Molecule.load(os.path.join(ROOT, "tests", "data", "1AMY.pdb"))
residues = {}
for atom in mol:
    key = (atom.chain_id, atom.residue_number, atom.insertion_code, atom.residue_name)
    residues.setdefault(key, []).append(atom)
lib = {}
for (chain, num, icode, res), atoms in residues.items():
    if res not in HEAVY or res in lib:
        continue
    names = [a.name for a in atoms if a.altloc in (" ", "A")]
    if len(names) != HEAVY[res] or len(set(names)) != len(names) or "CA" not in names:
        continue
    ca = next(a for a in atoms if a.name == "CA")
    lib[res] = {"names": names,
                "xyz": [[round(a.x - ca.x, 3), round(a.y - ca.y, 3), round(a.z - ca.z, 3)] for a in atoms
                        if a.altloc in (" ", "A")]}
assert len(lib) == 20, sorted(set(HEAVY) - set(lib))
with open(os.path.join(ROOT, "pyjess_b200", "data", "residue_library.json"), "w") as f:
    json.dump(lib, f, indent=0, sort_keys=True)
print({k: len(v["names"]) for k, v in sorted(lib.items())})
