"""A handful of single-structure queries (README example: 3 templates x pdb1a0p) for a kernel launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/gpu/single_query.py"""
import sys

sys.path.insert(0, ".")
from pyjess_b200 import Jess, Molecule, Template                              # noqa: E402

tpl = [Template.load(f"tests/data/template_0{i}.qry") for i in (1, 2, 3)]
mol = Molecule.load("tests/data/pdb1a0p.ent")
jess = Jess(tpl)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for _ in range(n):
    hits = list(jess.query(mol, 2.0, 3.0, 3.0, hit_order=sys.argv[2] if len(sys.argv) > 2 else "reference"))
print(len(hits), "hit(s)")
