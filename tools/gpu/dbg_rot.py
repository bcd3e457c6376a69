import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import oracle
from pyjess_b200 import synth, _capi
from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params
tpl, mols, thr, kw = synth.workload("cfg3dense", scale=0.05)
params = query_params(*thr, reorder=0, **kw)
ctx = _capi.default_context()
tt = TemplateTable(tpl); mb = MoleculeBatch(mols)
dev = _capi.DeviceTemplates(ctx, tt)
a, st = _capi.query_host(ctx, dev, mb, params)
b, _ = oracle.query(tt, mb, params, threads=8)
assert a.keys() == b.keys()
d = np.abs(a.rotation - b.rotation).reshape(len(a.rmsd), -1).max(1)
bad = np.nonzero(d > 1e-9)[0]
print("hits", a.count, "bad>1e-9", len(bad), "max", d.max(), "rmsd maxdiff", np.abs(a.rmsd-b.rmsd).max())
for i in bad[:10]:
    n = len(tpl[int(a.template[i])])
    atoms = a.atoms[i, :n]
    m = mols[int(a.molecule[i])]
    X = m._atoms["x"][atoms]; Y = tt.position[tt.atom_offset[a.template[i]]:tt.atom_offset[a.template[i]]+n]
    Xc = X - X.mean(0); Yc = Y - Y.mean(0)
    C = Xc.T @ Yc
    sv = np.linalg.svd(C, compute_uv=False)
    print(i, a.keys()[i][:2], "n", n, "diff", d[i], "sv", sv, "det", np.linalg.det(C), "detR gpu", np.linalg.det(a.rotation[i]), "cpu", np.linalg.det(b.rotation[i]), "rmsd", a.rmsd[i], b.rmsd[i])
