"""Where a single small query spends its time (README example: 3 templates x pdb1a0p)."""
import sys
import time

sys.path.insert(0, ".")
from pyjess_b200 import Jess, Molecule, Template, _capi                      # noqa: E402
from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params    # noqa: E402

tpl = [Template.load(f"tests/data/template_0{i}.qry") for i in (1, 2, 3)]
mol = Molecule.load("tests/data/pdb1a0p.ent")
ctx = _capi.default_context()
dev = _capi.DeviceTemplates(ctx, TemplateTable(tpl))


def timeit(fn, n=200):
    fn()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t) / n * 1e3


for order in ("canonical", "reference"):
    qp = query_params(2.0, 3.0, 3.0, hit_order=order)
    batch = MoleculeBatch([mol], serial=True)
    ms = timeit(lambda: _capi.query_host(ctx, dev, batch, qp))
    table, st = _capi.query_host(ctx, dev, batch, qp)
    print(f"query_host hit_order={order}: {ms:.3f} ms wall, kernel_ms {st['kernel_ms']:.3f}, launches {st['launches']}, hits {table.count}")
print(f"MoleculeBatch([mol], serial=True, pool): {timeit(lambda: MoleculeBatch([mol], serial=True, pool={})):.3f} ms")
jess = Jess(tpl)
print(f"Jess.query + list(): {timeit(lambda: list(jess.query(mol, 2.0, 3.0, 3.0))):.3f} ms")
print(f"Jess.query (no hits, rmsd 0.1) + list(): {timeit(lambda: list(jess.query(mol, 0.1, 3.0, 3.0))):.3f} ms")
