"""Cost of the reference iteration order: C-ABI query_host on cfg3 with hit_order canonical /
reference, and the drop-in Python API (`Jess.query_many`, hits materialised lazily)."""
import sys
import time

sys.path.insert(0, ".")
from pyjess_b200 import Jess, _capi, synth                                  # noqa: E402
from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params  # noqa: E402

tpl, mols, thr, kw = synth.workload("cfg3")
ctx = _capi.default_context()
dev = _capi.DeviceTemplates(ctx, TemplateTable(tpl))
for order in ("canonical", "reference", "reference-noreorder"):
    batch = MoleculeBatch(mols, pinned=True, serial=order != "canonical")
    qp = query_params(*thr, hit_order=order, **kw)
    for _ in range(3):
        table, st = _capi.query_host(ctx, dev, batch, qp)
    t0 = time.perf_counter()
    for _ in range(5):
        table, st = _capi.query_host(ctx, dev, batch, qp)
    dt = (time.perf_counter() - t0) / 5
    print(f"query_host hit_order={order}: {dt * 1e3:.2f} ms/step, {table.count} hits, launches {st['launches']}")
qp = query_params(*thr, hit_order="reference", all_tuples=True, **kw)
batch = MoleculeBatch(mols, pinned=True, serial=True)
for _ in range(2):
    table, st = _capi.query_host(ctx, dev, batch, qp)
t0 = time.perf_counter()
table, st = _capi.query_host(ctx, dev, batch, qp)
print(f"query_host all_tuples reference: {(time.perf_counter() - t0) * 1e3:.2f} ms, {table.count} tuples")
jess = Jess(tpl)
for rep in range(3):
    t0 = time.perf_counter()
    qs = jess.query_many(mols, *thr)
    first = next(iter(qs[0]), None)
    t1 = time.perf_counter()
    n = sum(1 for q in qs for _ in q)
    t2 = time.perf_counter()
    print(f"Jess.query_many: first hit after {(t1 - t0) * 1e3:.1f} ms; {n + (first is not None)} Hit objects in {(t2 - t1) * 1e3:.1f} ms")
