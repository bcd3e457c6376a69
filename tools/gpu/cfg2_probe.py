"""BASELINE config 2 (3 vendor templates x the 132 vendor structures): where a batch spends its time, and the
README's thread-pool recipe (one `Jess.query` per structure from a pool of threads) next to `query_many`."""
import os
import sys
import time
from multiprocessing.pool import ThreadPool

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import cfg2_molecules, load_template                                # noqa: E402
from pyjess_b200 import Jess, _capi                                            # noqa: E402
from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params     # noqa: E402

tpl = [load_template(f"template_0{i}.qry") for i in (1, 2, 3)]
mols = cfg2_molecules()
thr = (2.0, 3.0, 3.0)
print(f"{len(mols)} structures, {sum(len(m) for m in mols)} atoms, largest {max(len(m) for m in mols)}")
ctx = _capi.default_context()
dev = _capi.DeviceTemplates(ctx, TemplateTable(tpl))


def best(fn, reps=20, rounds=3):
    for _ in range(5):
        fn()
    b = 1e30
    for _ in range(rounds):
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        b = min(b, (time.perf_counter() - t) / reps * 1e3)
    return b


for order in ("canonical", "reference"):
    qp = query_params(*thr, hit_order=order)
    batch = MoleculeBatch(mols, serial=True, pinned=True)
    ms = best(lambda: _capi.query_host(ctx, dev, batch, qp))
    table, st = _capi.query_host(ctx, dev, batch, qp)
    keys = ("kernel_ms", "classify_ms", "pairs_ms", "expand_ms", "superpose_ms", "launches", "chunks")
    print(f"jb_query_host {order}: {ms:.3f} ms wall; " + ", ".join(f"{k} {st[k]:.3f}" if isinstance(st[k], float) else f"{k} {st[k]}" for k in keys if k in st),
          f"hits {table.count}")
    dm = _capi.DeviceMolecules(ctx, batch)
    def resident():
        r = _capi.query_run(ctx, dev, dm, qp); r.close()
    print(f"  resident jb_query_run {order}: {best(resident):.3f} ms wall")
    dm.close()
jess = Jess(tpl)
for order in ("canonical", "reference"):
    print(f"query_many {order} -> hit_table: {best(lambda: jess.query_many(mols, *thr, hit_order=order)[0].hit_table()):.3f} ms")
    print(f"query_many {order} -> list() of every query: {best(lambda: [list(q) for q in jess.query_many(mols, *thr, hit_order=order)], 10):.3f} ms")
print(f"serial loop of Jess.query: {best(lambda: [list(jess.query(m, *thr)) for m in mols], 3):.3f} ms")
for n in (4, 16):
    with ThreadPool(n) as pool:
        ms = best(lambda: pool.map(lambda m: list(jess.query(m, *thr)), mols), 5)
        tot = sum(len(h) for h in pool.map(lambda m: list(jess.query(m, *thr)), mols))
    print(f"ThreadPool({n}).map(Jess.query): {ms:.3f} ms, {tot} hits")
