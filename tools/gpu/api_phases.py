"""Where `Jess.query_many(...)` -> host hit table spends its wall time on a workload (default cfg3):
packing per slice, engine call per slice, host post-processing.   python tools/gpu/api_phases.py [workload] [order]"""
import sys
import time

sys.path.insert(0, ".")
from pyjess_b200 import Jess, _capi, _jess, _tables, synth                   # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
order = sys.argv[2] if len(sys.argv) > 2 else "canonical"
tpl, mols, thr, kw = synth.workload(name)
jess = Jess(tpl)
log = []
orig_qh, orig_mb = _capi.query_records, _tables.MoleculeBatch.__init__


def qh(*a, **k):
    t = time.perf_counter()
    r = orig_qh(*a, **k)
    log.append(("engine", t, time.perf_counter()))
    return r


def mb(self, *a, **k):
    t = time.perf_counter()
    orig_mb(self, *a, **k)
    log.append(("pack", t, time.perf_counter()))


_capi.query_records = qh
_tables.MoleculeBatch.__init__ = mb
for rep in range(6):
    log.clear()
    t0 = time.perf_counter()
    qs = jess.query_many(mols, *thr, hit_order=order, **kw)
    table, rows = qs[0].hit_table()
    t1 = time.perf_counter()
    parts = " ".join(f"{k}[{(a - t0) * 1e3:.2f}-{(b - t0) * 1e3:.2f}]" for k, a, b in sorted(log, key=lambda x: x[1]))
    print(f"rep {rep}: total {(t1 - t0) * 1e3:.2f} ms | {parts} | tables {len(qs[0]._run.tables)} hits {qs[0].stats['hits']}")
    del qs, table
